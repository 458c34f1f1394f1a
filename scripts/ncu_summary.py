"""Summarise an ncu report (.ncu-rep) + launch list (.csv) into a small text file for profiles/."""
import csv, subprocess, sys, io
rep, launches, out = sys.argv[1], sys.argv[2], sys.argv[3]
names = {0: "gates", 1: "propose", 2: "decode", 3: "lngelu(trunk7/trunk1)", 11: "lngelu_b2b(trunk: 7x7+LN+GELU+1x1+LN+GELU)", 4: "mix", 5: "bias_lrelu(q1/q3)", 6: "res_proj(q2)", 7: "res_id(q4)", 8: "sample(q5)"}
lines = []
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
tot = sum(float(r[vi].replace(",", "")) for r in rows[1:])
lines.append("# launch list of ONE event (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)")
for r in rows[1:]:
    t = float(r[vi].replace(",", ""))
    lines.append(f"{r[ki][:64]:64s} {t/1e3:9.1f} us  {100*t/tot:5.1f} %")
lines.append(f"{'total':64s} {tot/1e3:9.1f} us")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex_op_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__block_size", "launch__grid_size"]
idx = {w: hdr.index(w) for w in want if w in hdr}
units = rr[1]
lines.append("")
lines.append("# ncu --set full (one launch each): " + ", ".join(f"{w} [{units[i]}]" for w, i in idx.items()))
for r in rr[2:]:
    lines.append(r[hdr.index("Kernel Name")][:48] + " | " + " | ".join(r[i] for i in idx.values()))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
