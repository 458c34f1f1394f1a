"""profiles/roofline_traffic.json from an ncu --set full report of ONE event (scripts/gpu_profile.sh): per conv stage,
dram__bytes_read.sum + dram__bytes_write.sum of its launch, keyed '<stage>@<grid>x<batch>' as bench.py looks it up."""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
grid, batch = (sys.argv[3], sys.argv[4]) if len(sys.argv) > 4 else ("200", "8")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, units = rr[0], rr[1]
rows = [r for r in rr[2:] if "conv_stage_kernel" in r[h.index("Kernel Name")]]
names = ["gates", "propose", "decode", "trunk", "mix", "q1", "q2", "q3", "q4", "q5"]      # launch order of one event (fused trunk stage)
assert len(rows) == len(names), (len(rows), "conv stage launches in the report")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
d = {}
for n, r in zip(names, rows):
    d[f"{n}@{grid}x{batch}"] = int(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d, indent=1))
