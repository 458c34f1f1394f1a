"""Groups the ncu launch list of one forward (scripts/profile_forward.py) by kernel and by position: python scripts/forward_launch_summary.py in.csv out.txt"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
L = [(r[ki], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:]]
tot = sum(t for _, t in L)
by = collections.OrderedDict()
for k, t in L:
    k = k[:70]
    c = by.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += t
out = [f"# one eager FuturePredictionODE.forward, B = 8, BEV 200x200x64: {len(L)} launches, {tot/1e3:.2f} ms of kernel time (ncu, serialised, cold caches)"]
for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k:70s} x{n:4d} {t/1e3:8.3f} ms {100*t/tot:5.1f} %")
out.append("")
out.append("# in launch order (index, us, kernel)")
for i, (k, t) in enumerate(L):
    out.append(f"{i:4d} {t:9.1f} {k[:60]}")
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
