"""Summarize an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the
LAST `nsteps`-th part of the log (the profiled step). Usage: summarize_launches.py file.csv [nsteps] [-v]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 2
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = [(r[ki], float(r[vi].replace(",", "")), r[ui]) for r in rows[start:] if len(r) > vi]
step = data[len(data) - len(data) // nsteps:]
tot = sum(v for _, v, _ in step)
print(f"launches in the profiled step: {len(step)}, total {tot / 1e6:.3f} ms")
agg = {}
for i, (k, v, u) in enumerate(step):
    name = k.split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    if "-v" in sys.argv:
        print(f"{i:3d} {v / 1000:10.1f} us  {name[:70]}")
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v
    a[1] += 1
for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v / 1e6:9.3f} ms {100 * v / tot:5.1f}%  x{n:<3d} {k[:80]}")
