"""Per-kernel totals (time, DRAM read/write bytes) of the LAST step in an ncu csv log taken with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum.
Usage: summarize_traffic.py file.csv [nsteps | first-kernel-of-a-step]
With a kernel name (e.g. volume_stats_kernel) the step is the launches between its last two occurrences."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
arg = sys.argv[2] if len(sys.argv) > 2 else "2"
nsteps = int(arg) if arg.isdigit() else 0
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ii, ki, mi, vi, ui = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = {}
for r in rows[start:]:
    if len(r) <= vi:
        continue
    d = launches.setdefault(int(r[ii]), {"name": r[ki]})
    v = float(r[vi].replace(",", ""))
    u = r[ui].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(u, 1)
    d[r[mi]] = v * scale
ids = sorted(launches)
if nsteps:
    step = ids[len(ids) - len(ids) // nsteps:]
else:
    marks = [i for i in ids if arg in launches[i]["name"]]
    step = [i for i in ids if marks[-2] <= i < marks[-1]]
agg = {}
for i in step:
    d = launches[i]
    name = d["name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    a = agg.setdefault(name, [0.0, 0.0, 0.0, 0])
    a[0] += d.get("gpu__time_duration.sum", 0.0)
    a[1] += d.get("dram__bytes_read.sum", 0.0)
    a[2] += d.get("dram__bytes_write.sum", 0.0)
    a[3] += 1
tot = sum(a[0] for a in agg.values())
print(f"launches in the profiled step: {len(step)}, total {tot / 1e6:.3f} ms (serialised, cold caches)")
print(f"{'ms':>9s} {'share':>6s} {'n':>4s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s}  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{a[0] / 1e6:9.3f} {100 * a[0] / tot:5.1f}% x{a[3]:<3d} {a[1] / 1e6:11.1f} {a[2] / 1e6:11.1f}  {k[:70]}")
