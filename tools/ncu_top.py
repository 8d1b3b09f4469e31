"""Key metrics + top stalled SASS instructions of every kernel in a .ncu-rep (needs -lineinfo / --import-source).
Usage: python tools/ncu_top.py file.ncu-rep [ntop]"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size"]
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70])
    for k in KEYS:
        if k in hdr:
            print(f"   {k:90s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.25:
                print(f"      stall {h[34:-23]:24s} {v:.2f}")
src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and len(r) >= len(cur["hdr"]) - 2:
        cur["rows"].append(r)
for sec in secs:
    h = sec["hdr"]
    iS, iSrc, iE = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    tot = sum(int(r[iS] or 0) for r in sec["rows"])
    print(f"== {sec['name'][:60]}: {tot} samples, {len(sec['rows'])} SASS instructions")
    for r in sorted(sec["rows"], key=lambda r: -int(r[iS] or 0))[:ntop]:
        print(f"   {int(r[iS] or 0):6d} ({100 * int(r[iS] or 0) / max(tot, 1):4.1f}%) x{r[iE]:>9s}  {r[iSrc][:90]}")
