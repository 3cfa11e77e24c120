"""Top stall-sample SASS lines of an `ncu --page source --csv` export. usage: ncu_hot.py file.csv [N]"""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = r[1]
ci, cs, ce = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
rows = [x for x in r[2:] if len(x) > ci]
tot = sum(float(x[ci] or 0) for x in rows) or 1
idx = {id(x): i for i, x in enumerate(rows)}
for x in sorted(rows, key=lambda x: -float(x[ci] or 0))[:n]:
    top = sorted(((float(x[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{float(x[ci])/tot*100:5.1f}%  line {idx[id(x)]:5d} exec={x[ce]:>9s}  {x[cs].strip()[:70]:70s} {top[0][1]}={top[0][0]:.0f} {top[1][1]}={top[1][0]:.0f}")
