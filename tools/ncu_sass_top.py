"""Top stalled SASS instructions of one kernel of an ncu report (source page, CSV):
  ncu -i rep.ncu-rep --page source --csv --kernel-id ::regex:<name>:<nth> --print-source sass > k.csv
  python tools/ncu_sass_top.py k.csv [n] [which kernel of the export]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # the export may hold several kernels: pick one
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[which]
hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
hdr = rows[lo + 1]
data = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
rows = rows[lo:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in data)
agg = {}
for r in data:
    for c in stall_cols:
        agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c] or 0)
print(rows[0][1])
print("total samples", tot, "sass instructions", len(data))
print("stall mix:", ", ".join(f"{k[6:]} {v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:n]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} samples {r[isamp]:>6} executed {r[iex]:>9}  {r[isrc][:70]:70s} {st}")
