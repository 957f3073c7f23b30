#!/usr/bin/env python3
"""Hot spots of one kernel from an `ncu --page source --csv` export: top instructions by stall samples, with reasons.

usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv ; ncu_hot.py src.csv [top]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # which kernel section of the export
rows = rows[starts[sec]:(starts[sec + 1] if sec + 1 < len(starts) else len(rows))]
print(rows[0][1])
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
inst = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot}, warp instructions executed {inst}")
agg = {}
for r in data:
    for n in stall_cols:
        agg[n] = agg.get(n, 0) + int(r[col[n]] or 0)
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:top]
for i in sorted(idx):
    r = data[i]
    s = int(r[col["# Samples"]] or 0)
    reasons = sorted(((int(r[col[n]] or 0), n[6:]) for n in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {s:6d} {100.0 * s / tot:5.1f}%  ex={r[col['Instructions Executed']]:>8}  {r[col['Source']].strip()[:70]:70s} "
          + " ".join(f"{n}={v}" for v, n in reasons if v))
