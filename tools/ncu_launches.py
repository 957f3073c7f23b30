#!/usr/bin/env python3
"""Share-of-GPU-time table from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

usage: ncu_launches.py launches.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
n = 0
for r in rows[hi + 1:]:
    if len(r) != len(h):
        continue
    v = float(r[mv].replace(",", ""))
    ms = v / 1e6 if r[mu].startswith("n") else v / 1e3 if r[mu].startswith("u") else v
    agg[r[kn]][0] += 1
    agg[r[kn]][1] += ms
    n += 1
tot = sum(v[1] for v in agg.values())
print(f"Total kernel time {tot:.1f} ms over {n} launches.\n")
print("| share | total ms | launches | ms / launch | kernel |\n|---:|---:|---:|---:|---|")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"| {100 * ms / tot:.1f}% | {ms:.2f} | {c} | {ms / c:.3f} | `{k[:110]}` |")
tc2 = sum(v[1] for k, v in agg.items() if "tc2::" in k)
own = sum(v[1] for k, v in agg.items() if "tc2::" in k or "pev::" in k)
print(f"\n`tc2::*` {100 * tc2 / tot:.1f} % of GPU time, all `pev::`/`tc2::` kernels {100 * own / tot:.1f} %.")
