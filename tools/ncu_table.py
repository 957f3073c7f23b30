#!/usr/bin/env python3
"""Markdown-ish table of the key metrics of every kernel in an `ncu -i X.ncu-rep --page raw --csv` export.

usage: ncu_table.py raw.csv EDGES TILES
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
E, T = int(sys.argv[2]), int(sys.argv[3])
hdr, units, data = rows[0], rows[1], rows[2:]
names = [d[hdr.index("Kernel Name")].split('(')[0].replace('void ', '').replace('pev::', '').replace('tc2::', '')[:14] for d in data]
print(f"{'metric':58s} {'unit':7s}", " ".join(f"{n:>14s}" for n in names))
metrics = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
for m in metrics:
    if m in hdr:
        i = hdr.index(m)
        print(f"{m[:58]:58s} {units[i][:7]:7s}", " ".join(f"{d[i][:12]:>14s}" for d in data))
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
print(f"{'DRAM bytes per edge (read+write)':58s} {'B':7s}",
      " ".join(f"{(float(d[ir]) * SC[units[ir]] + float(d[iw]) * SC[units[iw]]) / E:14.1f}" for d in data))
ii = hdr.index("smsp__inst_executed.sum")
print(f"{'warp instructions per 128-edge tile':58s} {'':7s}", " ".join(f"{float(d[ii]) / T:14.0f}" for d in data))
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        v = [float(d[i]) for d in data]
        if max(v) > 0.3:
            print(f"{'stall ' + h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:58s} {'':7s}", " ".join(f"{x:14.2f}" for x in v))
