#!/usr/bin/env python3
"""Turn `ncu -i <rep> --page raw --csv` of tools/prof_edge.py into the markdown table kept under profiles/."""
import csv, sys
raw, E = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
names = [d[hdr.index("Kernel Name")].replace("void ", "").replace("(Params)", "") for d in data]
metrics = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print("| metric | unit | " + " | ".join(names) + " |")
print("|---|---|" + "---:|" * len(names))
for m in metrics:
    if m in hdr:
        i = hdr.index(m)
        print(f"| `{m}` | {units[i]} | " + " | ".join(d[i] for d in data) + " |")
ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
per = [(float(d[ir]) * SCALE[units[ir]] + float(d[iw]) * SCALE[units[iw]]) / E for d in data]
print("| DRAM bytes per edge (read+write) | B | " + " | ".join(f"{x:.1f}" for x in per) + " |")
print()
print("| stall reason (warps per issue-active cycle) | " + " | ".join(str(i + 1) for i in range(len(names))) + " |")
print("|---|" + "---:|" * len(names))
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        v = [float(d[i]) for d in data]
        if max(v) > 0.2:
            print(f"| {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | " + " | ".join(f"{x:.2f}" for x in v) + " |")
