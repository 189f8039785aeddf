#!/usr/bin/env python
"""Print selected metrics of every kernel in an .ncu-rep:  python profiles/ncu_metrics.py rep [regex...]"""
import csv, re, subprocess, sys
DEFAULT = [r"^gpu__time_duration.sum$", r"^dram__bytes_(read|write).sum$", r"^gpu__dram_throughput.avg.pct", r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           r"^sm__throughput.avg.pct", r"^launch__registers_per_thread$", r"^sm__warps_active.avg.pct", r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct",
           r"l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct", r"^lts__t_bytes.sum$", r"^lts__throughput.avg.pct", r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$",
           r"^sm__inst_executed_pipe_fma.*pct_of_peak_sustained_active", r"^smsp__inst_executed.sum$", r"^sm__cycles_elapsed.max$", r"lts__t_sector_hit_rate.pct",
           r"^l1tex__throughput.avg.pct", r"sm__inst_executed_pipe_xu.*pct_of_peak_sustained_active$"]
rep = sys.argv[1]
pats = [re.compile(p) for p in (sys.argv[2:] or DEFAULT)]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])
    print("====", name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(p.search(h) for p in pats):
            print(f"  {h:80s} {r[i]:>16s} {units[i]}")
