#!/usr/bin/env python
"""Turns gpurun_out ncu outputs into the tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv profiles/r1_launches.md "<command>"
  python profiles/summarize.py raw gpurun_out/prof.ncu-rep profiles/r1_kernels.csv
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def launches(src, dst, cmd):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", row["Kernel Name"])).replace("<unnamed>::", "")
        key = (name, row["Grid Size"], row["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n`{cmd}`\n\n")
        f.write("| kernel | grid | block | launches | total ms | avg us | share |\n|---|---|---|---:|---:|---:|---:|\n")
        for (name, grid, block), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {grid} | {block} | {n} | {t / 1e6:.3f} | {t / n / 1e3:.2f} | {100 * t / tot:.1f}% |\n")
        f.write(f"\ntotal {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([re.sub(r"\(.*", "", r[i]) if hdr[i] == "Kernel Name" else r[i] for i in idx])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        raw(sys.argv[2], sys.argv[3])
