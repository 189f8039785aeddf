#!/usr/bin/env python
"""Hottest SASS instructions (warp-stall samples) of one kernel launch in an .ncu-rep.
   python profiles/ncu_hot.py rep kernel_regex [launch_index] [top_n]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
sections, cur = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and len(r) > 10:
        if cur["hdr"] is None:
            cur["hdr"] = r
        else:
            cur["data"].append(r)
sec = sections[0]
hdr, data = sec["hdr"], sec["data"]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print("kernel", pat, "instructions", len(data), "samples", tot)
idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top]
for i in sorted(idx):
    r = data[i]
    why = sorted(((int(r[j]), hdr[j][6:]) for j in stalls if int(r[j]) > 0), reverse=True)[:2]
    print(f"{i:5d} {int(r[isamp]):7d} {100*int(r[isamp])/max(tot,1):5.1f}% ex={r[iex]:>8s} {r[ia].strip()[:70]:70s} {why}")
