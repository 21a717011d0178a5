#!/usr/bin/env python3
"""Top SASS instructions of an .ncu-rep source page by executed count and by stall samples.
    python tools/ncu_hot.py report.ncu-rep [N]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
ie, ss = ci["Instructions Executed"], ci["Warp Stall Sampling (All Samples)"]
tot_i = sum(int(r[ie]) for r in data)
tot_s = sum(int(r[ss]) for r in data)
print(f"total warp instructions {tot_i}, stall samples {tot_s}, {len(data)} SASS lines")
print("--- by instructions executed")
for k, r in sorted(enumerate(data), key=lambda x: -int(x[1][ie]))[:n]:
    print(f"{k:5d} {int(r[ie]):10d} {100*int(r[ie])/tot_i:5.1f}%  samples {int(r[ss]):6d}  {r[1].strip()[:80]}")
print("--- by stall samples")
for k, r in sorted(enumerate(data), key=lambda x: -int(x[1][ss]))[:n]:
    print(f"{k:5d} {int(r[ie]):10d}  samples {int(r[ss]):6d} {100*int(r[ss])/tot_s:5.1f}%  {r[1].strip()[:80]}")
