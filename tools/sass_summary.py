#!/usr/bin/env python3
"""Per-kernel SASS evidence of libhj_b200.so (runs without a GPU): instruction counts of the mnemonics
that prove the Blackwell-native paths — UBLKCP (cp.async.bulk, the 1-D TMA path), SYNCS (mbarrier),
REDUX / CREDUX (warp reduce), IDP.4A (dp4a lane masks), ATOMS / REDG / ATOMG (shared / global atomics),
ACQBULK / PREEXIT (griddepcontrol.wait / .launch_dependents: programmatic dependent launch), .SYS
(system-scope loads / stores: the peer-memory exchange over NVLink), NANOSLEEP (polling back-off).
No UTMALDG (the streams are 1-D: bulk copies need no tensor map) and no HMMA / UTC*MMA (nothing here is a
contraction).  The NVRTC-generated fused kernels are not in the .so; tools/ir_sass.py dumps theirs.

    python tools/sass_summary.py > profiles/sass_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "hephaestus-jit_b200", "libhj_b200.so")
PATTERNS = ["UBLKCP", "SYNCS", "REDUX", "CREDUX", "IDP.4A", "ATOMS", "REDG", "ATOMG", "ACQBULK", "PREEXIT", "NANOSLEEP", ".SYS", "LDG", "STG", "LDS", "STS", "SHFL", "UTMALDG", "HMMA", "UTC"]

sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
kernels = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"hj::\(anonymous namespace\)::", "", name)
        kernels[name] = collections.Counter()
        continue
    if name and re.search(r"/\*[0-9a-f]{4}\*/", line):
        kernels[name]["instructions"] += 1
        for p in PATTERNS:
            if p in line:
                kernels[name][p] += 1
print(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)}: {len(kernels)} kernels (sm_100a)")
print("# per kernel: total SASS instructions, then the count of lines containing each mnemonic")
tot = collections.Counter()
for k, c in kernels.items():
    short = k if len(k) < 150 else k[:147] + "..."
    cells = "  ".join(f"{p}={c[p]}" for p in PATTERNS if c[p])
    print(f"{short}\n    instr={c['instructions']}  {cells}")
    tot.update(c)
print("\n# totals over all kernels")
print("  ".join(f"{p}={tot[p]}" for p in ["instructions"] + PATTERNS))
