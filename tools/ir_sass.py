#!/usr/bin/env python3
"""SASS of the NVRTC-compiled headline kernel (BASELINE C2 chain), compiled here without a GPU:
mnemonic histogram of hj_kernel_vec — 128-bit LDG / STG, MUFU (sin / ex2), ACQBULK / PREEXIT (PDL)."""
import collections, importlib, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
irm = importlib.import_module("hephaestus-jit_b200.ir")
cubin = irm.compile_cubin(irm.c2_chain_ir().build())
with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
    f.write(cubin); f.flush()
    sass = subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True, check=True).stdout
name, hist = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); hist[name] = collections.Counter(); continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]+ )?([A-Z][A-Z0-9_.]*)", line)
    if name and m:
        hist[name][m.group(1)] += 1
for k, c in hist.items():
    print(f"{k}: {sum(c.values())} instructions")
    print("   " + "  ".join(f"{op}={n}" for op, n in c.most_common(40)))
