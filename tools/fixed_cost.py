#!/usr/bin/env python3
"""Fixed cost per launch of the hot kernels: time(n) = a + n / bw, fitted over n = 2^25 .. 2^30.

The per-GPU share of an 8-way split (2^27 elements) loses to `a`, not to bandwidth;
this prints `a` (us) and the asymptotic bandwidth for every kernel, two ways:
  isolated   one launch between two events (what a suite line measures)
  train      20 launches back to back between two events, divided by 20 (what a step of passes sees:
             with programmatic dependent launch the next kernel's prologue overlaps this one's tail)
Run it twice, with and without HJ_NO_PDL=1, to see what PDL buys.
Usage: python tools/fixed_cost.py [min_log2 max_log2]
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
irm = importlib.import_module("hephaestus-jit_b200.ir")

lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (25, 30)
torch.cuda.set_device(0)
dev = hj.Device.cuda(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
dev.set_stream(stream.cuda_stream)
wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
g = torch.Generator(device="cuda").manual_seed(0)
nmax = 1 << hi
xf = torch.rand(nmax, device="cuda", generator=g, dtype=torch.float32)
xu = torch.randint(0, 4, (nmax,), device="cuda", generator=g, dtype=torch.int32)
out = torch.empty(nmax, device="cuda", dtype=torch.int32)
m50 = (torch.rand(nmax, device="cuda", generator=g) < 0.5).to(torch.uint8)
m01 = (torch.rand(nmax, device="cuda", generator=g) < 0.01).to(torch.uint8)
keys = torch.randint(0, 1 << 16, (nmax,), device="cuda", generator=g, dtype=torch.int32)
hist = torch.zeros(1 << 16, device="cuda", dtype=torch.int32)
o1 = torch.zeros(16, device="cuda", dtype=torch.float32)
cnt = torch.zeros(4, device="cuda", dtype=torch.int32)
bf, bu, bo, b50, b01, bk, bh, b1, bc = map(wrap, (xf, xu, out, m50, m01, keys, hist, o1, cnt))
kern = dev.kernel(irm.c2_chain_ir())
yf = torch.empty(min(nmax, 1 << 28), device="cuda", dtype=torch.float32)
by = wrap(yf)

ops = {
    "map C2 (8 B/elem)": (lambda n: dev.launch(kern, n, [bf, by]), 8, 28),
    "reduce sum f32 (4 B/elem)": (lambda n: dev.reduce(hj.SUM, hj.F32, n, bf, b1), 4, 30),
    "scan incl u32 (8 B/elem)": (lambda n: dev.prefix_sum(hj.U32, n, True, bu, bo), 8, 30),
    "compress p=0.5 (3 B/elem)": (lambda n: dev.compress(n, bc, b50, bo), 3, 30),
    "compress p=0.01 (1.04 B/elem)": (lambda n: dev.compress(n, bc, b01, bo), 1.04, 30),
    "histogram 2^16 bins (4 B/key)": (lambda n: dev.scatter_reduce(hj.SUM, hj.U32, n, bk, None, 1, bh, 1 << 16), 4, 30),
}


def isolated(fn, iters=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3


def train(fn, reps=20, iters=5):
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    return float(np.median(ts)) * 1e3


print(f"PDL {'off (HJ_NO_PDL)' if os.environ.get('HJ_NO_PDL') else 'on'}; sizes 2^{lo}..2^{hi}; times in us")
for name, (fn, bpe, cap) in ops.items():
    ns = [1 << k for k in range(lo, min(hi, cap) + 1)]
    iso = [isolated(lambda: fn(n)) for n in ns]
    trn = [train(lambda: fn(n)) for n in ns]
    for label, t in (("isolated", iso), ("train", trn)):
        A = np.vstack([np.ones(len(ns)), np.array(ns, dtype=np.float64)]).T
        (a, b), *_ = np.linalg.lstsq(A, np.array(t), rcond=None)
        bw = bpe / b / 1e3 if b > 0 else float("nan")  # bytes/us -> GB/s: bpe[B] / b[us/elem] = MB/s... (1e-3 GB/s per B/us)
        cells = "  ".join(f"2^{int(np.log2(n))}:{x:8.1f}" for n, x in zip(ns, t))
        eff27 = [bpe * n / x / 1e3 for n, x in zip(ns, t) if n == (1 << 27)]
        print(f"{name:32s} {label:8s} fixed a = {a:6.1f} us  asymptotic {bw:7.0f} GB/s  at 2^27: {eff27[0] if eff27 else float('nan'):7.0f} GB/s | {cells}")
