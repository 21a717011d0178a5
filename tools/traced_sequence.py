#!/usr/bin/env python3
"""tr.array -> launch -> to_vec of the C2 chain over pinned host arrays: blocking upload vs tr.array_async
(chunk-wise upload | kernels | download), next to hj_kernel_map_host.  Usage: python tools/traced_sequence.py [log2 n]"""
import ctypes
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
tr = importlib.import_module("hephaestus-jit_b200.tr")
irm = importlib.import_module("hephaestus-jit_b200.ir")

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
dev = hj.Device.cuda(0)
hx = torch.rand(n, dtype=torch.float32).pin_memory().numpy()
hy = torch.empty(n, dtype=torch.float32).pin_memory().numpy()
kern = dev.kernel(irm.c2_chain_ir())


def traced(make):
    xv = make(hx, dev)
    tv = xv.fma(tr.literal(1.5, hj.F32), tr.literal(0.25, hj.F32))
    yv = tv.sin().select(xv.gt(tr.literal(0.0, hj.F32)), tv.exp2())
    yv.schedule()
    tr.compile().launch(dev)
    yv.to_vec(out=hy.view(np.uint8))


def bench(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 8 * n / float(np.median(ts)) / 1e9


hu = torch.randint(0, 4, (n,), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
hs = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)


def traced_scan(make):
    sv = make(hu, dev).prefix_sum(True)
    sv.schedule()
    tr.compile().launch(dev)
    sv.to_vec(out=hs.view(np.uint8))


want = None
for name, fn in (("scan: tr.array       -> launch -> to_vec", lambda: traced_scan(tr.array)),
                 ("scan: tr.array_async -> launch -> to_vec", lambda: traced_scan(tr.array_async)),
                 ("scan: hj_prefix_sum_host", lambda: dev.prefix_sum_host(hj.U32, n, True, hu.ctypes.data, hs.ctypes.data))):
    hs[:] = 0
    gbs = bench(fn)
    if want is None:
        want = hs.copy()
    print(f"{name:40s} {gbs:7.1f} GB/s (8 B/elem, n = 2^{int(np.log2(n))})  identical to the blocking result: {bool(np.array_equal(hs, want))}")

want = None
for name, fn in (("tr.array       -> launch -> to_vec", lambda: traced(tr.array)),
                 ("tr.array_async -> launch -> to_vec", lambda: traced(tr.array_async)),
                 ("hj_kernel_map_host (chunks 2^23)", lambda: dev.map_host(kern, n, [hx.ctypes.data, hy.ctypes.data], 1 << 23)),
                 ("hj_kernel_map_host (chunks 2^24)", lambda: dev.map_host(kern, n, [hx.ctypes.data, hy.ctypes.data], 1 << 24))):
    hy[:] = 0
    gbs = bench(fn)
    if want is None:
        want = hy.copy()
    same = bool(np.array_equal(hy.view(np.uint32), want.view(np.uint32)))
    print(f"{name:40s} {gbs:7.1f} GB/s (8 B/elem, n = 2^{int(np.log2(n))})  identical to the blocking result: {same}")
