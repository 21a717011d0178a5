#!/usr/bin/env python3
"""Quick per-kernel HBM throughput probe (development tool; bench.py is the judged harness).

Times each hand-written kernel with CUDA events on the stream it is launched on, inputs
resident in HBM and larger than L2, and prints GB/s of ALGORITHMIC bytes next to the measured
copy peak from MEASURED_PEAKS.json."""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")


def peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"]
    except Exception:
        return 6650.0


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    n = 1 << args.log2n
    torch.cuda.set_device(0)
    dev = hj.Device.cuda(0)
    # a non-default torch stream: events recorded by torch and kernels enqueued by the library
    # must share ONE stream (the legacy default stream handle is 0, which hj treats as "own")
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    dev.set_stream(side.cuda_stream)
    pk = peak()
    rows = []

    def wrap(t):
        return dev.wrap(t.data_ptr(), t.numel() * t.element_size())

    g = torch.Generator(device="cuda").manual_seed(0)
    xf = torch.rand(n, device="cuda", generator=g, dtype=torch.float32)
    xu = torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.int32)
    out1 = torch.zeros(16, device="cuda", dtype=torch.float32)
    outn = torch.empty(n, device="cuda", dtype=torch.int32)
    bf, bu, bo1, bon = wrap(xf), wrap(xu), wrap(out1), wrap(outn)

    # torch copy as the local roofline reference
    ms, best = timeit(lambda: outn.copy_(xu), args.iters)
    rows.append(("torch copy_ (8 B/elem)", 8 * n, ms, best))

    ms, best = timeit(lambda: dev.reduce(hj.SUM, hj.F32, n, bf, bo1), args.iters)
    rows.append(("reduce sum f32 (4 B/elem)", 4 * n, ms, best))
    ms, best = timeit(lambda: dev.reduce(hj.SUM, hj.U32, n, bu, bo1), args.iters)
    rows.append(("reduce sum u32 (4 B/elem)", 4 * n, ms, best))
    ms, best = timeit(lambda: dev.reduce(hj.MAX, hj.F32, n, bf, bo1), args.iters)
    rows.append(("reduce max f32 (4 B/elem)", 4 * n, ms, best))
    ms, best = timeit(lambda: dev.prefix_sum(hj.U32, n, True, bu, bon), args.iters)
    rows.append(("scan incl u32 (8 B/elem)", 8 * n, ms, best))
    ms, best = timeit(lambda: dev.prefix_sum(hj.U32, n, False, bu, bon), args.iters)
    rows.append(("scan excl u32 (8 B/elem)", 8 * n, ms, best))
    ms, best = timeit(lambda: dev.prefix_sum(hj.F32, n, True, bf, wrap(outn)), args.iters)
    rows.append(("scan incl f32 (8 B/elem)", 8 * n, ms, best))
    for p in (0.5, 0.01, 0.99):
        mask = (torch.rand(n, device="cuda", generator=g) < p).to(torch.uint8)
        cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
        bm, bc = wrap(mask), wrap(cnt)
        ms, best = timeit(lambda: dev.compress(n, bc, bm, bon), args.iters)
        c = int(cnt.item())
        rows.append((f"compress p={p} ((1+4p) B/elem)", n + 4 * c, ms, best))
    nk = min(n, 1 << 28)
    for nb in (1 << 10, 1 << 16):
        keys = torch.randint(0, nb, (nk,), device="cuda", generator=g, dtype=torch.int32)
        hist = torch.zeros(nb, device="cuda", dtype=torch.int32)
        bk, bh = wrap(keys), wrap(hist)
        ms, best = timeit(lambda: dev.scatter_reduce(hj.SUM, hj.U32, nk, bk, None, 1, bh, nb), args.iters)
        rows.append((f"histogram {nb} bins (4 B/key)", 4 * nk, ms, best))
    print(f"n = 2^{args.log2n}; measured copy peak {pk:.0f} GB/s")
    for name, nbytes, ms, best in rows:
        gbs = nbytes / ms / 1e6
        print(f"{name:38s} median {ms:8.3f} ms  {gbs:8.1f} GB/s  {gbs / pk * 100:6.1f}% of peak   (best {nbytes / best / 1e6:8.1f} GB/s)")


if __name__ == "__main__":
    main()
