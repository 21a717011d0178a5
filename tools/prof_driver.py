#!/usr/bin/env python3
"""Runs one hand-written kernel a few times on resident inputs (target for `ncu`)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")


def main():
    which = sys.argv[1]
    log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    n = 1 << log2n
    torch.cuda.set_device(0)
    dev = hj.Device.cuda(0)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    dev.set_stream(side.cuda_stream)
    wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
    g = torch.Generator(device="cuda").manual_seed(0)
    if which in ("scan", "scan_excl"):
        x = torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.int32)
        y = torch.empty_like(x)
        fn = lambda: dev.prefix_sum(hj.U32, n, which == "scan", wrap(x), wrap(y))
    elif which == "reduce":
        x = torch.rand(n, device="cuda", generator=g)
        y = torch.zeros(4, device="cuda")
        fn = lambda: dev.reduce(hj.SUM, hj.F32, n, wrap(x), wrap(y))
    elif which.startswith("compress"):
        p = float(which.split(":")[1]) if ":" in which else 0.5
        m = (torch.rand(n, device="cuda", generator=g) < p).to(torch.uint8)
        idx = torch.zeros(n, device="cuda", dtype=torch.int32)
        cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
        fn = lambda: dev.compress(n, wrap(cnt), wrap(m), wrap(idx))
    elif which.startswith("hist"):
        nb = int(which.split(":")[1]) if ":" in which else 1 << 16
        k = torch.randint(0, nb, (n,), device="cuda", generator=g, dtype=torch.int32)
        h = torch.zeros(nb, device="cuda", dtype=torch.int32)
        fn = lambda: dev.scatter_reduce(hj.SUM, hj.U32, n, wrap(k), None, 1, wrap(h), nb)
    elif which.startswith("gather"):
        log_t = int(which.split(":")[1]) if ":" in which else 28
        table = torch.rand(1 << log_t, device="cuda", generator=g)
        idx = torch.randint(0, 1 << log_t, (n,), device="cuda", generator=g, dtype=torch.int32)
        out = torch.empty(n, device="cuda")
        fn = lambda: dev.gather(4, n, wrap(table), wrap(idx), wrap(out))
    elif which == "map":
        irm = importlib.import_module("hephaestus-jit_b200.ir")
        x = torch.rand(n, device="cuda", generator=g) * 8 - 4
        y = torch.empty_like(x)
        k = dev.kernel(irm.c2_chain_ir())
        fn = lambda: dev.launch(k, n, [wrap(x), wrap(y)])
    elif which == "wavefront":   # Compress + the two DynSize kernels of the wavefront step, pass by pass
        irm = importlib.import_module("hephaestus-jit_b200.ir")
        a = torch.rand(n, device="cuda", generator=g)
        m = (torch.rand(n, device="cuda", generator=g) < 0.5).to(torch.uint8)
        idx = torch.zeros(n, device="cuda", dtype=torch.int32)
        cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
        passes, descs = irm.wavefront_step_passes(n, threshold=-1.0)
        graph = hj.PreparedGraph(dev, passes, [wrap(a), wrap(m), wrap(idx), wrap(cnt)], descs)
        fn = graph.run
    else:
        raise SystemExit(f"unknown kernel {which}")
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
