#!/usr/bin/env python3
"""Traced `mask.compress()` (zero-fill kernel + Compress pass) at 2^28 with and without the index
prefill elision of execute_graph (HJ_NO_PREFILL_ELISION=1)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200"); tr = importlib.import_module("hephaestus-jit_b200.tr")
torch.cuda.set_device(0); dev = hj.Device.cuda(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); dev.set_stream(st.cuda_stream)
n = 1 << 28
g = torch.Generator(device="cuda").manual_seed(0)
u = torch.rand(n, device="cuda", generator=g, dtype=torch.float32)
x = tr.from_buffer(dev.wrap(u.data_ptr(), 4 * n), hj.F32, n)
for p in (0.5, 0.01, 0.99):
    for fused_mask in (True, False):
        mask = x.lt(tr.literal(p, hj.F32))
        if not fused_mask:
            mask.schedule(); tr.compile().launch(dev)
        count, index = mask.compress()
        graph = tr.compile()
        for _ in range(4): graph.launch(dev)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in ev: a.record(); graph.launch(dev); b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)[5]
        c = int(count.to_vec(np.uint32)[0])
        print(f"p={p} mask {'computed in the fill kernel' if fused_mask else 'already evaluated'}: {ms*1e3:8.1f} us  count={c}", flush=True)
        del count, index, mask, graph
