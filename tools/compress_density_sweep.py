#!/usr/bin/env python3
"""compress over a sweep of mask densities (2^28 bytes by default): time and GB/s of algorithmic bytes (n + 4 count).
HJ_COMPRESS_STAGED=lo,mid moves the thresholds between the sparse / lane-major staged / quad staged paths
(1,3000 = lane-major everywhere, 1,1 = quads everywhere).  Usage: [log2 n]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
torch.cuda.set_device(0)
dev = hj.Device.cuda(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); dev.set_stream(st.cuda_stream)
wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
g = torch.Generator(device="cuda").manual_seed(0)
idx = torch.zeros(n, device="cuda", dtype=torch.int32); cnt = torch.zeros(4, device="cuda", dtype=torch.int32)
print(f"HJ_COMPRESS_STAGED={os.environ.get('HJ_COMPRESS_STAGED', '(default 176,960)')} n=2^{int(np.log2(n))}")
for p in (0.01, 0.05, 0.1, 0.15, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.78, 0.85, 0.99):
    m = (torch.rand(n, device="cuda", generator=g) < p).to(torch.uint8)
    want = torch.nonzero(m).flatten().to(torch.int32)
    fn = lambda: dev.compress(n, wrap(cnt), wrap(m), wrap(idx))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    c = int(cnt[0].item())
    ok = c == want.numel() and bool((idx[:c] == want).all())
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(15)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    us = float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3
    print(f"  p={p:4.2f}  {us:8.1f} us  {(n + 4 * c) / us / 1e3:7.0f} GB/s  exact={ok}")
    del m, want
