#!/usr/bin/env python3
"""Throughput over sizes 2^10 .. 2^30 — the sweep of the reference's criterion benches
(benches/vulkan.rs:157-192: compress_large and prefix_sum_large_u32 over n = 2^10 .. 2^29), plus
reduce.  CUDA events on the launch stream, median of 20; elements/s like the reference's
`Throughput::Elements`, and GB/s of algorithmic bytes."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
torch.cuda.set_device(0); dev = hj.Device.cuda(0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side); dev.set_stream(side.cuda_stream)
wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
g = torch.Generator(device="cuda").manual_seed(0)
N = 1 << 30
xu = torch.ones(N, device="cuda", dtype=torch.int32)
mask = torch.ones(N, device="cuda", dtype=torch.uint8)   # all true, like the reference bench
out = torch.empty(N, device="cuda", dtype=torch.int32)
cnt = torch.zeros(4, device="cuda", dtype=torch.int32)
bu, bm, bo, bc = wrap(xu), wrap(mask), wrap(out), wrap(cnt)
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev: a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[iters // 2]
print(f"{'log2 n':>6} | {'scan u32 us':>11} {'Gelem/s':>8} {'GB/s':>7} | {'compress us':>11} {'Gelem/s':>8} {'GB/s':>7} | {'reduce us':>9} {'Gelem/s':>8} {'GB/s':>7}")
for k in range(10, 31, 2):
    n = 1 << k
    ms_s = t(lambda: dev.prefix_sum(hj.U32, n, False, bu, bo))
    assert int(out[n - 1].item()) == n - 1   # benches/vulkan.rs:137 (true exclusive scan here)
    ms_c = t(lambda: dev.compress(n, bc, bm, bo))
    assert int(cnt[0].item()) == n           # benches/vulkan.rs:114
    ms_r = t(lambda: dev.reduce(hj.SUM, hj.U32, n, bu, bc))
    assert int(cnt[0].item()) == n
    print(f"{k:6d} | {ms_s*1e3:11.1f} {n/ms_s/1e6:8.2f} {8*n/ms_s/1e6:7.0f} | {ms_c*1e3:11.1f} {n/ms_c/1e6:8.2f} {5*n/ms_c/1e6:7.0f} | {ms_r*1e3:9.1f} {n/ms_r/1e6:8.2f} {4*n/ms_r/1e6:7.0f}")
