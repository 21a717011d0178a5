#!/usr/bin/env python3
"""Development check of the scatter-reduce histogram paths against torch.bincount, with timing."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
torch.cuda.set_device(0)
dev = hj.Device.cuda(0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side); dev.set_stream(side.cuda_stream)
wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
g = torch.Generator(device="cuda").manual_seed(0)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[iters // 2]
for n, nb, dist in ((1 << 28, 1 << 16, "uniform"), ((1 << 28) - 777, 1 << 16, "uniform"), (1 << 28, 1 << 16, "zipf"), (1 << 28, 1 << 10, "uniform"),
                    (1 << 28, 100000, "uniform"), ((1 << 20) + 3, 1 << 16, "uniform"), (1 << 26, 1 << 18, "uniform"), (1 << 26, 40000, "oob"),
                    (1 << 28, 1 << 10, "one"), (1 << 28, 256, "uniform"), (1 << 28, 4000, "uniform"), (1 << 28, 1 << 10, "zipf"),
                    (1 << 28, 1 << 18, "uniform"), (1 << 28, 1 << 20, "uniform"), (1 << 28, 1 << 22, "uniform"), (1 << 28, 1 << 18, "zipf")):
    if dist == "zipf":
        u = torch.rand(n, device="cuda", generator=g)
        keys = ((nb ** u - 1).clamp(0, nb - 1)).to(torch.int32)  # heavy head, log-uniform
    elif dist == "one":
        keys = torch.full((n,), 7, device="cuda", dtype=torch.int32)
    elif dist == "oob":
        keys = torch.randint(0, 2 * nb, (n,), device="cuda", generator=g, dtype=torch.int32)
    else:
        keys = torch.randint(0, nb, (n,), device="cuda", generator=g, dtype=torch.int32)
    hist = torch.full((nb,), 5, device="cuda", dtype=torch.int32)
    dev.scatter_reduce(hj.SUM, hj.U32, n, wrap(keys), None, 1, wrap(hist), nb)
    torch.cuda.synchronize()
    k64 = keys.to(torch.int64)
    want = torch.bincount(k64[k64 < nb], minlength=nb)[:nb] + 5
    ok = bool(torch.equal(hist.to(torch.int64), want))
    ms = timeit(lambda: dev.scatter_reduce(hj.SUM, hj.U32, n, wrap(keys), None, 1, wrap(hist), nb))
    print(f"n={n} bins={nb} {dist}: ok={ok} {ms:.3f} ms {4*n/ms/1e6:.1f} GB/s {n/ms/1e6:.1f} Gkeys/s", flush=True)
    del keys, hist, k64, want
