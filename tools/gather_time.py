import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
torch.cuda.set_device(0); dev = hj.Device.cuda(0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side); dev.set_stream(side.cuda_stream)
wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
g = torch.Generator(device="cuda").manual_seed(0); n = (1 << 28) + 3
out = torch.empty(n, device="cuda", dtype=torch.float32)
for log_t in (20, 24, 28):
    table = torch.rand(1 << log_t, device="cuda", generator=g, dtype=torch.float32)
    idx = torch.randint(0, 1 << log_t, (n,), device="cuda", generator=g, dtype=torch.int32)
    fn = lambda: dev.gather(4, n, wrap(table), wrap(idx), wrap(out))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ok = bool(torch.equal(out, table[idx.long()]))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev: a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[5]
    print(f"gather table 2^{log_t}: ok={ok} {ms:.3f} ms {n/ms/1e6:.1f} Gelem/s {12*n/ms/1e6:.0f} GB/s nominal")
    del table, idx
