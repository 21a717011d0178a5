import importlib, os, sys, torch
ROOT = os.environ.get("HJ_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
torch.cuda.set_device(0); dev = hj.Device.cuda(0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side); dev.set_stream(side.cuda_stream)
wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
g = torch.Generator(device="cuda").manual_seed(0); n = 1 << 28
for p in (0.5, 0.01, 0.99):
    m = (torch.rand(n, device="cuda", generator=g) < p).to(torch.uint8)
    idx = torch.zeros(n, device="cuda", dtype=torch.int32); cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
    fn = lambda: dev.compress(n, wrap(cnt), wrap(m), wrap(idx))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev: a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[5]
    print(f"cfg={os.environ.get('HJ_COMPRESS_CFG')} p={p}: {ms*1e3:.1f} us  ({(n + 4*p*n)/ms/1e6:.0f} GB/s-equivalent)")
