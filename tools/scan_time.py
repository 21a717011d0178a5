import importlib, os, sys, torch
ROOT = os.environ.get("HJ_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
torch.cuda.set_device(0); dev = hj.Device.cuda(0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side); dev.set_stream(side.cuda_stream)
wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
g = torch.Generator(device="cuda").manual_seed(0)
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = 1 << log2n
xu = torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.int32)
xf = torch.rand(n, device="cuda", generator=g, dtype=torch.float32)
out = torch.empty(n, device="cuda", dtype=torch.int32)
bu, bf, bo = wrap(xu), wrap(xf), wrap(out)
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev: a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts)//2], ts[0]
for rep in range(2):
    for name, ty, b in (("u32 incl", hj.U32, bu), ("f32 incl", hj.F32, bf)):
        med, best = t(lambda: dev.prefix_sum(ty, n, True, b, bo))
        print(f"{name} 2^{log2n}: median {8*n/med/1e6:.0f} GB/s best {8*n/best/1e6:.0f} GB/s")
