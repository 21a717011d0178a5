#!/usr/bin/env python3
"""Per-tile timeline of the scan kernel (development aid): where does a tile spend its life?"""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
from importlib import import_module
lib = import_module("hephaestus-jit_b200._lib").lib

n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
tile = 16384
n_tiles = (n + tile - 1) // tile
torch.cuda.set_device(0)
dev = hj.Device.cuda(0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side); dev.set_stream(side.cuda_stream)
x = torch.randint(0, 4, (n,), device="cuda", dtype=torch.int32)
y = torch.empty_like(x)
wrap = lambda t: dev.wrap(t.data_ptr(), t.numel() * t.element_size())
for _ in range(3):
    dev.prefix_sum(hj.U32, n, True, wrap(x), wrap(y))
tr = torch.zeros(n_tiles * 5, device="cuda", dtype=torch.int64)
lib.hj_debug_scan_trace.argtypes = [ctypes.c_void_p]
lib.hj_debug_scan_trace(ctypes.c_void_p(tr.data_ptr()))
dev.prefix_sum(hj.U32, n, True, wrap(x), wrap(y))
torch.cuda.synchronize()
lib.hj_debug_scan_trace(None)
t = tr.cpu().numpy().reshape(n_tiles, 5).astype(np.int64)
t0 = t[:, 0].min()
start, loaded, prefix, end, smid = (t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0, t[:, 3] - t0, t[:, 4])
print(f"tiles {n_tiles}, kernel span {end.max() / 1e3:.1f} us")
for name, a in (("load (start->local scan done)", loaded - start), ("chain (-> prefix known)", prefix - loaded),
                ("store issue (-> end)", end - prefix), ("lifetime", end - start)):
    print(f"{name:34s} mean {a.mean():8.0f} ns  p50 {np.median(a):8.0f}  p90 {np.percentile(a, 90):8.0f}  max {a.max():8.0f}")
# concurrency: how many tiles are in each phase at sample times
ts = np.linspace(end.max() * 0.3, end.max() * 0.7, 50)
for nm, a, b in (("loading", start, loaded), ("chain-wait", loaded, prefix), ("storing", prefix, end)):
    c = [(np.sum((a <= q) & (b > q))) for q in ts]
    print(f"avg tiles in phase {nm:10s}: {np.mean(c):7.1f}")
# order: is start time monotone in ticket? how far ahead do tiles start vs finish
print("start-time inversions (tile i+1 started before tile i):", int(np.sum(np.diff(start) < 0)))
k = n_tiles // 2
print("sample rows around the middle (tile, start, loaded, prefix, end, smid) in us:")
for i in range(k, k + 12):
    print(i, *(f"{v / 1e3:8.2f}" for v in (start[i], loaded[i], prefix[i], end[i])), smid[i])
np.save(os.path.join(ROOT, "gpurun_out", "scan_trace.npy"), t)
