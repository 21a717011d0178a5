#!/usr/bin/env python3
"""The host ceiling of an N-rank end-to-end run: every rank streams 1 GiB pinned H2D and 1 GiB D2H at
once (what bench.py's e2e does per chunk), all ranks concurrently; aggregate GB/s, with the pinned
buffers placed by hj_host_alloc (wherever the rank runs) and by hj_host_alloc_near (the GPU's NUMA node).
Launch: python -m torch.distributed.run --nproc-per-node N tools/pcie_probe_ranks.py"""
import ctypes, importlib, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200"); L = importlib.import_module("hephaestus-jit_b200._lib")
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("gloo")
torch.cuda.set_device(lr); dev = hj.Device.cuda(lr)
n = 1 << 30
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
def numa_of(ptr):
    try:
        import ctypes.util
        libc = ctypes.CDLL(None, use_errno=True)
        status = ctypes.c_int(-1); page = ctypes.c_void_p(ptr)
        rc = libc.syscall(279, 0, 1, ctypes.byref(page), None, ctypes.byref(status), 0)  # move_pages: query
        return status.value if rc == 0 else None
    except Exception:
        return None
for label, near in (("hj_host_alloc", False), ("hj_host_alloc_near", True)):
    hin, hout = ctypes.c_void_p(), ctypes.c_void_p()
    for p in (hin, hout):
        L.check(L.lib.hj_host_alloc_near(dev.handle, n, ctypes.byref(p)) if near else L.lib.hj_host_alloc(n, ctypes.byref(p)))
    res = {}
    for name, up, down in (("H2D", True, False), ("D2H", False, True), ("both", True, True)):
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize()
            if world > 1: dist.barrier()
            t0 = time.perf_counter()
            step = n // 16
            for c in range(16):
                if up: cudart.cudaMemcpyAsync(d_in.data_ptr() + c * step, hin.value + c * step, step, 1, s1.cuda_stream)
                if down: cudart.cudaMemcpyAsync(hout.value + c * step, d_out.data_ptr() + c * step, step, 2, s2.cuda_stream)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
            best = min(best, dt)
        res[name] = world * n * (up + down) / best / 1e9
    node = numa_of(hin.value)
    if rank == 0:
        print(f"{world} rank(s), {label:20s}: aggregate H2D {res['H2D']:7.1f}  D2H {res['D2H']:7.1f}  both {res['both']:7.1f} GB/s   (rank 0 buffer on NUMA node {node})", flush=True)
    L.lib.hj_host_free(hin); L.lib.hj_host_free(hout)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
