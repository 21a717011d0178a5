#!/usr/bin/env python3
"""e2e (pinned host -> hj_kernel_map_host -> pinned host) over chunk sizes, C2 chain, 2^28 f32."""
import ctypes, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200"); irm = importlib.import_module("hephaestus-jit_b200.ir")
L = importlib.import_module("hephaestus-jit_b200._lib")
dev = hj.Device.cuda(0)
n = 1 << 28
kernel = dev.kernel(irm.c2_chain_ir())
hx, hy = ctypes.c_void_p(), ctypes.c_void_p()
L.check(L.lib.hj_host_alloc(4 * n, ctypes.byref(hx))); L.check(L.lib.hj_host_alloc(4 * n, ctypes.byref(hy)))
host_x = np.ctypeslib.as_array(ctypes.cast(hx, ctypes.POINTER(ctypes.c_float)), shape=(n,))
host_y = np.ctypeslib.as_array(ctypes.cast(hy, ctypes.POINTER(ctypes.c_float)), shape=(n,))
host_x[:] = np.random.Generator(np.random.PCG64(0)).random(n, dtype=np.float32) * 8 - 4
for sh in [int(a) for a in sys.argv[1:]] or [20, 21, 22, 23, 24]:
    host_y[:] = 0
    dev.map_host(kernel, n, [hx.value, hy.value], 1 << sh)
    def want(x):  # numpy statement of the chain (a sanity check of the plumbing, not the parity test)
        t = x * np.float32(1.5) + np.float32(0.25)
        return np.where(x > 0, np.sin(t), np.exp2(t)).astype(np.float32)
    ok = all(np.allclose(host_y[o:o + 4096], want(host_x[o:o + 4096].copy()), rtol=1e-5, atol=1e-6)
             for o in (0, n // 3, n // 2 + 12345, n - 4096))
    t0 = time.perf_counter()
    for _ in range(5): dev.map_host(kernel, n, [hx.value, hy.value], 1 << sh)
    dev.sync()
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk 2^{sh}: {dt * 1e3:7.2f} ms  {8 * n / dt / 1e9:6.1f} GB/s  ok={ok}", flush=True)
