#!/usr/bin/env python3
"""Small end-to-end run of every hand-written kernel for compute-sanitizer (memcheck / racecheck /
synccheck): sizes chosen so that every kernel takes its ring / TMA path with a ragged last tile."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200")
dev = hj.Device.cuda(0)
rng = np.random.Generator(np.random.PCG64(0))
n = (1 << 21) + 4097
u = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
bu, bo, b1 = dev.create_buffer_from_slice(u), dev.create_buffer(4 * n), dev.create_buffer(8)
dev.reduce(hj.SUM, hj.U32, n, bu, b1)
assert b1.to_host(np.uint32)[0] == np.uint32(u.sum(dtype=np.uint64) & 0xFFFFFFFF)
for incl in (True, False):
    dev.prefix_sum(hj.U32, n, incl, bu, bo)
    want = np.cumsum(u, dtype=np.uint64)
    want = (want if incl else want - u).astype(np.uint32)
    assert np.array_equal(bo.to_host(np.uint32), want)
d = rng.random(n // 2)
bd, bdo = dev.create_buffer_from_slice(d), dev.create_buffer(8 * (n // 2))
dev.prefix_sum(hj.F64, n // 2, True, bd, bdo)
assert np.allclose(bdo.to_host(np.float64), np.cumsum(d), rtol=1e-9)
for p in (0.5, 0.02, 1.0):
    m = (rng.random(n) < p).astype(np.uint8)
    bm, bc = dev.create_buffer_from_slice(m), dev.create_buffer_from_slice(np.zeros(1, np.uint32))
    bo.fill_zero()
    dev.compress(n, bc, bm, bo)
    want = np.nonzero(m)[0].astype(np.uint32)
    c = int(bc.to_host(np.uint32)[0])
    assert c == want.size and np.array_equal(bo.to_host(np.uint32)[:c], want)
for nb, lit in ((1 << 16, 1), (1 << 16, 2), (1 << 10, 1), (100000, 1)):
    keys = rng.integers(0, nb, size=n).astype(np.uint32)
    keys[: n // 3] = 7  # hot bin: exercises the 0x8000 cashing of the packed counters
    bh = dev.create_buffer_from_slice(np.zeros(nb, np.uint32))
    dev.scatter_reduce(hj.SUM, hj.U32, n, dev.create_buffer_from_slice(keys), None, lit, bh, nb)
    assert np.array_equal(bh.to_host(np.uint32), (np.bincount(keys, minlength=nb) * lit).astype(np.uint32))
dev.sync()
print("sanitize workload ok")
