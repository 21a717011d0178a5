#!/usr/bin/env python3
"""Round-2 workload for compute-sanitizer (memcheck / synccheck): the kernels added or changed this round —
the packed-16 histogram with the fused fold kernel, the PDL-launched ring kernels in a captured graph, the
host-streamed ops (chunk carries, pinned count slots), apply_seed, the sharded pass interpreter with a
world-1 communicator (deferred-seed IR rewrite included), gather, and the traced program path."""
import ctypes, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
hj = importlib.import_module("hephaestus-jit_b200"); irm = importlib.import_module("hephaestus-jit_b200.ir")
L = importlib.import_module("hephaestus-jit_b200._lib"); sh = importlib.import_module("hephaestus-jit_b200.sharded")
tr = importlib.import_module("hephaestus-jit_b200.tr")
dev = hj.Device.cuda(0)
rng = np.random.Generator(np.random.PCG64(2))
n = (1 << 20) + 4097
u = rng.integers(0, 1 << 16, size=n).astype(np.uint32)
x = (rng.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
m = (rng.random(n) < 0.5).astype(np.uint8)
# histogram: packed-16 ring + fold kernel (PDL), odd bin count too
for nb in (1 << 16, 50001):
    keys = rng.integers(0, nb, size=n).astype(np.uint32); keys[: n // 4] = 11
    bh = dev.create_buffer_from_slice(np.ones(nb, np.uint32))
    dev.scatter_reduce(hj.SUM, hj.U32, n, dev.create_buffer_from_slice(keys), None, 1, bh, nb)
    assert np.array_equal(bh.to_host(np.uint32), np.bincount(keys, minlength=nb).astype(np.uint32) + 1)
# host-streamed ops with many small chunks
out = np.empty(n, np.uint32); idx = np.zeros(n, np.uint32)
dev.prefix_sum_host(hj.U32, n, False, u, out, 1 << 16); assert np.array_equal(out, np.cumsum(u, dtype=np.uint32) - u)
assert dev.reduce_host(hj.MAX, hj.U32, n, u, 1 << 16) == u.max()
c = dev.compress_host(n, m, idx, 0, 1 << 16); assert c == int(m.sum()) and np.array_equal(idx[:c], np.flatnonzero(m).astype(np.uint32))
y = np.empty(n, np.float32); dev.map_host(dev.kernel(irm.c2_chain_ir()), n, [x, y], 1 << 16)
# gather
tab = rng.random(1 << 16, dtype=np.float32); gi = rng.integers(0, 1 << 16, size=n).astype(np.uint32)
bo = dev.create_buffer(4 * n); dev.gather(4, n, dev.create_buffer_from_slice(tab), dev.create_buffer_from_slice(gi), bo)
assert np.array_equal(bo.to_host(np.float32), tab[gi])
# sharded pass interpreter, world 1, through the relaunch path (plain, captured, replayed)
comm = sh.Comm.local(dev, 0, 1, lambda h: [h])
vu, vx = tr.array_sharded(u, comm), tr.array_sharded(x, comm)
scan = vu.prefix_sum(True); z = scan.add(tr.sized_index(n)); z.schedule()
t = vx.fma(tr.literal(1.5, hj.F32), tr.literal(0.25, hj.F32)); yv = t.sin().select(vx.gt(tr.literal(0.0, hj.F32)), t.exp2()); yv.schedule()
tot = yv.abs().reduce_max(); tot.schedule()
cnt, ind = vu.and_(tr.literal(1, hj.U32)).eq(tr.literal(1, hj.U32)).compress()
g = tr.compile()
for _ in range(3):
    g.launch(dev)
assert np.array_equal(z.to_vec(np.uint32), np.cumsum(u, dtype=np.uint32) + np.arange(n, dtype=np.uint32))
assert int(cnt.to_vec(np.uint32)[0]) == int((u & 1).sum())
# apply_seed
seed = dev.create_buffer_from_slice(np.array([5], np.uint32)); bu = dev.create_buffer_from_slice(u)
L.check(L.lib.hj_apply_seed(dev.handle, hj.U32, n, bu.handle, seed.handle)); assert np.array_equal(bu.to_host(np.uint32), u + 5)
del vu, vx, scan, z, t, yv, tot, cnt, ind, g
comm.destroy(); dev.sync()
# asynchronous arrays: chunk-wise upload | kernels | download (map, integer scan, a consumer that has to wait)
os.environ.setdefault("HJ_ASYNC_CHUNK_ELEMS", "65536")
na = 1_000_003
ha = np.random.default_rng(5).integers(0, 1 << 20, na, dtype=np.uint32)
va = tr.array_async(ha, dev); ya = va.mul(tr.literal(3, hj.U32)); ya.schedule(); tr.compile().launch(dev)
assert np.array_equal(ya.to_vec(), ha * 3)
vs = tr.array_async(ha, dev).prefix_sum(True); vs.schedule(); tr.compile().launch(dev)
assert np.array_equal(vs.to_vec(), np.cumsum(ha, dtype=np.uint32))
vm = tr.array_async(ha, dev).reduce_max(); vm.schedule(); tr.compile().launch(dev)
assert int(vm.item()) == int(ha.max())
del va, ya, vs, vm
dev.sync()
print("sanitize r02 workload ok")
