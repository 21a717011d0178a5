"""Sharded execution on the GPU, parity against the oracle on the WHOLE array.

Two layers are covered at world 1, 2, 3 and 4 (the device ops also at 8):
  * the sharded device ops of include/hj.h with their exchange FUSED into the kernel (reduce, scan
    with a deferred seed, compress with the counts exchange, the packed-16 histogram fold+exchange);
  * whole traced programs over arrays partitioned with ``tr.array_sharded`` — Graph.launch ->
    hj_execute_graph_sharded, pass for pass what backend/vulkan/mod.rs:151-383 does on one device.

The communicator is bootstrapped WITHOUT NCCL (hj_comm_create_local / hj_comm_connect: the CUDA-IPC
handles of the peer mailboxes travel over a torch.distributed gloo group), so the ranks of a world
may share one GPU: on the one-GPU box the driver tests on, the kernels of the two or three
processes (up to eight) are time-sliced and the exchange code is exactly the multi-GPU one.  With enough GPUs
every rank takes its own (tests/test_sharded_gpu.py covers the NCCL bootstrap there).
"""
import importlib
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hj = importlib.import_module("hephaestus-jit_b200")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, body, args, errq):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ.setdefault("OMP_NUM_THREADS", "2")
    import torch.distributed as dist
    comm = None
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        hjw = importlib.import_module("hephaestus-jit_b200")
        sh = importlib.import_module("hephaestus-jit_b200.sharded")
        dev = hjw.Device.cuda(rank % max(hjw.device_count(), 1))
        comm = sh.Comm.local_from_torch(dev)
        info = comm.info()
        assert info["world"] == world and info["rank"] == rank and info["nccl"] == 0
        assert info["peer_memory"] == (1 if world > 1 else 0)
        globals()[body](rank, world, dev, comm, *args)
        dev.sync()
        dist.barrier()
    except BaseException:  # noqa: BLE001 - reported to the parent, which fails the test
        errq.put((rank, traceback.format_exc()))
    finally:
        try:
            if comm is not None:
                comm.destroy()
            dist.destroy_process_group()
        except Exception:
            pass


def _run(world, body, *args, timeout=600):
    ctx = mp.get_context("spawn")
    errq = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, body, args, errq)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
    hung = [p for p in procs if p.is_alive()]
    for p in hung:
        p.kill()
    errors = []
    while not errq.empty():
        errors.append(errq.get())
    assert not hung, f"{len(hung)} rank(s) did not finish within {timeout} s"
    assert not errors, "\n".join(f"--- rank {r}\n{tb}" for r, tb in errors)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


# ---- bodies (run in every rank's process) ------------------------------------------------------------------
def _device_ops(rank, world, dev, comm, n):
    import oracle
    sh = importlib.import_module("hephaestus-jit_b200.sharded")
    rng = np.random.Generator(np.random.PCG64(7))     # every rank generates the whole array
    u = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    f = rng.random(n, dtype=np.float32)
    s, e = sh.shard_bounds(n, world, rank)
    nl = e - s
    out1 = dev.create_buffer(8)
    for rep in range(3):   # repeated exchanges alternate the mailbox parities
        for op, ty, arr in ((hj.SUM, hj.U32, u), (hj.MAX, hj.U32, u), (hj.XOR, hj.U32, u), (hj.MIN, hj.F32, f)):
            comm.reduce(op, ty, nl, dev.create_buffer_from_slice(arr[s:e]), out1)
            assert out1.to_host(arr.dtype, 0, 1)[0] == oracle.reduce(op, ty, arr)[0], (op, ty)
    comm.reduce(hj.SUM, hj.F32, nl, dev.create_buffer_from_slice(f[s:e]), out1)
    exact = float(f.astype(np.float64).sum())
    assert abs(float(out1.to_host(np.float32, 0, 1)[0]) - exact) <= 1e-5 * exact
    # scan, materialised (12 B/elem) and with a deferred seed (8 B/elem, the exchange inside the scan kernel)
    dst, seed = dev.create_buffer(4 * nl), dev.create_buffer(8)
    src = dev.create_buffer_from_slice(u[s:e])
    for inclusive in (True, False):
        want = oracle.prefix_sum(oracle.U32, u, inclusive)[s:e]
        comm.prefix_sum(hj.U32, nl, inclusive, src, dst)
        assert np.array_equal(dst.to_host(np.uint32), want)
        comm.prefix_sum_deferred(hj.U32, nl, inclusive, src, dst, seed)
        local = dst.to_host(np.uint32)
        off = seed.to_host(np.uint32, 0, 1)[0]
        with np.errstate(over="ignore"):
            assert off == np.uint32(u[:s].astype(np.uint64).sum() & 0xFFFFFFFF)
            assert np.array_equal(local + off, want)
        L = importlib.import_module("hephaestus-jit_b200._lib")
        L.check(L.lib.hj_apply_seed(dev.handle, hj.U32, nl, dst.handle, seed.handle))
        assert np.array_equal(dst.to_host(np.uint32), want)
    # u64 / u8 deferred scans: two-word exchange payloads, narrow element types
    u64 = rng.integers(0, 2**63, size=n, dtype=np.uint64)
    d64 = dev.create_buffer(8 * nl)
    comm.prefix_sum_deferred(hj.U64, nl, True, dev.create_buffer_from_slice(u64[s:e]), d64, seed)
    with np.errstate(over="ignore"):
        assert np.array_equal(d64.to_host(np.uint64) + seed.to_host(np.uint64, 0, 1)[0], np.cumsum(u64, dtype=np.uint64)[s:e])
    # compress: per-rank segment with global indices, counts exchanged inside the kernel
    for p in (0.4, 0.01, 0.97):
        mask = (rng.random(n) < p).astype(np.uint8)
        idx = dev.create_buffer_from_slice(np.full(nl, 0xDEADBEEF, np.uint32))
        cnt, counts = dev.create_buffer(4), dev.create_buffer(4 * world)
        comm.compress(nl, s, dev.create_buffer_from_slice(mask[s:e]), idx, cnt, counts)
        gcnt, gidx = oracle.compress(mask)
        c = counts.to_host(np.uint32)
        off = int(c[:rank].sum())
        assert int(cnt.to_host(np.uint32)[0]) == gcnt and int(c.sum()) == gcnt
        got = idx.to_host(np.uint32)
        assert np.array_equal(got[: int(c[rank])], gidx[off: off + int(c[rank])])
        assert (got[int(c[rank]):] == 0xDEADBEEF).all()      # entries beyond the count are not touched
    # histogram 2^16 bins: packed-16 ring + fold and exchange in one kernel; dst's own contents are kept
    keys = rng.integers(0, 1 << 16, size=n).astype(np.uint32)
    start = rng.integers(0, 100, size=1 << 16).astype(np.uint32)
    for rep in range(2):
        hist = dev.create_buffer_from_slice(start if rank == 0 else np.zeros(1 << 16, np.uint32))
        comm.scatter_reduce(hj.SUM, hj.U32, nl, dev.create_buffer_from_slice(keys[s:e]), None, 1, hist, 1 << 16)
        assert np.array_equal(hist.to_host(np.uint32), oracle.histogram_u32_mt(keys, 1 << 16) + start)
    # a skewed histogram (every key equal: the sweep path) and an odd bin count
    same = np.full(n, 12345, np.uint32)
    hist = dev.create_buffer_from_slice(np.zeros(50001, np.uint32))
    comm.scatter_reduce(hj.SUM, hj.U32, nl, dev.create_buffer_from_slice(same[s:e]), None, 1, hist, 50001)
    want = np.zeros(50001, np.uint32)
    want[12345] = n
    assert np.array_equal(hist.to_host(np.uint32), want)


def _traced_program(rank, world, dev, comm, n):
    import oracle
    tr = importlib.import_module("hephaestus-jit_b200.tr")
    sh = importlib.import_module("hephaestus-jit_b200.sharded")
    rng = np.random.Generator(np.random.PCG64(21))
    x = (rng.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
    u = rng.integers(0, 1 << 20, size=n).astype(np.uint32)
    keys = rng.integers(0, 1 << 16, size=n).astype(np.uint32)
    table = rng.integers(0, 1 << 30, size=4096).astype(np.uint32)
    s, e = sh.shard_bounds(n, world, rank)

    vx, vu, vk = tr.array_sharded(x, comm), tr.array_sharded(u, comm), tr.array_sharded(keys, comm)
    vt = tr.array(table, dev)                                   # a replica: the gather table
    assert vx.shard() == (s, e - s, False) and vt.shard() is None
    # C2 chain + reduction of its result
    t = vx.fma(tr.literal(1.5, hj.F32), tr.literal(0.25, hj.F32))
    y = t.sin().select(vx.gt(tr.literal(0.0, hj.F32)), t.exp2())
    y.schedule()
    total = y.abs().reduce_max()
    total.schedule()
    # scan with consumers: the fused kernel behind it adds the deferred seed at load; Index is global
    scan = vu.prefix_sum(True)
    z = scan.add(tr.sized_index(n)).add(vt.gather(vu.and_(tr.literal(4095, hj.U32))))
    z.schedule()
    scan.schedule()
    scan2 = vk.prefix_sum(False)    # never scheduled itself: lives and dies inside the graph
    ssum = scan2.reduce_max()       # a device op on a deferred result: materialised first
    ssum.schedule()
    # compress of a traced mask, histogram-shaped scatter-reduce
    mask = vu.and_(tr.literal(3, hj.U32)).eq(tr.literal(1, hj.U32))
    count, index = mask.compress()
    hist = tr.sized_literal(0, 1 << 16, hj.U32)
    tr.sized_literal(1, n, hj.U32).scatter_reduce(hist, vk, hj.SUM)
    hist.schedule()
    g = tr.compile()
    report = g.launch(dev, timed=True)
    names = [p[0] for p in report.passes]
    if world > 1:
        assert any("deferred seed" in nm for nm in names), names
    assert any(nm.startswith("Histogram") for nm in names) or n < (1 << 20), names

    want_y = oracle.c2_chain(x)
    assert y.shard() == (s, e - s, False)
    assert np.allclose(y.to_vec(), want_y[s:e], rtol=4e-7, atol=1e-7)
    assert float(total.item()) == float(np.abs(want_y).max()) or np.isclose(float(total.item()), float(np.abs(want_y).max()), rtol=4e-7)
    want_scan = np.cumsum(u, dtype=np.uint32)
    assert np.array_equal(scan.to_vec(np.uint32), want_scan[s:e])
    with np.errstate(over="ignore"):
        want_scan2 = np.cumsum(keys, dtype=np.uint32) - keys
    assert int(ssum.item()) == int(want_scan2.max())
    assert scan.shard() == (s, e - s, world > 1)      # still (local scan, offset) on more than one rank
    with np.errstate(over="ignore"):
        want_z = want_scan + np.arange(n, dtype=np.uint32) + table[u & 4095]
    assert np.array_equal(z.to_vec(np.uint32), want_z[s:e])
    m = ((u & 3) == 1).astype(np.uint8)
    gcnt, gidx = oracle.compress(m)
    assert int(count.to_vec(np.uint32)[0]) == gcnt
    lcnt = int(m[s:e].sum())
    off = int(m[:s].sum())
    got = index.to_vec(np.uint32)
    assert index.shard() == (s, e - s, False) and len(got) == e - s
    assert np.array_equal(got[:lcnt], gidx[off: off + lcnt]) and (got[lcnt:] == 0).all()   # zero tail like the reference
    assert np.array_equal(hist.to_vec(np.uint32), oracle.histogram_u32_mt(keys, 1 << 16))

    # a second launch consumes results of the first: a deferred scan result as an input of a later graph
    w = scan.add(tr.literal(7, hj.U32))
    w.schedule()
    tr.compile().launch(dev)
    assert np.array_equal(w.to_vec(np.uint32), want_scan[s:e] + np.uint32(7))
    scan.materialise()
    assert scan.shard() == (s, e - s, False)
    assert np.array_equal(scan.to_vec(np.uint32), want_scan[s:e])


def _unsupported_shapes(rank, world, dev, comm, n):
    tr = importlib.import_module("hephaestus-jit_b200.tr")
    u = np.arange(n, dtype=np.uint32)
    vu = tr.array_sharded(u, comm)
    perm = tr.array((n - 1 - u).astype(np.uint32), dev)
    # a sharded array read through a computed index: replicas only
    bad = vu.gather(perm)
    bad.schedule()
    with pytest.raises(hj.HjError, match="computed index|replica"):
        tr.compile().launch(dev)
    # a plain scatter into a replica from a kernel over sharded data
    dst = tr.sized_literal(0, 64, hj.U32)
    dst.schedule()
    tr.compile().launch(dev)
    vu.scatter(dst, vu.and_(tr.literal(63, hj.U32)))
    with pytest.raises(hj.HjError, match="replica"):
        tr.compile().launch(dev)


def _cached_sharded_graph(rank, world, dev, comm, n):
    """The relaunch path: the same sharded pass list launched again and again is captured into ONE CUDA
    graph and replayed (hj_execute_graph_sharded_cached) — possible because the exchange epochs of the
    sharded kernels live in device memory.  Inputs change between launches; every result is checked."""
    import oracle
    irm = importlib.import_module("hephaestus-jit_b200.ir")
    L = importlib.import_module("hephaestus-jit_b200._lib")
    sh = importlib.import_module("hephaestus-jit_b200.sharded")
    s, e = sh.shard_bounds(n, world, rank)
    nl = e - s
    bx, by = dev.create_buffer(4 * nl), dev.create_buffer(4 * nl)
    bf, bs = dev.create_buffer(4 * nl), dev.create_buffer(4)
    bu, bscan, bseed = dev.create_buffer(4 * nl), dev.create_buffer(4 * nl), dev.create_buffer(16)
    bm, bidx, bcnt = dev.create_buffer(nl), dev.create_buffer(4 * nl), dev.create_buffer(4)
    passes = [{"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": irm.c2_chain_ir(), "size": n},
              {"kind": hj.PASS_REDUCE, "arg": hj.MAX, "resources": [3, 2]},
              {"kind": hj.PASS_PREFIX_SUM, "arg": 1, "resources": [5, 4]},
              {"kind": hj.PASS_COMPRESS, "resources": [7, 8, 6]}]
    descs = [(n, hj.F32, 4), (n, hj.F32, 4), (n, hj.U32, 4), (1, hj.U32, 4), (n, hj.U32, 4), (n, hj.U32, 4), (n, hj.BOOL, 1),
             (n, hj.U32, 4), (1, hj.U32, 4)]
    S, R = L.RES_SHARDED, L.RES_REPLICATED
    g = hj.PreparedGraph(dev, passes, [bx, by, bf, bs, bu, bscan, bm, bidx, bcnt], descs, comm, [S, S, S, R, S, S, S, S, R],
                         [None] * 5 + [bseed] + [None] * 3, graph_key=0xC0FFEE + n)
    hows = []
    for it in range(5):
        rng = np.random.Generator(np.random.PCG64(100 + it))       # every rank generates the whole arrays
        x = (rng.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
        f = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        u = rng.integers(0, 1 << 12, size=n).astype(np.uint32)
        m = (rng.random(n) < 0.3 + 0.1 * it).astype(np.uint8)
        for b, a in ((bx, x), (bf, f), (bu, u), (bm, m)):
            b.upload(a[s:e])
        bidx.fill_zero()
        g.run()
        hows.append(g.how.value)
        assert np.allclose(by.to_host(np.float32), oracle.c2_chain(x)[s:e], rtol=4e-7, atol=1e-7), it
        assert bs.to_host(np.uint32)[0] == f.max(), it
        deferred = g.deferred()[5]
        assert deferred == (world > 1)
        off = bseed.to_host(np.uint32, 0, 1)[0] if deferred else np.uint32(0)
        with np.errstate(over="ignore"):
            assert np.array_equal(bscan.to_host(np.uint32) + off, np.cumsum(u, dtype=np.uint32)[s:e]), it
        gcnt, gidx = oracle.compress(m)
        lc, before = int(m[s:e].sum()), int(m[:s].sum())
        assert int(bcnt.to_host(np.uint32)[0]) == gcnt, it
        assert np.array_equal(bidx.to_host(np.uint32)[:lc], gidx[before: before + lc]), it
    assert hows == [0, 1, 2, 2, 2], hows       # executed, captured, replayed ...


def _random_sharded_programs(rank, world, dev, comm, n, n_programs, seed):
    """Random traced programs over sharded arrays against their numpy statement on the WHOLE arrays: every
    rank builds the same program (same seed) over its blocks; elementwise chains, comparisons and selects,
    scans (whose deferred seed the next fused kernel must add), reductions broadcast back, gathers from a
    replicated table, the global Index, casts, recorded loops, compress — in random order and nesting."""
    import random
    tr = importlib.import_module("hephaestus-jit_b200.tr")
    sh = importlib.import_module("hephaestus-jit_b200.sharded")
    U32, F32 = hj.U32, hj.F32
    s, e = sh.shard_bounds(n, world, rank)
    rnd = random.Random(seed)
    data_rng = np.random.Generator(np.random.PCG64(seed))
    k_ = np.arange(n, dtype=np.uint32)
    table_np = data_rng.integers(0, 1 << 20, size=1024).astype(np.uint32)
    for prog in range(n_programs):
        table = tr.array(table_np, dev)
        a0 = data_rng.integers(0, 1 << 12, size=n).astype(np.uint32)
        b0 = data_rng.integers(0, 1 << 31, size=n).astype(np.uint32)
        pool = [(tr.array_sharded(a0, comm), a0), (tr.array_sharded(b0, comm), b0), (tr.sized_index(n), k_.copy())]
        bools = [(pool[0][0].lt(tr.literal(2048, U32)), a0 < 2048)]
        terminal, trace_ops = [], []
        with np.errstate(over="ignore"):
            for _ in range(rnd.randrange(3, 12)):
                op = rnd.randrange(11)
                (va, a), (vb, b) = rnd.choice(pool), rnd.choice(pool)
                trace_ops.append(op)
                if op == 0:
                    which = rnd.randrange(5)
                    pool.append([(va.add(vb), a + b), (va.sub(vb), a - b), (va.min(vb), np.minimum(a, b)),
                                 (va.or_(vb), a | b), (va.xor(vb), a ^ b)][which])
                elif op == 1:
                    c = rnd.randrange(1, 9)
                    pool.append((va.mul(tr.literal(c, U32)), a * np.uint32(c)))
                elif op == 2:
                    bools.append((va.lt(vb), a < b))
                elif op == 3:
                    vm, m = rnd.choice(bools)
                    pool.append((va.select(vm, vb), np.where(m, a, b)))
                elif op == 4:   # scan: the result stays (local scan, offset) until somebody needs the values
                    inc = rnd.random() < 0.5
                    cs = np.cumsum(a, dtype=np.uint32)
                    pool.append((va.prefix_sum(inc), cs if inc else cs - a))
                elif op == 5:   # reduction broadcast back: a replica read at a computed (literal) index
                    pool.append((vb.add(va.reduce_sum().gather(tr.literal(0, U32))), b + a.sum(dtype=np.uint32)))
                elif op == 6:   # gather from the replicated table at a computed index
                    pool.append((table.gather(va.and_(tr.literal(1023, U32))), table_np[a & np.uint32(1023)]))
                elif op == 7:   # compress: per-rank segment + global count
                    vm, m = rnd.choice(bools)
                    cnt, idx = vm.compress()
                    terminal.append(("compress", cnt, idx, m))
                elif op == 8:   # recorded loop: x = 3x + 1, `reps` times
                    reps = rnd.randrange(1, 4)

                    def body(c, vs, reps=reps):
                        x, it = vs
                        it = it.add(tr.literal(1, U32))
                        return it.lt(tr.literal(reps, U32)), [x.mul(tr.literal(3, U32)).add(tr.literal(1, U32)), it]

                    _, (x1, _it) = tr.loop_record(tr.literal(True), [va, tr.sized_literal(0, n, U32)], body)
                    w = a.copy()
                    for _ in range(reps):
                        w = w * np.uint32(3) + np.uint32(1)
                    pool.append((x1, w))
                elif op == 9:   # through f32 and back, exact for small integers
                    small_v, small = va.and_(tr.literal(1023, U32)), a & np.uint32(1023)
                    pool.append((small_v.cast(F32).fma(tr.literal(2.0, F32), tr.literal(1.0, F32)).cast(U32),
                                 small * np.uint32(2) + np.uint32(1)))
                else:           # min / max reductions of a (possibly deferred) value
                    terminal.append(("reduce", va.reduce_max(), None, a.max()))
        # every other program also takes a wavefront step: compact one of its masks, update the selected
        # elements of a sharded array in place through the compacted indices (a DynSize kernel per segment)
        wave = random.Random(seed * 7919 + prog)
        if wave.random() < 0.6:
            vm, m = wave.choice(bools)
            sel = np.flatnonzero(m).astype(np.uint32)
            d0 = data_rng.integers(0, 1 << 16, size=n).astype(np.uint32)
            dst = tr.array_sharded(d0, comm)
            idx = vm.compress_dyn()
            with np.errstate(over="ignore"):
                variant = wave.randrange(3)
                if variant == 0:    # dst[i] = 3 dst[i] + table[i & 1023]
                    val = dst.gather(idx).mul(tr.literal(3, U32)).add(table.gather(idx.and_(tr.literal(1023, U32))))
                    want_d = d0.copy()
                    want_d[sel] = d0[sel] * np.uint32(3) + table_np[sel & np.uint32(1023)]
                elif variant == 1:  # dst[i] = a0[i] ^ b0[i] + i: two other sharded arrays and the index value itself
                    val = pool[0][0].gather(idx).xor(pool[1][0].gather(idx)).add(idx)
                    want_d = d0.copy()
                    want_d[sel] = (a0[sel] ^ b0[sel]) + sel
                else:               # dst[i] = min(dst[i], a0[i])
                    val = dst.gather(idx).min(pool[0][0].gather(idx))
                    want_d = d0.copy()
                    want_d[sel] = np.minimum(d0[sel], a0[sel])
            val.scatter(dst, idx)
            terminal.append(("scatter", dst, None, want_d))
            del idx, val
        outs = [pool[-1], rnd.choice(pool[3:] or pool)]
        for v, _ in outs:
            v.schedule()
        for t in terminal:
            t[1].schedule()
        g = tr.compile()
        g.launch(dev)
        ctx = f"program {prog} ops {trace_ops} world {world} rank {rank}"
        for v, want in outs:
            sh_info = v.shard()
            got = v.to_vec(np.uint32)
            assert np.array_equal(got, want[s:e] if sh_info is not None else want), ctx
        for t in terminal:
            if t[0] == "reduce":
                assert int(t[1].item(np.uint32)) == int(t[3]), ctx
            elif t[0] == "scatter":
                assert np.array_equal(t[1].to_vec(np.uint32), t[3][s:e]), ctx + " (wavefront step)"
            else:
                _, cnt, idx, m = t
                sel = np.flatnonzero(m).astype(np.uint32)
                assert int(cnt.to_vec(np.uint32)[0]) == sel.size, ctx
                lc, before = int(m[s:e].sum()), int(m[:s].sum())
                got = idx.to_vec(np.uint32)
                assert np.array_equal(got[:lc], sel[before: before + lc]) and (got[lc:] == 0).all(), ctx
        del pool, bools, terminal, outs, g, table


@pytest.mark.parametrize("world,n", [(1, (1 << 17) + 5), (2, (1 << 18) + 11), (3, 50_001)])
def test_random_programs_over_sharded_arrays(world, n):
    _run(world, "_random_sharded_programs", n, 16, 1234 + world)


def _recorded_sharded_function(rank, world, dev, comm, n):
    """record(f) over sharded inputs (record.rs:93-210): traced once, then relaunched with new blocks; the
    relaunches go through hj_execute_graph_sharded_cached.  The function scans, feeds the (deferred) scan
    into a fused kernel, reduces and compresses."""
    tr = importlib.import_module("hephaestus-jit_b200.tr")
    rec = importlib.import_module("hephaestus-jit_b200.record")
    sh = importlib.import_module("hephaestus-jit_b200.sharded")
    U32 = hj.U32
    s, e = sh.shard_bounds(n, world, rank)
    calls = []

    def fn(x):
        calls.append(1)
        y = x.mul(tr.literal(3, U32)).add(tr.literal(1, U32))
        scan = y.prefix_sum(True)
        z = scan.xor(tr.sized_index(n))
        top = z.reduce_max()
        cnt, idx = x.and_(tr.literal(7, U32)).eq(tr.literal(0, U32)).compress()
        return [z, top, cnt, idx]

    f = rec.record(fn)
    c0, r0, _ = dev.graph_cache_stats()
    for it in range(12):
        x = np.random.Generator(np.random.PCG64(50 + it)).integers(0, 1 << 10, size=n).astype(np.uint32)
        (z, top, cnt, idx), _ = f(dev, tr.array_sharded(x, comm))
        with np.errstate(over="ignore"):
            want_z = np.cumsum(x * np.uint32(3) + np.uint32(1), dtype=np.uint32) ^ np.arange(n, dtype=np.uint32)
        assert np.array_equal(z.to_vec(np.uint32), want_z[s:e]), it
        assert int(top.item(np.uint32)) == int(want_z.max()), it
        m = (x & 7) == 0
        sel = np.flatnonzero(m).astype(np.uint32)
        assert int(cnt.to_vec(np.uint32)[0]) == sel.size, it
        lc, before = int(m[s:e].sum()), int(m[:s].sum())
        got = idx.to_vec(np.uint32)
        assert np.array_equal(got[:lc], sel[before: before + lc]) and (got[lc:] == 0).all(), it
        del z, top, cnt, idx
    assert len(calls) == 1, "the function must be traced once"
    c1, r1, p1 = dev.graph_cache_stats()
    print(f"recorded sharded function: captured {c1 - c0}, replayed {r1 - r0} of 12 launches")
    # the pool hands a relaunch the same addresses again (possibly after a cycle of a few launches): some
    # launches must have been replays of a captured CUDA graph
    assert c1 - c0 >= 1 and r1 - r0 >= 1, (c0, r0, c1, r1)


def _wavefront_reference(a0, steps, index_as_value=False):
    """The wavefront loop on the whole array: lanes alive in `mask` are scaled and stay alive while > 0.1
    (index_as_value: they take their position in the compacted sequence instead)."""
    a, alive = a0.copy(), np.ones(a0.size, bool)
    history = []
    for _ in range(steps):
        idx = np.flatnonzero(alive).astype(np.uint32)
        v = np.arange(idx.size, dtype=np.float32) if index_as_value else a[idx] * np.float32(0.9)
        alive[idx] = v > np.float32(0.1)
        a[idx] = v
        history.append((idx, a.copy(), alive.copy()))
    return history


def _wavefront_pass_list(rank, world, dev, comm, n):
    """Compress + the two DynSize kernels behind it (tests/test_sharded_cpu.py: _wavefront_passes) through
    hj_execute_graph_sharded(_cached): every rank runs the kernels over its own segment, sized by its own
    count; the relaunches replay one captured CUDA graph.  Checked against the loop on the whole array."""
    from test_sharded_cpu import _wavefront_passes
    sh = importlib.import_module("hephaestus-jit_b200.sharded")
    S, R = sh.RES_SHARDED, sh.RES_REPLICATED
    s, e = sh.shard_bounds(n, world, rank)
    nl = e - s
    a0 = np.random.Generator(np.random.PCG64(5)).random(n, dtype=np.float32)
    steps = 14
    history = _wavefront_reference(a0, steps)
    passes, descs = _wavefront_passes(n)
    ba, bm = dev.create_buffer_from_slice(a0[s:e]), dev.create_buffer_from_slice(np.ones(nl, np.uint8))
    bidx, bcnt, bseed = dev.create_buffer_from_slice(np.zeros(nl, np.uint32)), dev.create_buffer(4), dev.create_buffer(16)
    g = hj.PreparedGraph(dev, passes, [ba, bm, bidx, bcnt], descs, comm=comm, placement=[S, S, S, R],
                         seeds=[None, None, bseed, None], graph_key=0x77AF0)
    c0, r0, _ = dev.graph_cache_stats()
    for it, (idx, a, alive) in enumerate(history):
        g.run()
        lo, hi = np.searchsorted(idx, s), np.searchsorted(idx, e)
        assert int(bcnt.to_host(np.uint32, 0, 1)[0]) == idx.size, it
        assert int(bseed.to_host(np.uint32, 0, 1)[0]) == hi - lo, it
        assert np.array_equal(bidx.to_host(np.uint32)[: hi - lo], idx[lo:hi]), it
        assert np.array_equal(ba.to_host(np.float32), a[s:e]), it
        assert np.array_equal(bm.to_host(np.uint8).astype(bool), alive[s:e]), it
        assert g.segments() == [False, False, True, False]
    assert history[-1][0].size < history[0][0].size * 3 // 4, "the wavefront must have shrunk"
    c1, r1, _ = dev.graph_cache_stats()
    assert c1 - c0 == 1 and r1 - r0 == steps - 2, (c1 - c0, r1 - r0)
    # without a seed on the index segment there is nothing to size the kernels with: refused, not mis-sized
    g2 = hj.PreparedGraph(dev, passes, [ba, bm, bidx, bcnt], descs, comm=comm, placement=[S, S, S, R])
    with pytest.raises(hj.HjError, match="segment"):
        g2.run()
    # a conditional read of the segment (an inactive lane would address element 0 of another rank's block): refused
    bad, _ = _wavefront_passes(n, conditional=True)
    g3 = hj.PreparedGraph(dev, bad, [ba, bm, bidx, bcnt], descs, comm=comm, placement=[S, S, S, R], seeds=[None, None, bseed, None])
    with pytest.raises(hj.HjError, match="segment"):
        g3.run()
    # KernelOp::Index as a VALUE is the position in the global compacted sequence, as on one GPU
    ba2, bm2 = dev.create_buffer_from_slice(a0[s:e]), dev.create_buffer_from_slice(np.ones(nl, np.uint8))
    byidx, _ = _wavefront_passes(n, index_as_value=True)
    g5 = hj.PreparedGraph(dev, byidx, [ba2, bm2, bidx, bcnt], descs, comm=comm, placement=[S, S, S, R],
                          seeds=[None, None, bseed, None])
    for it, (idx, a, alive) in enumerate(_wavefront_reference(a0, 3, index_as_value=True)):
        g5.run()
        lo, hi = np.searchsorted(idx, s), np.searchsorted(idx, e)
        assert list(bseed.to_host(np.uint32, 0, 2)) == [hi - lo, lo], it
        assert np.array_equal(ba2.to_host(np.float32), a[s:e]), it
        assert np.array_equal(bm2.to_host(np.uint8).astype(bool), alive[s:e]), it
    # a device op over a segment: refused
    red = [{"kind": hj.PASS_COMPRESS, "resources": [2, 3, 1]}, {"kind": hj.PASS_REDUCE, "arg": hj.MAX, "resources": [3, 2]}]
    g4 = hj.PreparedGraph(dev, red, [ba, bm, bidx, bcnt], descs, comm=comm, placement=[S, S, S, R], seeds=[None, None, bseed, None])
    with pytest.raises(hj.HjError, match="compacted segment"):
        g4.run()


def _traced_wavefront(rank, world, dev, comm, n):
    """jit/test.rs:1020-1062 (`example`) over sharded arrays, traced: indices = mask.compress_dyn();
    b = a.gather(indices) * 0.9; (b > 0.1).scatter(mask, indices); b.scatter(a, indices) — recorded once,
    launched again and again; the compacted indices stay per-rank segments."""
    tr = importlib.import_module("hephaestus-jit_b200.tr")
    rec = importlib.import_module("hephaestus-jit_b200.record")
    sh = importlib.import_module("hephaestus-jit_b200.sharded")
    F32 = hj.F32
    s, e = sh.shard_bounds(n, world, rank)
    a0 = np.random.Generator(np.random.PCG64(6)).random(n, dtype=np.float32)
    steps = 10
    history = _wavefront_reference(a0, steps)
    a = tr.array_sharded(a0, comm)
    mask = tr.array_sharded(np.ones(n, np.bool_), comm)
    # traced and launched step by step: `indices` is a live DynSize variable, to_vec reads the rank's own count
    for it, (idx, want_a, alive) in enumerate(history[:5]):
        indices = mask.compress_dyn()
        b = a.gather(indices).mul(tr.literal(0.9, F32))
        b.gt(tr.literal(0.1, F32)).scatter(mask, indices)
        b.scatter(a, indices)
        a.schedule()
        indices.schedule()
        tr.compile().launch(dev)
        lo, hi = np.searchsorted(idx, s), np.searchsorted(idx, e)
        assert indices.is_segment() and indices.shard() == (s, hi - lo, False), (it, indices.shard())
        assert np.array_equal(indices.to_vec(np.uint32), idx[lo:hi]), it   # the rank's part of the compacted sequence
        assert np.array_equal(a.to_vec(np.float32), want_a[s:e]), it
        assert np.array_equal(mask.to_vec(np.uint8).astype(bool), alive[s:e]), it
        del indices, b
    # recorded once, relaunched (record.rs:93-210); outputs of a recorded function have a static extent
    # (graph.rs:345: Extent::Size(desc.size)): the index output is the block with its zero tail
    calls = []

    def step():
        calls.append(1)
        indices = mask.compress_dyn()
        b = a.gather(indices).mul(tr.literal(0.9, F32))
        b.gt(tr.literal(0.1, F32)).scatter(mask, indices)
        b.scatter(a, indices)
        a.schedule()
        return [indices]

    f = rec.record(step)
    for it, (idx, want_a, alive) in enumerate(history[5:]):
        (indices,), _ = f(dev)
        lo, hi = np.searchsorted(idx, s), np.searchsorted(idx, e)
        got = indices.to_vec(np.uint32)
        assert indices.is_segment() and len(got) == e - s, it
        assert np.array_equal(got[: hi - lo], idx[lo:hi]) and (got[hi - lo:] == 0).all(), it
        assert np.array_equal(a.to_vec(np.float32), want_a[s:e]), it
        assert np.array_equal(mask.to_vec(np.uint8).astype(bool), alive[s:e]), it
        del indices
    assert len(calls) == 1, "traced once"
    assert not a.is_segment() and a.shard() == (s, e - s, False)
    # KernelOp::Index as a value inside the DynSize kernel: the position in the GLOBAL compacted sequence,
    # as on one GPU (trace.rs:578-597 dynamic_index) — every surviving lane takes its rank in the wavefront
    alive = history[-1][2]
    sel = np.flatnonzero(alive).astype(np.uint32)
    count, index = mask.compress()
    j = tr.dynamic_index(n, count)
    j.cast(F32).scatter(a, index.gather(j))
    a.schedule()
    tr.compile().launch(dev)
    want = history[-1][1].copy()
    want[sel] = np.arange(sel.size, dtype=np.float32)
    assert int(count.to_vec(np.uint32)[0]) == sel.size
    assert np.array_equal(a.to_vec(np.float32), want[s:e])
    del count, index, j
    # a DynSize array written by one launch (aligned with the segment: the rank holds its own part of it) and
    # consumed, together with the indices, by a later launch
    lo, hi = np.searchsorted(sel, s), np.searchsorted(sel, e)
    indices = mask.compress_dyn()
    vals = a.gather(indices).add(tr.literal(1.0, F32))
    vals.schedule()
    indices.schedule()
    tr.compile().launch(dev)
    assert vals.is_segment() and indices.is_segment() and vals.shard() == (s, hi - lo, False)
    assert np.array_equal(vals.to_vec(np.float32), (want[sel] + np.float32(1.0))[lo:hi])
    vals.mul(tr.literal(2.0, F32)).scatter(a, indices)
    a.schedule()
    tr.compile().launch(dev)
    want[sel] = (want[sel] + np.float32(1.0)) * np.float32(2.0)
    assert np.array_equal(a.to_vec(np.float32), want[s:e])
    # device ops do not run over a segment: refused, not computed over the undefined tail
    bad = vals.reduce_max()
    bad.schedule()
    with pytest.raises(hj.HjError, match="compacted segment"):
        tr.compile().launch(dev)
    del bad, vals, indices


def _traced_nested_compaction(rank, world, dev, comm, n):
    """jit/test.rs:976-1019 (`dynamic_index`) over a sharded array: compact, gather, compact the gathered values
    again, gather through the second indices.  The second Compress is DynSize (its mask is aligned with the first
    segment); its indices are positions in the rank's part of the first sequence and address the rank's segments
    in place.  Every rank ends up with its part of the result, in order."""
    tr = importlib.import_module("hephaestus-jit_b200.tr")
    sh = importlib.import_module("hephaestus-jit_b200.sharded")
    I32 = hj.I32
    s, e = sh.shard_bounds(n, world, rank)
    lo, hi = 3, 7
    for seed in (13, 14):
        src = np.random.Generator(np.random.PCG64(seed)).integers(0, 10, size=n).astype(np.int32)
        src_var = tr.array_sharded(src, comm)
        indices = src_var.lt(tr.literal(hi, I32)).compress_dyn()
        values = src_var.gather(indices)
        indices2 = values.gt(tr.literal(lo, I32)).compress_dyn()
        values2 = values.gather(indices2)
        values2.schedule()
        indices2.schedule()
        tr.compile().launch(dev)
        mine = src[s:e]
        first = mine[mine < hi]                      # the rank's part of the first compacted sequence
        want = first[first > lo]
        assert values2.is_segment() and values2.shard() == (s, want.size, False)
        assert np.array_equal(values2.to_vec(np.int32), want)
        assert indices2.is_segment()
        assert np.array_equal(indices2.to_vec(np.uint32), np.flatnonzero(first > lo).astype(np.uint32))   # rank-local positions
        del src_var, indices, values, indices2, values2


@pytest.mark.parametrize("world", [1, 2, 3])
def test_nested_compaction_over_sharded_arrays(world):
    _run(world, "_traced_nested_compaction", (1 << 17) + 9)


@pytest.mark.parametrize("world", [1, 2, 3])
def test_wavefront_pass_list_runs_per_segment(world):
    _run(world, "_wavefront_pass_list", (1 << 18) + 33)


@pytest.mark.parametrize("world", [1, 2, 3])
def test_traced_wavefront_over_sharded_arrays(world):
    _run(world, "_traced_wavefront", (1 << 18) + 5)


@pytest.mark.parametrize("world", [1, 2])
def test_recorded_function_over_sharded_inputs(world):
    _run(world, "_recorded_sharded_function", (1 << 19) + 77)


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_pass_list_replays_as_one_cuda_graph(world):
    _run(world, "_cached_sharded_graph", (1 << 20) + 4099)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_sharded_device_ops_fused_exchange(world):
    _run(world, "_device_ops", (1 << 21) + 77)


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_traced_program_over_sharded_arrays(world):
    _run(world, "_traced_program", (1 << 21) + 13)


def test_small_shards_take_the_unfused_paths():
    _run(2, "_device_ops_small", 5003)


def _device_ops_small(rank, world, dev, comm, n):
    import oracle
    sh = importlib.import_module("hephaestus-jit_b200.sharded")
    rng = np.random.Generator(np.random.PCG64(3))
    u = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    mask = (rng.random(n) < 0.5).astype(np.uint8)
    s, e = sh.shard_bounds(n, world, rank)
    nl = e - s
    dst, seed = dev.create_buffer(4 * nl), dev.create_buffer(8)
    comm.prefix_sum_deferred(hj.U32, nl, True, dev.create_buffer_from_slice(u[s:e]), dst, seed)
    with np.errstate(over="ignore"):
        assert np.array_equal(dst.to_host(np.uint32) + seed.to_host(np.uint32, 0, 1)[0], oracle.prefix_sum(oracle.U32, u, True)[s:e])
    idx, cnt, counts = dev.create_buffer_from_slice(np.zeros(nl, np.uint32)), dev.create_buffer(4), dev.create_buffer(4 * world)
    comm.compress(nl, s, dev.create_buffer_from_slice(mask[s:e]), idx, cnt, counts)
    gcnt, gidx = oracle.compress(mask)
    c = counts.to_host(np.uint32)
    assert int(cnt.to_host(np.uint32)[0]) == gcnt
    assert np.array_equal(idx.to_host(np.uint32)[: int(c[rank])], gidx[int(c[:rank].sum()): int(c[:rank + 1].sum())])


def test_shapes_that_do_not_shard_fail_loudly():
    _run(2, "_unsupported_shapes", 1 << 16)
