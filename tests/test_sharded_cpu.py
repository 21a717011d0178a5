"""Host logic of the multi-GPU layer on CPU: shard partitioning, and the exchange protocol
(partials -> all-gather -> rank-order fold; totals -> exclusive offsets; counts -> global offsets)
run over torch.distributed `gloo` with world_size 2 and 3, the local work done by the oracle."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sharded = importlib.import_module("hephaestus-jit_b200.sharded")


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 1000, (1 << 30) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [sharded.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_exclusive_offsets():
    assert sharded.exclusive_offsets([3, 0, 5, 2]) == [0, 3, 3, 8]
    off = sharded.exclusive_offsets([np.uint32(0xFFFFFFFF), np.uint32(2), np.uint32(1)])
    with np.errstate(over="ignore"):
        assert [int(x) for x in off] == [0, 0xFFFFFFFF, 1]  # wraps like the device arithmetic


def test_rebalance_plan_moves_every_element_once():
    rng = np.random.Generator(np.random.PCG64(3))
    for world in (1, 2, 3, 4, 8):
        for counts in ([0] * world, [5] * world, [100] + [0] * (world - 1), [0] * (world - 1) + [17],
                       list(rng.integers(0, 1000, size=world))):
            total = int(sum(counts))
            sent = 0
            inbox = {r: [] for r in range(world)}
            for r in range(world):
                sends, recvs = sharded.rebalance_plan(counts, r)
                # my sends tile my segment, my receives tile my block, both in order
                assert [o for _, o, _ in sends] == list(np.cumsum([0] + [n for _, _, n in sends[:-1]])) if sends else True
                assert sum(n for _, _, n in sends) == counts[r]
                lo, hi = sharded.shard_bounds(total, world, r)
                assert sum(n for _, _, n in recvs) == hi - lo
                sent += sum(n for _, _, n in sends)
                for peer, _, n in sends:
                    inbox[peer].append((r, n))
            assert sent == total
            for r in range(world):  # what the peers send to r is what r expects to receive
                _, recvs = sharded.rebalance_plan(counts, r)
                assert inbox[r] == [(p, n) for p, _, n in recvs]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    import oracle
    sh = importlib.import_module("hephaestus-jit_b200.sharded")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.Generator(np.random.PCG64(42))  # every rank generates the whole array
        u = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        f = rng.random(n, dtype=np.float32)
        mask = (rng.random(n) < 0.3).astype(np.uint8)
        s, e = sh.shard_bounds(n, world, rank)

        # reduce: local partial (oracle) -> exchange -> equals the single-array reduction
        with np.errstate(over="ignore"):
            tot = sh.exchange_reduce(oracle.reduce(oracle.SUM, oracle.U32, u[s:e])[0], lambda a, b: a + b)
        assert tot == oracle.reduce(oracle.SUM, oracle.U32, u)[0]
        mx = sh.exchange_reduce(oracle.reduce(oracle.MAX, oracle.F32, f[s:e])[0], max)
        assert mx == f.max()
        x = sh.exchange_reduce(oracle.reduce(oracle.XOR, oracle.U32, u[s:e])[0], lambda a, b: a ^ b)
        assert x == np.bitwise_xor.reduce(u)
        fs = sh.exchange_reduce(oracle.reduce(oracle.SUM, oracle.F32, f[s:e])[0], lambda a, b: np.float32(a + b))
        assert abs(float(fs) - float(f.astype(np.float64).sum())) <= 1e-5 * n

        # scan: shard total -> exclusive offset -> seeded local scan == slice of the global scan
        with np.errstate(over="ignore"):
            off = sh.exchange_scan_offset(oracle.reduce(oracle.SUM, oracle.U32, u[s:e])[0])
            local = oracle.prefix_sum(oracle.U32, u[s:e], True) + np.uint32(off)
        assert np.array_equal(local, oracle.prefix_sum(oracle.U32, u, True)[s:e])

        # compress: local compaction with global indices; counts -> offsets; concatenation == global
        cnt, idx = oracle.compress(mask[s:e], index_base=s)
        counts, offsets, total = sh.exchange_counts(cnt)
        gcnt, gidx = oracle.compress(mask)
        assert total == gcnt and counts[rank] == cnt
        assert np.array_equal(idx[:cnt], gidx[offsets[rank]: offsets[rank] + cnt])

        # rebalance: a skewed compaction (most survivors on the first ranks) re-partitioned evenly
        skew = (rng.random(n) < np.linspace(0.9, 0.02, n)).astype(np.uint8)
        cnt, idx = oracle.compress(skew[s:e], index_base=s)
        gcnt, gidx = oracle.compress(skew)
        mine = sh.exchange_rebalance(idx[:cnt])
        lo, hi = sh.shard_bounds(gcnt, world, rank)
        assert np.array_equal(mine, gidx[lo:hi])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_protocol_gloo(world):
    mp.spawn(_worker, args=(world, _free_port(), 100003), nprocs=world, join=True)


# ---- placement planning of the sharded pass interpreter (hj_shard_plan, host only) ------------------------
def _c2_like_passes(n, bins=1 << 10):
    """kernel y = f(x); reduce; scan; compress (with its zero-fill kernel); histogram-shaped kernel."""
    hj = importlib.import_module("hephaestus-jit_b200")
    irm = importlib.import_module("hephaestus-jit_b200.ir")
    # resources: 0 x, 1 y, 2 sum, 3 scan, 4 mask, 5 index, 6 count, 7 keys, 8 hist, 9 table, 10 gathered
    fill = irm.IRBuilder()
    u32 = fill.scalar(hj.U32)
    fill.scatter(fill.buffer_ref(u32), fill.literal(hj.U32, 0), fill.index())
    hist = irm.IRBuilder()
    u32 = hist.scalar(hj.U32)
    dst, keys = hist.buffer_ref(u32), hist.buffer_ref(u32)
    hist.scatter_reduce(hj.SUM, dst, hist.literal(hj.U32, 1), hist.gather(u32, keys, hist.index()))
    gat = irm.IRBuilder()   # out[i] = table[x[i]]: a replica read through a computed index
    u32 = gat.scalar(hj.U32)
    x, table, idx = gat.buffer_ref(u32), gat.buffer_ref(u32), gat.index()
    gat.scatter(gat.buffer_ref(u32), gat.gather(u32, table, gat.gather(u32, x, idx)), idx)
    passes = [
        {"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": irm.c2_chain_ir(), "size": n},
        {"kind": hj.PASS_REDUCE, "arg": hj.SUM, "resources": [2, 1]},
        {"kind": hj.PASS_PREFIX_SUM, "arg": 1, "resources": [3, 7]},
        {"kind": hj.PASS_KERNEL, "resources": [5], "ir": fill, "size": n},
        {"kind": hj.PASS_COMPRESS, "resources": [5, 6, 4]},
        {"kind": hj.PASS_KERNEL, "resources": [8, 7], "ir": hist, "size": n},
        {"kind": hj.PASS_KERNEL, "resources": [7, 9, 10], "ir": gat, "size": n},
    ]
    descs = [(n, hj.F32, 4), (n, hj.F32, 4), (1, hj.F32, 4), (n, hj.U32, 4), (n, hj.BOOL, 1), (n, hj.U32, 4),
             (1, hj.U32, 4), (n, hj.U32, 4), (bins, hj.U32, 4), (4096, hj.U32, 4), (n, hj.U32, 4)]
    return passes, descs


def test_shard_plan_propagates_placement():
    S, R, A = sharded.RES_SHARDED, sharded.RES_REPLICATED, sharded.RES_AUTO
    passes, descs = _c2_like_passes(1 << 20)
    #        x  y  sum scan mask index count keys hist table gathered
    given = [S, A, A,  A,   S,   A,    A,    S,   R,   R,    A]
    got = sharded.shard_plan(passes, descs, given)
    assert got == [S, S, R, S, S, S, R, S, R, R, S]
    # nothing sharded going in: everything is a replica, the single-GPU program
    assert sharded.shard_plan(passes, descs, [A] * len(descs)) == [R] * len(descs)
    # (index, resource 5: its zero-fill kernel has no sharded input of its own — the demand flows
    # backwards from the Compress pass whose mask is sharded)


def _wavefront_passes(n, index_as_value=False, conditional=False):
    """Compress + the two DynSize kernels of the wavefront step (hephaestus-jit_b200/ir.py)."""
    irm = importlib.import_module("hephaestus-jit_b200.ir")
    return irm.wavefront_step_passes(n, index_as_value=index_as_value, conditional=conditional)


def test_shard_plan_places_the_wavefront_step():
    """DynSize kernels behind a sharded Compress run per segment: what they reach through the compacted indices
    stays sharded, the count is a replica.  Nothing sharded going in -> the single-GPU program."""
    S, R, A = sharded.RES_SHARDED, sharded.RES_REPLICATED, sharded.RES_AUTO
    passes, descs = _wavefront_passes(1 << 16)
    assert sharded.shard_plan(passes, descs, [A, S, A, A]) == [S, S, S, R]
    assert sharded.shard_plan(passes, descs, [S, S, A, A]) == [S, S, S, R]
    assert sharded.shard_plan(passes, descs, [A, A, A, A]) == [R, R, R, R]
    # KernelOp::Index as a value (the position in the global compacted sequence) is fine
    passes, descs = _wavefront_passes(1 << 16, index_as_value=True)
    assert sharded.shard_plan(passes, descs, [A, S, A, A]) == [S, S, S, R]
    # a conditional read of the segment: not a segment kernel, `a` stays undecided (a replica), and the
    # launch then refuses the pass instead of computing something else
    passes, descs = _wavefront_passes(1 << 16, conditional=True)
    assert sharded.shard_plan(passes, descs, [A, S, A, A]) == [R, S, S, R]


def test_shard_plan_places_a_nested_compaction():
    """jit/test.rs:976-1019 (`dynamic_index`) as a pass list: compact, gather, compact the gathered values again
    (a DynSize Compress whose mask is aligned with the first segment), gather through the second indices.
    Everything that lives per compacted sequence is per rank; the two counts are replicas."""
    hj = importlib.import_module("hephaestus-jit_b200")
    irm = importlib.import_module("hephaestus-jit_b200.ir")
    S, R, A = sharded.RES_SHARDED, sharded.RES_REPLICATED, sharded.RES_AUTO
    n = 1 << 14

    def through(write_mask):   # out[i] = src[idx[i]]  (or out[i] = src[idx[i]] > 3)
        b = irm.IRBuilder()
        i32, u32, bl = b.scalar(hj.I32), b.scalar(hj.U32), b.scalar(hj.BOOL)
        rsrc, ridx = b.buffer_ref(i32), b.buffer_ref(u32)
        i = b.index()
        t = b.literal(hj.BOOL, 1)
        v = b.gather(i32, rsrc, b.gather(u32, ridx, i, t), t)
        if write_mask:
            b.scatter(b.buffer_ref(bl), b.bop(irm.BOP_GT, bl, v, b.literal(hj.I32, 3)), i)
        else:
            b.scatter(b.buffer_ref(i32), v, i)
        return b

    # resources: 0 count1, 1 src, 2 idx1, 3 mask1, 4 mask2, 5 count2, 6 idx2, 7 values1, 8 values2
    passes = [
        {"kind": hj.PASS_COMPRESS, "resources": [2, 0, 3]},
        {"kind": hj.PASS_KERNEL, "resources": [1, 2, 4], "ir": through(True), "size": n, "size_buffer": 0},
        {"kind": hj.PASS_COMPRESS, "resources": [6, 5, 4], "size_buffer": 0},
        {"kind": hj.PASS_KERNEL, "resources": [1, 2, 7], "ir": through(False), "size": n, "size_buffer": 0},
        {"kind": hj.PASS_KERNEL, "resources": [7, 6, 8], "ir": through(False), "size": n, "size_buffer": 5},
    ]
    descs = [(1, hj.U32, 4), (n, hj.I32, 4), (n, hj.U32, 4), (n, hj.BOOL, 1), (n, hj.BOOL, 1), (1, hj.U32, 4),
             (n, hj.U32, 4), (n, hj.I32, 4), (n, hj.I32, 4)]
    got = sharded.shard_plan(passes, descs, [A, S, A, S, A, A, A, A, A])
    assert got == [R, S, S, S, S, R, S, S, S]
    assert sharded.shard_plan(passes, descs, [A] * 9) == [R] * 9


def test_shard_plan_rejects_malformed_pass_lists():
    hj = importlib.import_module("hephaestus-jit_b200")
    passes, descs = _c2_like_passes(1 << 12)
    bad = [dict(passes[0], resources=[0, 99])]
    with pytest.raises(hj.HjError, match="out of range"):
        sharded.shard_plan(bad, descs, [sharded.RES_AUTO] * len(descs))
    with pytest.raises(hj.HjError, match="binds"):
        sharded.shard_plan([dict(passes[0], resources=[0])], descs, [sharded.RES_AUTO] * len(descs))


def test_hj_shard_bounds_matches_the_python_statement():
    import ctypes
    L = importlib.import_module("hephaestus-jit_b200._lib")
    for n in (0, 1, 7, 1000, (1 << 30) + 5):
        for world in (1, 2, 3, 8):
            for r in range(world):
                a, b = ctypes.c_uint64(), ctypes.c_uint64()
                L.lib.hj_shard_bounds(n, world, r, ctypes.byref(a), ctypes.byref(b))
                assert (a.value, b.value) == sharded.shard_bounds(n, world, r)
