"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the sharded reduce / scan /
compress / histogram of include/hj.h (local sm_100a kernel + NCCL exchange on the device stream)
against the oracle on the whole array.  One process per GPU, spawned here; the NCCL unique id is
handed to the workers directly (no torch.distributed needed)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hj = importlib.import_module("hephaestus-jit_b200")
sharded = importlib.import_module("hephaestus-jit_b200.sharded")


def _worker(rank, world, uid, n):
    sys.path.insert(0, ROOT)
    import oracle
    hjw = importlib.import_module("hephaestus-jit_b200")
    sh = importlib.import_module("hephaestus-jit_b200.sharded")
    dev = hjw.Device.cuda(rank)
    comm = sh.Comm(dev, uid, rank, world)
    try:
        rng = np.random.Generator(np.random.PCG64(7))
        u = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        f = rng.random(n, dtype=np.float32)
        mask = (rng.random(n) < 0.4).astype(np.uint8)
        keys = rng.integers(0, 1 << 16, size=n).astype(np.uint32)
        s, e = sh.shard_bounds(n, world, rank)
        nl = e - s
        out1 = dev.create_buffer(8)
        for op, ty, arr in ((hjw.SUM, hjw.U32, u), (hjw.MAX, hjw.U32, u), (hjw.XOR, hjw.U32, u), (hjw.MIN, hjw.F32, f)):
            comm.reduce(op, ty, nl, dev.create_buffer_from_slice(arr[s:e]), out1)
            assert out1.to_host(arr.dtype, 0, 1)[0] == oracle.reduce(op, ty, arr)[0], (op, ty)
        comm.reduce(hjw.SUM, hjw.F32, nl, dev.create_buffer_from_slice(f[s:e]), out1)
        exact = float(f.astype(np.float64).sum())
        assert abs(float(out1.to_host(np.float32, 0, 1)[0]) - exact) <= 1e-5 * exact
        # scan: this rank's slice of the global scan, bit-exact
        dst = dev.create_buffer(4 * nl)
        for inclusive in (True, False):
            comm.prefix_sum(hjw.U32, nl, inclusive, dev.create_buffer_from_slice(u[s:e]), dst)
            assert np.array_equal(dst.to_host(np.uint32), oracle.prefix_sum(oracle.U32, u, inclusive)[s:e])
        # compress: per-rank segment with global indices + counts table
        idx = dev.create_buffer_from_slice(np.zeros(nl, np.uint32))
        cnt, counts = dev.create_buffer(4), dev.create_buffer(4 * world)
        comm.compress(nl, s, dev.create_buffer_from_slice(mask[s:e]), idx, cnt, counts)
        gcnt, gidx = oracle.compress(mask)
        c = counts.to_host(np.uint32)
        off = int(c[:rank].sum())
        assert int(cnt.to_host(np.uint32)[0]) == gcnt
        assert np.array_equal(idx.to_host(np.uint32)[: int(c[rank])], gidx[off: off + int(c[rank])])
        # histogram: privatised + all-reduce
        hist = dev.create_buffer_from_slice(np.zeros(1 << 16, np.uint32))
        comm.scatter_reduce(hjw.SUM, hjw.U32, nl, dev.create_buffer_from_slice(keys[s:e]), None, 1, hist, 1 << 16)
        assert np.array_equal(hist.to_host(np.uint32), oracle.histogram_u32_mt(keys, 1 << 16))
        ored = dev.create_buffer_from_slice(np.zeros(1 << 10, np.uint32))
        comm.scatter_reduce(hjw.OR, hjw.U32, nl, dev.create_buffer_from_slice(keys[s:e] & 1023),
                            dev.create_buffer_from_slice(u[s:e]), 0, ored, 1 << 10)
        want = np.zeros(1 << 10, np.uint32)
        np.bitwise_or.at(want, keys & 1023, u)
        assert np.array_equal(ored.to_host(np.uint32), want)
        # array exchange over peer memory with other types / operators, and repeated calls (box parities)
        for rep in range(3):
            fv = rng.random(n, dtype=np.float32)
            facc = dev.create_buffer_from_slice(np.zeros(4096, np.float32))
            comm.scatter_reduce(hjw.SUM, hjw.F32, nl, dev.create_buffer_from_slice(keys[s:e] & 4095),
                                dev.create_buffer_from_slice(fv[s:e]), 0.0, facc, 4096)
            wantf = np.zeros(4096, np.float64)
            np.add.at(wantf, keys & 4095, fv.astype(np.float64))
            assert np.allclose(facc.to_host(np.float32), wantf, rtol=1e-4)
            iv = rng.integers(-2**31, 2**31, size=n, dtype=np.int64).astype(np.int32)
            imin = dev.create_buffer_from_slice(np.full(2048, 2**31 - 1, np.int32))
            comm.scatter_reduce(hjw.MIN, hjw.I32, nl, dev.create_buffer_from_slice(keys[s:e] & 2047),
                                dev.create_buffer_from_slice(iv[s:e]), 0, imin, 2048)
            wanti = np.full(2048, 2**31 - 1, np.int32)
            np.minimum.at(wanti, keys & 2047, iv)
            assert np.array_equal(imin.to_host(np.int32), wanti)
        # rebalance: a skewed compaction re-partitioned evenly, order preserved
        skew = (rng.random(n) < np.linspace(0.9, 0.02, n)).astype(np.uint8)
        idx = dev.create_buffer_from_slice(np.zeros(nl, np.uint32))
        comm.compress(nl, s, dev.create_buffer_from_slice(skew[s:e]), idx, cnt, counts)
        gcnt, gidx = oracle.compress(skew)
        lo, hi = sh.shard_bounds(gcnt, world, rank)
        bal, new_cnt = dev.create_buffer(4 * max(hi - lo, 1)), dev.create_buffer(4)
        got = comm.rebalance(4, idx, counts, bal, new_cnt)
        assert got == hi - lo and int(new_cnt.to_host(np.uint32)[0]) == hi - lo
        assert np.array_equal(bal.to_host(np.uint32)[: hi - lo], gidx[lo:hi])
        small = dev.create_buffer(4)
        with pytest.raises(hjw.HjError, match="balanced block"):
            comm.rebalance(4, idx, counts, small, None)
    finally:
        comm.destroy()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharded_ops_match_oracle(world):
    if hj.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    uid = sharded.Comm.unique_id()
    mp.spawn(_worker, args=(world, uid, (1 << 20) + 77), nprocs=world, join=True)
