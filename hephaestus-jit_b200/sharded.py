"""Multi-GPU layer (host side): contiguous sharding of 1-D arrays over the ranks of one node and
the NCCL-backed combine ops of include/hj.h (``hj_sharded_*``).

One process per GPU.  The reference has no multi-GPU code; the partitioning
follows BASELINE.json: reductions all-reduce their partials, scans and compaction apply the
exclusive scan of the per-GPU totals as an offset, histograms are privatised per GPU.

``exchange_*`` are the host-visible statement of the exchange protocol (what crosses the fabric
and how it is folded); they run on any torch.distributed backend and are what the world-size-2
gloo tests exercise on CPU.  On GPUs the same protocol runs inside libhj_b200.so on the device
stream (csrc/comm.cu) — nothing is copied to the host.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import Buffer, Device, TYPE_SIZE, _NP, _lib, marshal_graph
from ._lib import HJ_IPC_HANDLE_BYTES, HJ_UNIQUE_ID_BYTES, RES_AUTO, RES_REPLICATED, RES_SHARDED, check, lib


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of rank ``rank``: sizes differ by at most one, order preserved."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def exclusive_offsets(totals):
    """offset[r] = fold(totals[0..r)) — the carry each rank seeds its scan / compaction with."""
    out, acc = [], type(totals[0])(0) if len(totals) else 0
    for t in totals:
        out.append(acc)
        acc = acc + t
    return out


class Comm:
    """An NCCL communicator bound to one Device (``hj_comm``)."""

    def __init__(self, dev: Device, unique_id: bytes, rank: int, world: int):
        assert len(unique_id) == HJ_UNIQUE_ID_BYTES
        self.dev, self.rank, self.world = dev, rank, world
        out = ctypes.c_void_p()
        buf = (ctypes.c_uint8 * HJ_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        check(lib.hj_comm_create(dev.handle, buf, rank, world, ctypes.byref(out)))
        self._h = out

    @staticmethod
    def unique_id() -> bytes:
        buf = (ctypes.c_uint8 * HJ_UNIQUE_ID_BYTES)()
        check(lib.hj_comm_unique_id(buf))
        return bytes(buf)

    @staticmethod
    def from_torch(dev: Device) -> "Comm":
        """Create the communicator of the current torch.distributed job (id broadcast from rank 0)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return Comm(dev, box[0], rank, world)

    @staticmethod
    def local(dev: Device, rank: int, world: int, all_gather_bytes) -> "Comm":
        """The same communicator WITHOUT NCCL (``hj_comm_create_local`` + ``hj_comm_connect``): every rank
        exports the CUDA-IPC handle of its mailbox, ``all_gather_bytes(handle: bytes) -> list[bytes]``
        (any host transport: torch.distributed over gloo, a multiprocessing queue, ...) returns the
        handles of all ranks in rank order.  Ranks may share one GPU."""
        self = Comm.__new__(Comm)
        self.dev, self.rank, self.world = dev, rank, world
        out = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * HJ_IPC_HANDLE_BYTES)()
        check(lib.hj_comm_create_local(dev.handle, rank, world, ctypes.byref(out), handle))
        self._h = out
        handles = all_gather_bytes(bytes(handle))
        assert len(handles) == world and all(len(h) == HJ_IPC_HANDLE_BYTES for h in handles)
        blob = (ctypes.c_uint8 * (HJ_IPC_HANDLE_BYTES * world)).from_buffer_copy(b"".join(handles))
        check(lib.hj_comm_connect(self._h, blob))
        return self

    @staticmethod
    def local_from_torch(dev: Device) -> "Comm":
        """``Comm.local`` with the handles all-gathered over the current torch.distributed group (any
        backend; the handles travel as host bytes)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()

        def gather(handle: bytes):
            box = [None] * world
            dist.all_gather_object(box, handle)
            return box
        return Comm.local(dev, rank, world, gather)

    @property
    def handle(self):
        return self._h

    def info(self) -> dict:
        v = [ctypes.c_int32() for _ in range(4)]
        check(lib.hj_comm_info(self._h, *[ctypes.byref(x) for x in v]))
        return dict(zip(("rank", "world", "peer_memory", "nccl"), (x.value for x in v)))

    def destroy(self):
        h, self._h = self._h, None
        if h:
            check(lib.hj_comm_destroy(h))

    def bounds(self, n: int) -> tuple[int, int]:
        return shard_bounds(n, self.world, self.rank)

    # ---- sharded ops (device-side exchange) ---------------------------------------------------
    def reduce(self, op: int, ty: int, n_local: int, src: Buffer, dst: Buffer) -> None:
        check(lib.hj_sharded_reduce(self._h, op, ty, n_local, src.handle, dst.handle))

    def prefix_sum(self, ty: int, n_local: int, inclusive: bool, src: Buffer, dst: Buffer) -> None:
        check(lib.hj_sharded_prefix_sum(self._h, ty, n_local, int(inclusive), src.handle, dst.handle))

    def prefix_sum_deferred(self, ty: int, n_local: int, inclusive: bool, src: Buffer, dst: Buffer, seed_out: Buffer) -> None:
        """Local scan + this rank's exclusive offset in ``seed_out`` (one kernel, 2 x sizeof(T) bytes per
        element); the global scan is ``dst[i] + seed_out[0]``."""
        check(lib.hj_sharded_prefix_sum_deferred(self._h, ty, n_local, int(inclusive), src.handle, dst.handle,
                                                 seed_out.handle))

    def execute_graph(self, passes, env, descs, placement, seeds=None, timed: bool = False):
        """``hj_execute_graph_sharded``: ``placement[i]`` in {RES_REPLICATED, RES_SHARDED, RES_AUTO} per
        resource (AUTO ones are planned with ``hj_shard_plan`` first), ``seeds[i]`` an optional
        one-element Buffer that lets a sharded integer scan into resource i stay deferred (and that receives
        the rank's own count when resource i is the index segment of a sharded Compress).  Returns
        ``(placement, deferred, report)`` after the call."""
        c_passes, n, c_env, c_desc, _keep = marshal_graph(passes, env, descs)
        nres = len(env)
        sh = (_lib.ShardDesc * max(nres, 1))()
        for i in range(nres):
            sh[i].placement = placement[i]
            sh[i].deferred = 0
            sh[i].seed = seeds[i].handle if seeds and seeds[i] is not None else None
        check(lib.hj_shard_plan(c_passes, n, c_desc, nres, sh))
        report = _lib.Report()
        reps = (_lib.PassReport * max(n, 1))()
        if timed:
            report.passes = reps
            report.passes_capacity = n
        check(lib.hj_execute_graph_sharded(self._h, c_passes, n, c_env, c_desc, nres, sh, ctypes.byref(report)))
        rep = [(reps[i].name.decode(), reps[i].start_us, reps[i].duration_us) for i in range(n)] if timed else None
        return [sh[i].placement for i in range(nres)], [sh[i].deferred == 1 for i in range(nres)], rep

    def compress(self, n_local: int, index_base: int, src_mask: Buffer, index_out: Buffer, out_count: Buffer,
                 counts_out: Buffer | None = None) -> None:
        check(lib.hj_sharded_compress(self._h, n_local, index_base, src_mask.handle, index_out.handle,
                                      out_count.handle, counts_out.handle if counts_out else None))

    def scatter_reduce(self, op: int, ty: int, n_local: int, idx: Buffer, src: Buffer | None, literal,
                       dst: Buffer, n_dst: int) -> None:
        lit = int(np.array([literal], dtype=_NP[ty]).view(
            {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[TYPE_SIZE[ty]])[0])
        check(lib.hj_sharded_scatter_reduce(self._h, op, ty, n_local, idx.handle,
                                            src.handle if src else None, lit, dst.handle, n_dst))

    def rebalance(self, elem_bytes: int, src: Buffer, counts: Buffer, dst: Buffer,
                  out_count: Buffer | None = None) -> int:
        """Even re-partition of a sharded compacted sequence (order preserved); returns this
        rank's new element count."""
        n = ctypes.c_uint64()
        check(lib.hj_sharded_rebalance(self._h, elem_bytes, src.handle, counts.handle, dst.handle,
                                       out_count.handle if out_count else None, ctypes.byref(n)))
        return n.value


def shard_plan(passes, descs, placement):
    """Host-only placement propagation (``hj_shard_plan``): returns the placement of every resource."""
    c_passes, n, _env, c_desc, _keep = marshal_graph(passes, [None] * len(descs), descs)
    sh = (_lib.ShardDesc * max(len(descs), 1))()
    for i, p in enumerate(placement):
        sh[i].placement = p
    check(lib.hj_shard_plan(c_passes, n, c_desc, len(descs), sh))
    return [sh[i].placement for i in range(len(descs))]


def rebalance_plan(counts, rank: int):
    """What rank ``rank`` sends and receives when the concatenation of per-rank segments of lengths
    ``counts`` is re-partitioned into ``shard_bounds`` blocks: two lists of
    ``(peer, offset in my segment / my block, length)``; ``peer == rank`` is the local copy."""
    world = len(counts)
    src0 = [0] + list(np.cumsum(np.asarray(counts, dtype=np.int64)))
    total = int(src0[-1])
    dst = [shard_bounds(total, world, q) for q in range(world)]
    sends, recvs = [], []
    for q in range(world):
        lo, hi = max(src0[rank], dst[q][0]), min(src0[rank + 1], dst[q][1])
        if lo < hi:
            sends.append((q, int(lo - src0[rank]), int(hi - lo)))
        lo, hi = max(src0[q], dst[rank][0]), min(src0[q + 1], dst[rank][1])
        if lo < hi:
            recvs.append((q, int(lo - dst[rank][0]), int(hi - lo)))
    return sends, recvs


# ---- the exchange protocol on host tensors (any backend; used by the gloo tests) ---------------

def _all_gather_scalar(value: np.generic):
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    raw = np.frombuffer(np.asarray(value).tobytes().ljust(8, b"\0"), dtype=np.uint8).copy()
    mine = torch.from_numpy(raw)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    n = np.asarray(value).dtype.itemsize
    return [np.frombuffer(o.numpy().tobytes()[:n], dtype=np.asarray(value).dtype)[0] for o in outs]


def exchange_reduce(partial: np.generic, fold):
    """All-gather the per-rank partials, fold them in rank order on every rank."""
    parts = _all_gather_scalar(partial)
    acc = parts[0]
    for p in parts[1:]:
        acc = fold(acc, p)
    return acc


def exchange_scan_offset(local_total: np.generic):
    """All-gather the shard totals; this rank's scan is seeded with the exclusive prefix."""
    import torch.distributed as dist
    with np.errstate(over="ignore"):
        return exclusive_offsets(_all_gather_scalar(local_total))[dist.get_rank()]


def exchange_counts(local_count: int):
    """All-gather the per-rank compaction counts: (counts, exclusive offsets, global count)."""
    counts = [int(c) for c in _all_gather_scalar(np.uint32(local_count))]
    return counts, exclusive_offsets(counts), sum(counts)


def exchange_rebalance(local: np.ndarray) -> np.ndarray:
    """Host statement of ``hj_sharded_rebalance``: all-gather the counts, then point-to-point
    transfers of exactly the overlapping slices."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    counts, _, total = exchange_counts(len(local))
    lo, hi = shard_bounds(total, len(counts), rank)
    out = np.empty(hi - lo, dtype=local.dtype)
    sends, recvs = rebalance_plan(counts, rank)
    ops, keep = [], []
    for peer, off, n in recvs:
        if peer == rank:
            continue
        t = torch.empty(n * local.dtype.itemsize, dtype=torch.uint8)
        keep.append((t, off, n))
        ops.append(dist.P2POp(dist.irecv, t, peer))
    for peer, off, n in sends:
        if peer == rank:
            src_off = off
            dst_off = next(o for p, o, _ in recvs if p == rank)
            out[dst_off:dst_off + n] = local[src_off:src_off + n]
            continue
        t = torch.from_numpy(np.frombuffer(local[off:off + n].tobytes(), dtype=np.uint8).copy())
        ops.append(dist.P2POp(dist.isend, t, peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for t, off, n in keep:
        out[off:off + n] = np.frombuffer(t.numpy().tobytes(), dtype=local.dtype)
    return out
