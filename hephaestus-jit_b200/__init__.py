"""hephaestus-jit_b200 — B200-native (sm_100a) execution backend for hephaestus-jit.

Host-side mirror of the reference's backend interface
(hephaestus-jit/src/backend/mod.rs:26-155: ``Device`` / ``Buffer``) over the C ABI in
include/hj.h.  The directory name contains a hyphen, so import it with
``importlib.import_module("hephaestus-jit_b200")`` (tests/ and bench.py do).

Nothing here computes on the CPU: every operation is a call into libhj_b200.so, which
enqueues hand-written or NVRTC-compiled sm_100a kernels.  Without the shared library the
import fails; without a CUDA device ``Device.cuda`` raises ``HjError(ERR_NO_DEVICE)``.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import HjError, check, lib

__all__ = ["Device", "Buffer", "HjError", "device_count", "VOID", "BOOL", "I8", "U8", "I16", "U16",
           "I32", "U32", "I64", "U64", "F16", "F32", "F64", "MAX", "MIN", "SUM", "PROD", "OR", "AND",
           "XOR", "dtype_of", "type_of_dtype"]

# hj_type_kind — order of VarType (vartype.rs:89-104)
VOID, BOOL, I8, U8, I16, U16, I32, U32, I64, U64, F16, F32, F64 = range(13)
VEC, ARRAY, MAT, STRUCT = 13, 14, 15, 16
# hj_reduce_op — order of ReduceOp (op.rs:90-99)
MAX, MIN, SUM, PROD, OR, AND, XOR = range(7)

_NP = {BOOL: np.bool_, I8: np.int8, U8: np.uint8, I16: np.int16, U16: np.uint16, I32: np.int32,
       U32: np.uint32, I64: np.int64, U64: np.uint64, F16: np.float16, F32: np.float32,
       F64: np.float64}
_FROM_NP = {np.dtype(v): k for k, v in _NP.items()}
TYPE_SIZE = {VOID: 0, BOOL: 1, I8: 1, U8: 1, I16: 2, U16: 2, I32: 4, U32: 4, I64: 8, U64: 8, F16: 2,
             F32: 4, F64: 8}


def dtype_of(ty: int) -> np.dtype:
    return np.dtype(_NP[ty])


def type_of_dtype(dt) -> int:
    return _FROM_NP[np.dtype(dt)]


def device_count() -> int:
    return int(lib.hj_device_count())


class Buffer:
    """Mirror of ``backend::Buffer`` (backend/mod.rs:131-155): a ref-counted device allocation."""

    __slots__ = ("_h", "device", "__weakref__")

    def __init__(self, handle: int, device: "Device"):
        self._h = ctypes.c_void_p(handle)
        self.device = device

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:  # module globals are torn down before objects at interpreter exit
            lib.hj_buffer_release(h)

    @property
    def handle(self) -> ctypes.c_void_p:
        return self._h

    @property
    def size(self) -> int:
        out = ctypes.c_size_t()
        check(lib.hj_buffer_size(self._h, ctypes.byref(out)))
        return out.value

    @property
    def ptr(self) -> int:
        out = ctypes.c_void_p()
        check(lib.hj_buffer_device_ptr(self._h, ctypes.byref(out)))
        return out.value or 0

    def to_host(self, dtype, start: int = 0, end: int | None = None) -> np.ndarray:
        """``BackendBuffer::to_host::<T>(range)`` (backend/mod.rs:39): element range -> host."""
        dt = np.dtype(dtype)
        if end is None:
            end = self.size // dt.itemsize
        n = max(0, end - start)
        out = np.empty(n, dtype=dt)
        check(lib.hj_buffer_to_host(self._h, start * dt.itemsize, n * dt.itemsize,
                                    out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def upload(self, data: np.ndarray, offset_bytes: int = 0) -> None:
        data = np.ascontiguousarray(data)
        check(lib.hj_buffer_upload(self._h, offset_bytes, data.ctypes.data_as(ctypes.c_void_p),
                                   data.nbytes))

    def fill_zero(self) -> None:
        check(lib.hj_buffer_fill_zero(self._h))


class Device:
    """Mirror of ``backend::Device`` (backend/mod.rs:66-126), CUDA arm."""

    def __init__(self, handle: int, ordinal: int):
        self._h = ctypes.c_void_p(handle)
        self.ordinal = ordinal

    @staticmethod
    def cuda(ordinal: int = 0) -> "Device":
        """``Device::cuda(id)`` (backend/mod.rs:73-75)."""
        out = ctypes.c_void_p()
        check(lib.hj_device_create(ordinal, ctypes.byref(out)))
        return Device(out.value, ordinal)

    @property
    def handle(self) -> ctypes.c_void_p:
        return self._h

    # -- buffers ------------------------------------------------------------------------
    def create_buffer(self, size: int) -> Buffer:
        out = ctypes.c_void_p()
        check(lib.hj_buffer_create(self._h, size, ctypes.byref(out)))
        return Buffer(out.value, self)

    def create_buffer_from_slice(self, data) -> Buffer:
        a = np.ascontiguousarray(data)
        out = ctypes.c_void_p()
        check(lib.hj_buffer_create_from_slice(self._h, a.ctypes.data_as(ctypes.c_void_p), a.nbytes,
                                              ctypes.byref(out)))
        return Buffer(out.value, self)

    def wrap(self, device_ptr: int, nbytes: int) -> Buffer:
        out = ctypes.c_void_p()
        check(lib.hj_buffer_wrap(self._h, ctypes.c_void_p(device_ptr), nbytes, ctypes.byref(out)))
        return Buffer(out.value, self)

    # -- control ------------------------------------------------------------------------
    def sync(self) -> None:
        check(lib.hj_device_sync(self._h))

    @property
    def stream(self) -> int:
        out = ctypes.c_void_p()
        check(lib.hj_device_stream(self._h, ctypes.byref(out)))
        return out.value or 0

    def set_stream(self, stream: int | None) -> None:
        check(lib.hj_device_set_stream(self._h, ctypes.c_void_p(stream or 0)))

    def info(self) -> dict:
        sm, ma, mi = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        mem, l2 = ctypes.c_uint64(), ctypes.c_uint64()
        check(lib.hj_device_info(self._h, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi),
                                 ctypes.byref(mem), ctypes.byref(l2)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value,
                "l2_bytes": l2.value}

    def pool_stats(self) -> dict:
        v = [ctypes.c_uint64() for _ in range(4)]
        check(lib.hj_device_pool_stats(self._h, *[ctypes.byref(x) for x in v]))
        return dict(zip(("bytes_live", "bytes_cached", "n_alloc", "n_free"), (x.value for x in v)))

    def launch_count(self) -> int:
        out = ctypes.c_uint64()
        check(lib.hj_device_launch_count(self._h, ctypes.byref(out)))
        return out.value

    # -- device ops (direct entry points; execute_graph dispatches to the same kernels) ----
    def reduce(self, op: int, ty: int, n: int, src: Buffer, dst: Buffer) -> None:
        check(lib.hj_reduce(self._h, op, ty, n, src.handle, dst.handle))

    def prefix_sum(self, ty: int, n: int, inclusive: bool, src: Buffer, dst: Buffer,
                   seed: Buffer | None = None) -> None:
        check(lib.hj_prefix_sum(self._h, ty, n, int(inclusive), src.handle, dst.handle,
                                seed.handle if seed else None))

    def compress(self, n: int, out_count: Buffer, src_mask: Buffer, index_out: Buffer,
                 size_buf: Buffer | None = None, index_base: int = 0) -> None:
        check(lib.hj_compress(self._h, n, size_buf.handle if size_buf else None, out_count.handle,
                              src_mask.handle, index_out.handle, index_base))

    def scatter_reduce(self, op: int, ty: int, n: int, idx: Buffer, src: Buffer | None, literal,
                       dst: Buffer, n_dst: int) -> None:
        lit = int(np.array([literal], dtype=_NP[ty]).view(
            {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[TYPE_SIZE[ty]])[0])
        check(lib.hj_scatter_reduce(self._h, op, ty, n, idx.handle, src.handle if src else None, lit,
                                    dst.handle, n_dst))

    def gather(self, elem_bytes: int, n: int, src: Buffer, idx: Buffer, dst: Buffer) -> None:
        check(lib.hj_gather(self._h, elem_bytes, n, src.handle, idx.handle, dst.handle))

    # -- fused kernels (NVRTC path) -------------------------------------------------------
    def kernel(self, ir) -> "Kernel":
        """Compile (or fetch from the cache keyed by IR hash) the fused kernel for ``ir``
        (an ``_lib.Ir`` or an ``ir.IRBuilder``)."""
        keep = None
        if hasattr(ir, "build"):
            keep, ir = ir, ir.build()
        out = ctypes.c_void_p()
        check(lib.hj_kernel_get(self._h, ctypes.byref(ir), ctypes.byref(out)))
        return Kernel(out.value, keep)

    def launch(self, kernel: "Kernel", size: int, buffers, size_buf: Buffer | None = None,
               index_base: int = 0) -> None:
        arr = (ctypes.c_void_p * max(len(buffers), 1))(*[b.handle for b in buffers])
        check(lib.hj_kernel_launch(self._h, kernel.handle, size, size_buf.handle if size_buf else None,
                                   arr, len(buffers), index_base))

    def map_host(self, kernel: "Kernel", n: int, host_arrays, chunk_elems: int = 0) -> None:
        """Out-of-core elementwise map: ``host_arrays[i]`` (numpy arrays or raw addresses, pinned for
        full speed) is buffer slot ``i`` of the kernel; upload, kernel and download are pipelined
        chunk by chunk (hj_kernel_map_host)."""
        ptrs = [a.ctypes.data if hasattr(a, "ctypes") else int(a) for a in host_arrays]
        arr = (ctypes.c_void_p * max(len(ptrs), 1))(*ptrs)
        check(lib.hj_kernel_map_host(self._h, kernel.handle, n, arr, len(ptrs), chunk_elems))

    # -- device ops over HOST arrays (chunked, upload | kernel | download overlapped) -------------
    @staticmethod
    def _host_ptr(a):
        return a.ctypes.data if hasattr(a, "ctypes") else int(a)

    def reduce_host(self, op: int, ty: int, n: int, host_src, chunk_elems: int = 0):
        """``hj_reduce_host``: fold of a host array (numpy array or raw address); returns the scalar."""
        out = np.zeros(1, dtype=_NP[ty])
        check(lib.hj_reduce_host(self._h, op, ty, n, self._host_ptr(host_src), out.ctypes.data, chunk_elems))
        return out[0]

    def prefix_sum_host(self, ty: int, n: int, inclusive: bool, host_src, host_dst, chunk_elems: int = 0) -> None:
        check(lib.hj_prefix_sum_host(self._h, ty, n, int(inclusive), self._host_ptr(host_src), self._host_ptr(host_dst),
                                     chunk_elems))

    def compress_host(self, n: int, host_mask, host_index_out, index_base: int = 0, chunk_elems: int = 0) -> int:
        """``hj_compress_host``: returns the count; ``host_index_out[:count]`` holds the indices."""
        cnt = ctypes.c_uint32()
        check(lib.hj_compress_host(self._h, n, self._host_ptr(host_mask), self._host_ptr(host_index_out),
                                   ctypes.byref(cnt), index_base, chunk_elems))
        return cnt.value

    def kernel_cache_stats(self) -> dict:
        v = [ctypes.c_uint64() for _ in range(3)]
        check(lib.hj_device_kernel_cache_stats(self._h, *[ctypes.byref(x) for x in v]))
        return dict(zip(("compiled", "hits", "disk_hits"), (x.value for x in v)))

    def graph_cache_stats(self):
        """(captured, replayed, plain) counters of ``hj_execute_graph_cached``."""
        c, r, p = _lib._u64(), _lib._u64(), _lib._u64()
        check(lib.hj_graph_cache_stats(self._h, ctypes.byref(c), ctypes.byref(r), ctypes.byref(p)))
        return c.value, r.value, p.value

    def execute_graph(self, passes, env, descs, timed: bool = False, graph_key: int = 0):
        """``BackendDevice::execute_graph`` (backend/mod.rs:33).

        passes: list of dicts {kind, arg, resources, size_buffer, ir (IRBuilder|None), size};
        env: list of Buffer (or None); descs: list of (size_elems, ty, elem_bytes).
        Returns the list of (name, start_us, duration_us) when ``timed``.  A non-zero
        ``graph_key`` names a pass list that is launched repeatedly: from the second launch with
        the same buffers on it is replayed as one captured CUDA graph (returns 0 / 1 / 2 =
        executed / captured / replayed)."""
        c_passes, n, c_env, c_desc, _keep = marshal_graph(passes, env, descs)
        report = _lib.Report()
        reps = (_lib.PassReport * max(n, 1))()
        if timed:
            report.passes = reps
            report.passes_capacity = n
        if graph_key and not timed:
            how = ctypes.c_uint32()
            check(lib.hj_execute_graph_cached(self._h, graph_key, c_passes, n, c_env, c_desc, len(env), ctypes.byref(how)))
            return how.value
        check(lib.hj_execute_graph(self._h, c_passes, n, c_env, c_desc, len(env), ctypes.byref(report)))
        if timed:
            return [(reps[i].name.decode(), reps[i].start_us, reps[i].duration_us) for i in range(n)]
        return None


PASS_KERNEL, PASS_REDUCE, PASS_PREFIX_SUM, PASS_COMPRESS = range(4)


def marshal_graph(passes, env, descs):
    """C views of a pass list / environment (see ``Device.execute_graph``); the last element of the
    returned tuple keeps the ctypes arrays alive."""
    n = len(passes)
    c_passes = (_lib.Pass * max(n, 1))()
    keep = []
    for i, p in enumerate(passes):
        res = (ctypes.c_uint32 * max(len(p["resources"]), 1))(*p["resources"])
        ir_ptr = None
        if p.get("ir") is not None:
            ir = p["ir"].build() if hasattr(p["ir"], "build") else p["ir"]
            keep.append(ir)
            ir_ptr = ctypes.pointer(ir)
        keep.append(res)
        c_passes[i] = _lib.Pass(p["kind"], p.get("arg", 0), res, len(p["resources"]),
                                p.get("size_buffer", -1) if p.get("size_buffer") is not None else -1,
                                ir_ptr, p.get("size", 0))
    c_env = (ctypes.c_void_p * max(len(env), 1))(*[(b.handle if b is not None else None) for b in env])
    c_desc = (_lib.BufferDesc * max(len(descs), 1))(*[_lib.BufferDesc(*d) for d in descs])
    return c_passes, n, c_env, c_desc, keep


class PreparedGraph:
    """A pass list marshalled ONCE for repeated launches (what Graph::launch_with rebuilds per call is a
    few hundred bytes; through ctypes that costs more than a 2^27-element kernel runs).  ``comm``
    (``sharded.Comm``) selects ``hj_execute_graph_sharded``; ``placement`` / ``seeds`` as in
    ``Comm.execute_graph``."""

    def __init__(self, device: "Device", passes, env, descs, comm=None, placement=None, seeds=None, graph_key: int = 0):
        self.device, self.comm, self.graph_key = device, comm, graph_key
        self.how = ctypes.c_uint32()
        self._keep_env = list(env)
        self.c_passes, self.n, self.c_env, self.c_desc, self._keep = marshal_graph(passes, env, descs)
        self.n_res = len(env)
        self.report = _lib.Report()
        self.reps = (_lib.PassReport * max(self.n, 1))()
        self.shards = None
        if comm is not None:
            self.shards = (_lib.ShardDesc * max(self.n_res, 1))()
            self._seeds = list(seeds) if seeds else [None] * self.n_res
            for i in range(self.n_res):
                self.shards[i].placement = placement[i] if placement else _lib.RES_AUTO
                self.shards[i].seed = self._seeds[i].handle if self._seeds[i] is not None else None
            check(lib.hj_shard_plan(self.c_passes, self.n, self.c_desc, self.n_res, self.shards))
            self._deferred_in = [0] * self.n_res

    def run(self, timed: bool = False):
        """One launch; with ``timed`` blocks and returns [(pass name, start_us, duration_us)]."""
        self.report.passes = self.reps if timed else None
        self.report.passes_capacity = self.n if timed else 0
        cached = self.graph_key and not timed   # relaunch path: replay of one captured CUDA graph
        if self.comm is None:
            if cached:
                check(lib.hj_execute_graph_cached(self.device.handle, self.graph_key, self.c_passes, self.n, self.c_env,
                                                  self.c_desc, self.n_res, ctypes.byref(self.how)))
            else:
                check(lib.hj_execute_graph(self.device.handle, self.c_passes, self.n, self.c_env, self.c_desc, self.n_res,
                                           ctypes.byref(self.report) if timed else None))
        else:
            for i in range(self.n_res):
                self.shards[i].deferred = self._deferred_in[i]   # the inputs' state; outputs are rewritten by the call
            if cached:
                check(lib.hj_execute_graph_sharded_cached(self.comm.handle, self.graph_key, self.c_passes, self.n, self.c_env,
                                                          self.c_desc, self.n_res, self.shards, ctypes.byref(self.how)))
            else:
                check(lib.hj_execute_graph_sharded(self.comm.handle, self.c_passes, self.n, self.c_env, self.c_desc,
                                                   self.n_res, self.shards, ctypes.byref(self.report) if timed else None))
        if timed:
            return [(self.reps[i].name.decode(), self.reps[i].start_us, self.reps[i].duration_us) for i in range(self.n)]
        return None

    def deferred(self):
        return [self.shards[i].deferred == 1 for i in range(self.n_res)] if self.shards is not None else [False] * self.n_res

    def segments(self):
        """Per resource: True when it is a per-rank compacted segment after the last launch (``HJ_SHARD_SEGMENT``:
        the index output of a sharded Compress, its seed buffer holds the rank's own count)."""
        return [self.shards[i].deferred in (2, 3) for i in range(self.n_res)] if self.shards is not None else [False] * self.n_res

    def placement(self):
        return [self.shards[i].placement for i in range(self.n_res)] if self.shards is not None else [0] * self.n_res


class Kernel:
    """A compiled fused kernel (NVRTC -> sm_100a cubin), owned by the per-device cache."""

    def __init__(self, handle: int, keep=None):
        self._h = ctypes.c_void_p(handle)
        self._keep = keep

    @property
    def handle(self) -> ctypes.c_void_p:
        return self._h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:  # module globals are torn down before objects at interpreter exit
            lib.hj_kernel_release(h)
