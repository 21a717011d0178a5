"""``tr`` — host-side mirror of the reference's tracing API (hephaestus-jit/src/trace.rs,
re-exported as ``hephaestus_jit::tr``), over the C++ trace / schedule / graph layer in
libhj_b200.so (csrc/trace.cpp, tgraph.cpp).  Same names, argument meaning and error behaviour,
so the parity tests in tests/test_trace_*.py read like hephaestus-jit/src/test.rs:

    i = tr.sized_index(10); j = tr.sized_index(5)
    j.add(tr.literal(1, U32)).scatter(i, j)
    j.schedule()
    graph = tr.compile(); graph.launch(device)
    i.to_vec()  # [1, 2, 3, 4, 5, 5, 6, 7, 8, 9]

Nothing here computes: every call appends a variable to the trace held by the library; graphs
execute on the GPU through hj_execute_graph.
"""
from __future__ import annotations

import ctypes
import struct as _struct

import numpy as np

from . import _lib
from . import (BOOL, F16, F32, F64, I8, I16, I32, I64, U8, U16, U32, U64, VOID, MAX, MIN, SUM, PROD, OR, AND, XOR,
               Buffer, Device, _NP, _FROM_NP)
from ._lib import check, lib

# hj_bop / hj_uop numbering (op.rs:2-46)
(B_ADD, B_SUB, B_MUL, B_DIV, B_MOD, B_MIN, B_MAX, B_INNER, B_AND, B_OR, B_XOR, B_SHL, B_SHR, B_EQ, B_NEQ, B_LT, B_LE,
 B_GT, B_GE) = range(19)
U_CAST, U_BITCAST, U_NEG, U_SQRT, U_ABS, U_SIN, U_COS, U_EXP2, U_LOG2 = range(9)

_u64 = ctypes.c_uint64
_PACK = {BOOL: _struct.Struct("<?"), I8: _struct.Struct("<b"), U8: _struct.Struct("<B"), I16: _struct.Struct("<h"),
         U16: _struct.Struct("<H"), I32: _struct.Struct("<i"), U32: _struct.Struct("<I"), I64: _struct.Struct("<q"),
         U64: _struct.Struct("<Q"), F16: _struct.Struct("<e"), F32: _struct.Struct("<f"), F64: _struct.Struct("<d")}


def _bits(value, ty: int) -> int:
    """The u64 the reference stores for a literal (trace.rs:604-606: the value's bytes, zero padded)."""
    fmt = _PACK.get(ty)
    if fmt is not None and type(value) in (int, float, bool):  # plain Python scalars: no numpy round trip
        try:
            return int.from_bytes(fmt.pack(value), "little")
        except (_struct.error, OverflowError, TypeError):
            pass  # out of range / float into an integer type: numpy's conversion rules decide
    a = np.array([value], dtype=_NP[ty])
    return int.from_bytes(a.tobytes().ljust(8, b"\0"), "little")


def _infer_type(value) -> int:
    # Rust's literal defaults: integer literals are i32, float literals f64; we follow i32 / f32
    # (every float test in the reference writes an explicit f32 suffix)
    if isinstance(value, (bool, np.bool_)):
        return BOOL
    if isinstance(value, (int, np.integer)):
        return type_of_np(value) if isinstance(value, np.integer) else I32
    if isinstance(value, (float, np.floating)):
        return type_of_np(value) if isinstance(value, np.floating) else F32
    raise TypeError(f"cannot make a literal from {type(value)}")


def type_of_np(value) -> int:
    return _FROM_NP[np.dtype(type(value))]


# ---- types -----------------------------------------------------------------------------------
def vector(ty: int, num: int) -> int:
    return lib.hj_tr_type_vector(ty, num)


def array_type(ty: int, num: int) -> int:
    return lib.hj_tr_type_array(ty, num)


def matrix(ty: int, cols: int, rows: int) -> int:
    return lib.hj_tr_type_matrix(ty, cols, rows)


def struct(tys) -> int:
    arr = (ctypes.c_uint32 * len(tys))(*tys)
    return lib.hj_tr_type_struct(arr, len(tys))


def type_size(ty: int) -> int:
    return lib.hj_tr_type_size(ty)


def type_alignment(ty: int) -> int:
    return lib.hj_tr_type_alignment(ty)


def type_offset(ty: int, elem: int) -> int:
    return lib.hj_tr_type_offset(ty, elem)


class VarRef:
    """Mirror of ``tr::VarRef`` (trace.rs:244-291): an owned reference to a trace variable."""

    __slots__ = ("_id", "_keepalive", "__weakref__")

    def __init__(self, handle: int):
        self._id = int(handle)

    def __del__(self):
        h, self._id = getattr(self, "_id", 0), 0
        if h:
            try:
                lib.hj_tr_var_release(h)
            except Exception:
                pass

    # -- plumbing -----------------------------------------------------------------------------
    def id(self) -> int:
        return self._id

    def clone(self) -> "VarRef":
        check(lib.hj_tr_var_retain(self._id))
        return VarRef(self._id)

    def _info(self):
        ty, dyn, ext, ev, rc, dirty = (ctypes.c_uint32(), ctypes.c_int32(), _u64(), ctypes.c_int32(), _u64(),
                                       ctypes.c_int32())
        check(lib.hj_tr_var_info(self._id, ctypes.byref(ty), ctypes.byref(dyn), ctypes.byref(ext), ctypes.byref(ev),
                                 ctypes.byref(rc), ctypes.byref(dirty)))
        return ty.value, bool(dyn.value), ext.value, bool(ev.value), rc.value, bool(dirty.value)

    def ty(self) -> int:
        return self._info()[0]

    def capacity(self) -> int:
        return self._info()[2]

    def size(self) -> int:
        ty, dyn, ext, *_ = self._info()
        if dyn:
            raise NotImplementedError("size() of a dynamically sized variable (todo!() in the reference, trace.rs:1380)")
        return ext

    def is_evaluated(self) -> bool:
        return self._info()[3]

    def is_unsized(self) -> bool:
        return self._info()[2] == 0

    def is_dynamic(self) -> bool:
        return self._info()[1]

    def rc(self) -> int:
        return self._info()[4]

    def dirty(self) -> bool:
        return self._info()[5]

    def hash(self) -> int:
        return lib.hj_tr_var_hash(self._id)

    def buffer(self):
        out = ctypes.c_void_p()
        check(lib.hj_tr_var_buffer(self._id, ctypes.byref(out)))
        return out.value

    def schedule(self) -> None:
        check(lib.hj_tr_schedule(self._id))

    # -- elementwise (trace.rs:968-1082) ------------------------------------------------------
    def _bop(self, op: int, rhs) -> "VarRef":
        if not isinstance(rhs, VarRef):  # the type is only needed to turn a plain value into a literal
            rhs = _into(rhs, self.ty())
        out = _u64()
        check(lib.hj_tr_bop(op, self._id, rhs._id, ctypes.byref(out)))
        return VarRef(out.value)

    def _uop(self, op: int) -> "VarRef":
        out = _u64()
        check(lib.hj_tr_uop(op, self._id, ctypes.byref(out)))
        return VarRef(out.value)

    def add(self, rhs): return self._bop(B_ADD, rhs)
    def sub(self, rhs): return self._bop(B_SUB, rhs)
    def mul(self, rhs): return self._bop(B_MUL, rhs)
    def div(self, rhs): return self._bop(B_DIV, rhs)
    def modulus(self, rhs): return self._bop(B_MOD, rhs)
    def min(self, rhs): return self._bop(B_MIN, rhs)
    def max(self, rhs): return self._bop(B_MAX, rhs)
    def inner(self, rhs): return self._bop(B_INNER, rhs)
    def and_(self, rhs): return self._bop(B_AND, rhs)
    def or_(self, rhs): return self._bop(B_OR, rhs)
    def xor(self, rhs): return self._bop(B_XOR, rhs)
    def shl(self, rhs): return self._bop(B_SHL, rhs)
    def shr(self, rhs): return self._bop(B_SHR, rhs)
    def eq(self, rhs): return self._bop(B_EQ, rhs)
    def neq(self, rhs): return self._bop(B_NEQ, rhs)
    def lt(self, rhs): return self._bop(B_LT, rhs)
    def le(self, rhs): return self._bop(B_LE, rhs)
    def gt(self, rhs): return self._bop(B_GT, rhs)
    def ge(self, rhs): return self._bop(B_GE, rhs)
    def neg(self): return self._uop(U_NEG)
    def sqrt(self): return self._uop(U_SQRT)
    def abs(self): return self._uop(U_ABS)
    def sin(self): return self._uop(U_SIN)
    def cos(self): return self._uop(U_COS)
    def exp2(self): return self._uop(U_EXP2)
    def log2(self): return self._uop(U_LOG2)

    def cast(self, ty: int) -> "VarRef":
        out = _u64()
        check(lib.hj_tr_cast(self._id, ty, ctypes.byref(out)))
        return VarRef(out.value)

    def bitcast(self, ty: int) -> "VarRef":
        out = _u64()
        check(lib.hj_tr_bitcast(self._id, ty, ctypes.byref(out)))
        return VarRef(out.value)

    def fma(self, b, c) -> "VarRef":
        if not (isinstance(b, VarRef) and isinstance(c, VarRef)):
            ty = self.ty()
            b, c = _into(b, ty), _into(c, ty)
        out = _u64()
        check(lib.hj_tr_fma(self._id, b._id, c._id, ctypes.byref(out)))
        return VarRef(out.value)

    def select(self, condition, false_val) -> "VarRef":
        """``true_val.select(&cond, &false_val)`` (trace.rs:1483-1504)."""
        condition = _into(condition, BOOL)
        if not isinstance(false_val, VarRef):
            false_val = _into(false_val, self.ty())
        out = _u64()
        check(lib.hj_tr_select(self._id, condition._id, false_val._id, ctypes.byref(out)))
        return VarRef(out.value)

    def extract(self, elem: int) -> "VarRef":
        out = _u64()
        check(lib.hj_tr_extract(self._id, elem, ctypes.byref(out)))
        return VarRef(out.value)

    def extract_dyn(self, elem) -> "VarRef":
        elem = _into(elem, U32)
        out = _u64()
        check(lib.hj_tr_extract_dyn(self._id, elem._id, ctypes.byref(out)))
        return VarRef(out.value)

    # -- gather / scatter (trace.rs:1122-1334) -----------------------------------------------------
    def gather(self, idx) -> "VarRef":
        return self.gather_if(idx, True)

    def gather_if(self, idx, active) -> "VarRef":
        idx, active = _into(idx, U32), _into(active, BOOL)
        out = _u64()
        check(lib.hj_tr_gather(self._id, idx._id, active._id, ctypes.byref(out)))
        return VarRef(out.value)

    def scatter(self, dst, idx) -> None:
        check(lib.hj_tr_scatter(self._id, dst._id, _into(idx, U32)._id, 0))

    def scatter_if(self, dst, idx, active) -> None:
        check(lib.hj_tr_scatter(self._id, dst._id, _into(idx, U32)._id, _into(active, BOOL)._id))

    def scatter_reduce(self, dst, idx, op: int) -> None:
        check(lib.hj_tr_scatter_reduce(self._id, dst._id, _into(idx, U32)._id, 0, op))

    def scatter_reduce_if(self, dst, idx, active, op: int) -> None:
        check(lib.hj_tr_scatter_reduce(self._id, dst._id, _into(idx, U32)._id, _into(active, BOOL)._id, op))

    def scatter_atomic(self, dst, idx, op: int) -> "VarRef":
        out = _u64()
        check(lib.hj_tr_scatter_atomic(self._id, dst._id, _into(idx, U32)._id, 0, op, ctypes.byref(out)))
        return VarRef(out.value)

    def scatter_atomic_if(self, dst, idx, active, op: int) -> "VarRef":
        out = _u64()
        check(lib.hj_tr_scatter_atomic(self._id, dst._id, _into(idx, U32)._id, _into(active, BOOL)._id, op,
                                       ctypes.byref(out)))
        return VarRef(out.value)

    def atomic_inc(self, idx, active) -> "VarRef":
        out = _u64()
        check(lib.hj_tr_atomic_inc(self._id, _into(idx, U32)._id, _into(active, BOOL)._id, ctypes.byref(out)))
        return VarRef(out.value)

    # -- device ops (trace.rs:1583-1674) --------------------------------------------------------------
    def compress(self):
        """Returns ``(count, indices)`` (trace.rs:1595-1620)."""
        c, i = _u64(), _u64()
        check(lib.hj_tr_compress(self._id, ctypes.byref(c), ctypes.byref(i)))
        return VarRef(c.value), VarRef(i.value)

    def compress_dyn(self) -> "VarRef":
        out = _u64()
        check(lib.hj_tr_compress_dyn(self._id, ctypes.byref(out)))
        return VarRef(out.value)

    def prefix_sum(self, inclusive: bool) -> "VarRef":
        out = _u64()
        check(lib.hj_tr_prefix_sum(self._id, int(inclusive), ctypes.byref(out)))
        return VarRef(out.value)

    def reduce(self, op: int) -> "VarRef":
        out = _u64()
        check(lib.hj_tr_reduce(self._id, op, ctypes.byref(out)))
        return VarRef(out.value)

    def reduce_max(self): return self.reduce(MAX)
    def reduce_min(self): return self.reduce(MIN)
    def reduce_sum(self): return self.reduce(SUM)
    def reduce_prod(self): return self.reduce(PROD)
    def reduce_or(self): return self.reduce(OR)
    def reduce_and(self): return self.reduce(AND)
    def reduce_xor(self): return self.reduce(XOR)

    # -- read back (trace.rs:1400-1438) --------------------------------------------------------------
    def to_vec(self, dtype=None, start: int = 0, end: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """``to_vec::<T>(range)``: the evaluated values as ``dtype`` (default: the variable's own
        scalar type); a DynSize variable reads its device-resident count first.  ``out``: a contiguous
        destination of the right size (pinned memory makes the download asynchronous per chunk)."""
        ty = self.ty()
        n = _u64()
        check(lib.hj_tr_var_size(self._id, ctypes.byref(n)))
        size = n.value
        sh = self.shard()
        if sh is not None:  # a sharded variable reads this rank's block; start / end address the block
            size = sh[1]
        end = size if end is None else min(end, size)
        start = min(start, end)
        es = lib.hj_tr_type_size(ty)
        if dtype is None:
            dtype = _NP.get(ty, np.uint8)
        dt = np.dtype(dtype)
        nbytes = (end - start) * es
        assert nbytes % dt.itemsize == 0 and (start * es) % dt.itemsize == 0
        if out is not None:
            assert out.flags["C_CONTIGUOUS"] and out.nbytes == nbytes, "to_vec: `out` does not match the range"
            raw = out
        else:
            raw = np.empty(nbytes, dtype=np.uint8)
        if nbytes:
            check(lib.hj_tr_to_host(self._id, start, end - start, raw.ctypes.data_as(ctypes.c_void_p)))
        return raw if out is not None else raw.view(dt)

    def item(self, dtype=None):
        assert self.size() == 1
        return self.to_vec(dtype, 0, 1)[0]

    def shard(self):
        """None for an ordinary variable; ``(start, count, deferred)`` when the variable's buffer is the
        block ``[start, start + count)`` of its global extent on this rank (``hj_tr_var_shard``).
        ``deferred``: a scan result kept as (local scan, offset); ``to_vec`` adds the offset.  For a per-rank
        compacted segment (``is_segment``: the indices of a sharded ``compress``) ``count`` is the rank's own
        count, so ``to_vec`` returns the rank's part of the compacted sequence."""
        sharded, deferred, start, count = ctypes.c_int32(), ctypes.c_int32(), _u64(), _u64()
        check(lib.hj_tr_var_shard(self._id, ctypes.byref(sharded), ctypes.byref(start), ctypes.byref(count),
                                  ctypes.byref(deferred)))
        return (start.value, count.value, deferred.value == 1) if sharded.value else None

    def is_segment(self) -> bool:
        """True when the buffer is this rank's segment of a compacted sequence (``HJ_SHARD_SEGMENT``)."""
        sharded, deferred = ctypes.c_int32(), ctypes.c_int32()
        check(lib.hj_tr_var_shard(self._id, ctypes.byref(sharded), None, None, ctypes.byref(deferred)))
        return bool(sharded.value) and deferred.value in (2, 3)

    def materialise(self) -> None:
        """Adds the offset of a deferred scan result on the device (``hj_tr_materialise``)."""
        check(lib.hj_tr_materialise(self._id))


def _into(value, ty_hint: int) -> VarRef:
    """``impl<T: AsVarType> From<T> for VarRef`` (trace.rs:617-621): plain values become literals."""
    if isinstance(value, VarRef):
        return value
    if isinstance(value, (bool, np.bool_)):
        return literal(bool(value), BOOL)
    return literal(value, ty_hint)


# ---- constructors (trace.rs:547-663) ---------------------------------------------------------------
def index() -> VarRef:
    out = _u64()
    check(lib.hj_tr_index(ctypes.byref(out)))
    return VarRef(out.value)


def sized_index(size: int) -> VarRef:
    out = _u64()
    check(lib.hj_tr_sized_index(size, ctypes.byref(out)))
    return VarRef(out.value)


def dynamic_index(capacity: int, size: VarRef) -> VarRef:
    out = _u64()
    check(lib.hj_tr_dynamic_index(capacity, size._id, ctypes.byref(out)))
    return VarRef(out.value)


def literal(value, ty: int | None = None) -> VarRef:
    ty = _infer_type(value) if ty is None else ty
    out = _u64()
    check(lib.hj_tr_literal(ty, _bits(value, ty), ctypes.byref(out)))
    return VarRef(out.value)


def sized_literal(value, size: int, ty: int | None = None) -> VarRef:
    ty = _infer_type(value) if ty is None else ty
    out = _u64()
    check(lib.hj_tr_sized_literal(ty, _bits(value, ty), size, ctypes.byref(out)))
    return VarRef(out.value)


def array(data, device: Device, ty: int | None = None) -> VarRef:
    """``tr::array(&slice, &device)`` (trace.rs:647-663): uploads immediately."""
    a = np.ascontiguousarray(data)
    if a.dtype == np.int64 and ty is None and not isinstance(data, np.ndarray):
        a = a.astype(np.int32)  # Rust integer literals default to i32
    if a.dtype == np.float64 and ty is None and not isinstance(data, np.ndarray):
        a = a.astype(np.float32)
    if ty is None:
        ty = _FROM_NP[a.dtype]
    else:
        a = a.astype(_NP[ty])
    out = _u64()
    check(lib.hj_tr_array(device.handle, ty, a.ctypes.data_as(ctypes.c_void_p), a.size, ctypes.byref(out)))
    return VarRef(out.value)


def array_async(data: np.ndarray, device: Device, ty: int | None = None) -> VarRef:
    """``tr::array`` that returns at once (``hj_tr_array_async``): ``data`` — a contiguous numpy array over
    PINNED memory — is uploaded chunk by chunk on a side stream; a launch of one fused kernel over it and the
    ``to_vec`` of its results then run chunk-wise behind the upload, so the three steps overlap.  The array is
    kept alive by the variable and must not be modified until a dependent ``to_vec`` / ``device.sync()`` returned."""
    assert isinstance(data, np.ndarray) and data.flags["C_CONTIGUOUS"], "array_async takes a contiguous numpy array"
    if ty is None:
        ty = _FROM_NP[data.dtype]
    assert np.dtype(_NP[ty]).itemsize == data.dtype.itemsize
    out = _u64()
    check(lib.hj_tr_array_async(device.handle, ty, data.ctypes.data_as(ctypes.c_void_p), data.size, ctypes.byref(out)))
    v = VarRef(out.value)
    v._keepalive = data
    return v


def array_sharded(data, comm, ty: int | None = None, is_global: bool = True) -> VarRef:
    """A variable of ``n`` elements partitioned over the ranks of ``comm`` (``sharded.Comm``): every rank
    uploads its contiguous block.  ``data`` is the GLOBAL array (each rank slices its block out of it —
    convenient for tests) or, with ``is_global=False``, ``(local_block, n_global)``."""
    if is_global:
        a = np.ascontiguousarray(data)
        n_global = a.size
        lo, hi = comm.bounds(n_global)
        a = np.ascontiguousarray(a[lo:hi])
    else:
        a, n_global = np.ascontiguousarray(data[0]), int(data[1])
        lo, hi = comm.bounds(n_global)
        assert a.size == hi - lo, "the local block does not match hj_shard_bounds"
    if ty is None:
        ty = _FROM_NP[a.dtype]
    else:
        a = a.astype(_NP[ty])
    out = _u64()
    check(lib.hj_tr_array_sharded(comm.handle, ty, a.ctypes.data_as(ctypes.c_void_p), n_global, ctypes.byref(out)))
    return VarRef(out.value)


def from_buffer_sharded(buf: Buffer, ty: int, n_global: int, comm) -> VarRef:
    out = _u64()
    check(lib.hj_tr_from_buffer_sharded(comm.handle, buf.handle, ty, n_global, ctypes.byref(out)))
    return VarRef(out.value)


def from_buffer(buf: Buffer, ty: int, size: int) -> VarRef:
    out = _u64()
    check(lib.hj_tr_from_buffer(buf.handle, ty, size, ctypes.byref(out)))
    return VarRef(out.value)


def _handles(refs):
    return (ctypes.c_uint64 * max(len(refs), 1))(*[r._id for r in refs])


def composite(refs) -> VarRef:
    out = _u64()
    check(lib.hj_tr_composite(_handles(refs), len(refs), ctypes.byref(out)))
    return VarRef(out.value)


def vec(refs) -> VarRef:
    out = _u64()
    check(lib.hj_tr_vec(_handles(refs), len(refs), ctypes.byref(out)))
    return VarRef(out.value)


def mat(columns) -> VarRef:
    """``tr::mat`` (trace.rs:734-756): a matrix from its column vectors."""
    out = _u64()
    check(lib.hj_tr_mat(_handles(columns), len(columns), ctypes.byref(out)))
    return VarRef(out.value)


def arr(refs) -> VarRef:
    out = _u64()
    check(lib.hj_tr_arr(_handles(refs), len(refs), ctypes.byref(out)))
    return VarRef(out.value)


# ---- recorded control flow (trace.rs:408-521) ---------------------------------------------------------
def _scope_start(is_loop: bool, state_vars):
    n = len(state_vars)
    scope = _u64()
    out = (ctypes.c_uint64 * n)()
    check(lib.hj_tr_scope_start(int(is_loop), _handles(state_vars), n, ctypes.byref(scope), out))
    return VarRef(scope.value), [VarRef(out[i]) for i in range(n)]


def _scope_end(scope: VarRef, state_vars):
    n = len(state_vars)
    out = (ctypes.c_uint64 * n)()
    check(lib.hj_tr_scope_end(scope._id, _handles(state_vars), n, out))
    return [VarRef(out[i]) for i in range(n)]


def loop_start(state_vars):
    """``tr::loop_start(&[cond, vars...])`` -> (loop_start, state)."""
    return _scope_start(True, state_vars)


def loop_end(loop_start_var: VarRef, state_vars):
    return _scope_end(loop_start_var, state_vars)


def if_start(state_vars):
    return _scope_start(False, state_vars)


def if_end(if_start_var: VarRef, state_vars):
    return _scope_end(if_start_var, state_vars)


def loop_record(cond: VarRef, vars_, body):
    """``loop_record!([vars] while cond { body })`` (record.rs:56-73): ``body(cond, vars) -> (cond, vars)``."""
    start, state = loop_start([cond] + list(vars_))
    cond, vars_ = body(state[0], state[1:])
    state = loop_end(start, [cond] + list(vars_))
    return state[0], state[1:]


def if_record(cond: VarRef, vars_, body):
    """``if_record!([vars] if cond { body })`` (record.rs:74-91)."""
    start, state = if_start([cond] + list(vars_))
    cond, vars_ = body(state[0], state[1:])
    state = if_end(start, [cond] + list(vars_))
    return state[0], state[1:]


# ---- schedule / graph (trace.rs:528-545, graph.rs) ---------------------------------------------------------
def schedule_eval() -> None:
    check(lib.hj_tr_schedule_eval())


def is_empty() -> bool:
    return bool(lib.hj_tr_is_empty())


def n_live() -> int:
    return lib.hj_tr_n_live()


class Report:
    """``graph::Report`` (graph.rs:138-143) + per-pass GPU times (backend/report.rs:2-19)."""

    def __init__(self, aliasing_rate, aliasing_duration_us, backend_cpu_us, passes):
        self.aliasing_rate = aliasing_rate
        self.aliasing_duration_us = aliasing_duration_us
        self.backend_cpu_us = backend_cpu_us
        self.passes = passes  # list of (name, start_us, duration_us) or None


class Graph:
    """``graph::Graph`` (graph.rs:145-151)."""

    def __init__(self, handle: int):
        self._h = ctypes.c_void_p(handle)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:  # module globals are torn down before objects at interpreter exit
            lib.hj_graph_release(h)

    def n_passes(self) -> int:
        return lib.hj_graph_n_passes(self._h)

    def n_outputs(self) -> int:
        return lib.hj_graph_n_outputs(self._h)

    def debug_string(self) -> str:
        out = ctypes.c_void_p()
        check(lib.hj_graph_debug_string(self._h, ctypes.byref(out)))
        s = ctypes.string_at(out).decode()
        lib.hj_free_string(out)
        return s

    def pass_ir(self, index: int):
        """The flat IR (``_lib.Ir``) of kernel pass ``index``, or None for a device-op pass; valid while
        this Graph is alive (``ir.codegen`` / ``ir.compile_cubin`` accept it)."""
        out = ctypes.POINTER(_lib.Ir)()
        check(lib.hj_graph_pass_ir(self._h, index, ctypes.byref(out)))
        return out.contents if out else None

    def serialize(self) -> bytes:
        """Wire format of the graph (passes + kernel IR + resources + captured buffer contents);
        the reference keeps graphs in memory only (graph.rs:145-151)."""
        out, n = ctypes.c_void_p(), ctypes.c_size_t()
        check(lib.hj_graph_serialize(self._h, ctypes.byref(out), ctypes.byref(n)))
        data = ctypes.string_at(out, n.value)
        lib.hj_free_string(out)
        return data

    @staticmethod
    def deserialize(data: bytes, device: "Device | None" = None) -> "Graph":
        """Rebuild a graph in this process; `device` is needed when it captured buffers."""
        out = ctypes.c_void_p()
        check(lib.hj_graph_deserialize(device.handle if device is not None else None, data, len(data), ctypes.byref(out)))
        return Graph(out.value)

    def launch(self, device: Device, timed: bool = False, comm=None) -> Report:
        return self.launch_with(device, [], timed, comm)[0]

    def launch_with(self, device: Device, inputs, timed: bool = False, comm=None):
        """``Graph::launch_with(device, inputs)`` -> (Report, outputs).  ``comm`` (``sharded.Comm``): run over
        arrays partitioned across its ranks; a graph that meets a sharded variable does so by itself."""
        n_out = lib.hj_graph_n_outputs(self._h)
        outs = (ctypes.c_uint64 * max(n_out, 1))()
        rep = _lib.GraphReport()
        n_p = self.n_passes()
        pr = (_lib.PassReport * max(n_p, 1))()
        if timed:
            rep.passes = pr
            rep.passes_capacity = n_p
        check(lib.hj_graph_launch_sharded(self._h, device.handle, comm.handle if comm is not None else None,
                                          _handles(inputs), len(inputs), outs, ctypes.byref(rep)))
        passes = [(pr[i].name.decode(), pr[i].start_us, pr[i].duration_us) for i in range(n_p)] if timed else None
        report = Report(rep.aliasing_rate, rep.aliasing_duration_us, rep.backend_cpu_us, passes)
        return report, [VarRef(outs[i]) for i in range(n_out)]


def compile() -> Graph:  # noqa: A001 - the reference's name
    """``tr::compile()`` (trace.rs:528-536)."""
    out = ctypes.c_void_p()
    check(lib.hj_tr_compile(ctypes.byref(out)))
    return Graph(out.value)


def compile_fn(inputs, outputs) -> Graph:
    out = ctypes.c_void_p()
    check(lib.hj_tr_compile_fn(_handles(inputs), len(inputs), _handles(outputs), len(outputs), ctypes.byref(out)))
    return Graph(out.value)
