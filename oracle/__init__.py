"""CPU oracle for the hephaestus-jit hot path (TEST INFRASTRUCTURE — see hj_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` leg
may import this package, and only as the checker or the timed CPU baseline.  The product
(hephaestus-jit_b200/) never imports it.

Two halves:
  * ``hj_oracle.c`` (ctypes, this file): the device ops — reduce / prefix_sum / compress /
    scatter_reduce / gather — restated step for step from the reference's GLSL kernels.
  * ``oracle.ir_interp`` (numpy): an interpreter for the fused-kernel IR with the per-op
    meaning of the reference's GLSL code generator.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhj_oracle.so")

# hj_type_kind (include/hj.h; order of VarType in vartype.rs:89-104)
VOID, BOOL, I8, U8, I16, U16, I32, U32, I64, U64, F16, F32, F64 = range(13)
# hj_reduce_op (order of ReduceOp in op.rs:90-99)
MAX, MIN, SUM, PROD, OR, AND, XOR = range(7)

NP_DTYPE = {
    BOOL: np.uint8, I8: np.int8, U8: np.uint8, I16: np.int16, U16: np.uint16,
    I32: np.int32, U32: np.uint32, I64: np.int64, U64: np.uint64,
    F16: np.float16, F32: np.float32, F64: np.float64,
}
TYPE_NAME = {
    VOID: "Void", BOOL: "Bool", I8: "I8", U8: "U8", I16: "I16", U16: "U16", I32: "I32",
    U32: "U32", I64: "I64", U64: "U64", F16: "F16", F32: "F32", F64: "F64",
}
OP_NAME = {MAX: "max", MIN: "min", SUM: "sum", PROD: "prod", OR: "or", AND: "and", XOR: "xor"}


def build(force: bool = False) -> str:
    """Compile libhj_oracle.so with the committed Makefile if it is missing or stale."""
    src = os.path.join(_HERE, "hj_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "libhj_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        sz, vp, i32, u32, u64 = (ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int,
                                 ctypes.c_uint32, ctypes.c_uint64)
        L.hjo_set_threads.argtypes = [i32]
        L.hjo_get_threads.restype = i32
        L.hjo_type_size.argtypes = [i32]
        L.hjo_type_size.restype = sz
        L.hjo_reduce_supported.argtypes = [i32, i32]
        L.hjo_reduce.argtypes = [i32, i32, sz, vp, vp, i32]
        L.hjo_prefix_sum.argtypes = [i32, sz, i32, i32, vp, vp]
        L.hjo_prefix_sum_u32_mt.argtypes = [sz, i32, vp, vp]
        L.hjo_compress.argtypes = [sz, vp, vp, vp, u32]
        L.hjo_compress_mt.argtypes = [sz, vp, vp, vp, u32]
        L.hjo_scatter_reduce.argtypes = [i32, i32, sz, vp, vp, u64, vp, sz]
        L.hjo_histogram_u32_mt.argtypes = [sz, vp, vp, sz]
        L.hjo_gather.argtypes = [sz, sz, vp, sz, vp, vp]
        L.hjo_c2_chain_f32.argtypes = [sz, vp, vp]
        L.hjo_c2_chain_f32_fast.argtypes = [sz, vp, vp]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _check(rc: int, what: str) -> None:
    if rc == -2:
        raise NotImplementedError(f"{what}: unsupported by the reference (todo!())")
    if rc != 0:
        raise OracleError(f"{what}: oracle returned {rc}")


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def set_threads(n: int) -> None:
    lib().hjo_set_threads(int(n))


def get_threads() -> int:
    return int(lib().hjo_get_threads())


def reduce_supported(op: int, ty: int) -> bool:
    return bool(lib().hjo_reduce_supported(op, ty))


def reduce(op: int, ty: int, src: np.ndarray, faithful_copy: bool = False) -> np.ndarray:
    """Radix-32 tree reduction of ``src`` (reduce.rs:22-314 / reduce.glsl). Returns shape (1,)."""
    dt = NP_DTYPE[ty]
    src = np.ascontiguousarray(src).view(dt) if src.dtype != np.bool_ else \
        np.ascontiguousarray(src).view(np.uint8)
    out = np.zeros(1, dtype=dt)
    _check(lib().hjo_reduce(op, ty, src.size, _ptr(src), _ptr(out), int(faithful_copy)),
           f"reduce({OP_NAME[op]}, {TYPE_NAME[ty]})")
    return out


def prefix_sum(ty: int, src: np.ndarray, inclusive: bool, ref_compat: bool = False) -> np.ndarray:
    """Partitioned decoupled-look-back scan (prefix_sum_large.glsl) — serialised schedule."""
    dt = NP_DTYPE[ty]
    src = np.ascontiguousarray(src).view(dt)
    out = np.empty_like(src)
    _check(lib().hjo_prefix_sum(ty, src.size, int(inclusive), int(ref_compat), _ptr(src),
                                _ptr(out)), f"prefix_sum({TYPE_NAME[ty]})")
    return out


def prefix_sum_u32_mt(src: np.ndarray, inclusive: bool) -> np.ndarray:
    src = np.ascontiguousarray(src, dtype=np.uint32)
    out = np.empty_like(src)
    _check(lib().hjo_prefix_sum_u32_mt(src.size, int(inclusive), _ptr(src), _ptr(out)),
           "prefix_sum_u32_mt")
    return out


def compress(mask: np.ndarray, index_out: np.ndarray | None = None, index_base: int = 0,
             mt: bool = False):
    """Returns (count, index_out). ``index_out`` entries >= count are left untouched, as in the
    reference; when not supplied it starts zeroed (the scheduler's zero-fill pass)."""
    mask = np.ascontiguousarray(mask).view(np.uint8)
    if index_out is None:
        index_out = np.zeros(mask.size, dtype=np.uint32)
    count = np.zeros(1, dtype=np.uint32)
    fn = lib().hjo_compress_mt if mt else lib().hjo_compress
    _check(fn(mask.size, _ptr(mask), _ptr(index_out), _ptr(count), index_base), "compress")
    return int(count[0]), index_out


def scatter_reduce(op: int, ty: int, idx: np.ndarray, src, dst: np.ndarray) -> np.ndarray:
    """dst[idx[i]] = op(dst[idx[i]], src[i] or literal) in place; ``src`` array or scalar."""
    dt = NP_DTYPE[ty]
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    assert dst.dtype == dt and dst.flags.c_contiguous
    if isinstance(src, np.ndarray):
        src = np.ascontiguousarray(src, dtype=dt)
        sp, lit = _ptr(src), 0
    else:
        sp = None
        lit = int(np.array([src], dtype=dt).view(
            {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[np.dtype(dt).itemsize])[0])
    _check(lib().hjo_scatter_reduce(op, ty, idx.size, _ptr(idx), sp, lit, _ptr(dst), dst.size),
           "scatter_reduce")
    return dst


def histogram_u32_mt(idx: np.ndarray, n_bins: int) -> np.ndarray:
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    dst = np.zeros(n_bins, dtype=np.uint32)
    _check(lib().hjo_histogram_u32_mt(idx.size, _ptr(idx), _ptr(dst), n_bins), "histogram")
    return dst


def gather(src: np.ndarray, idx: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src)
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    out = np.empty(idx.size, dtype=src.dtype)
    _check(lib().hjo_gather(src.dtype.itemsize, idx.size, _ptr(src), src.size, _ptr(idx),
                            _ptr(out)), "gather")
    return out


def c2_chain(x: np.ndarray, fast: bool = False) -> np.ndarray:
    """BASELINE C2: y = select(x > 0, sin(fma(x,1.5,0.25)), exp2(fma(x,1.5,0.25)))."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    fn = lib().hjo_c2_chain_f32_fast if fast else lib().hjo_c2_chain_f32
    _check(fn(x.size, _ptr(x), _ptr(y)), "c2_chain")
    return y
