"""Builder for the flat fused-kernel IR (``hj_ir`` in include/hj.h; mirror of
hephaestus-jit/src/ir.rs:12-46 and the interned VarType tree, vartype.rs:89-122).

Used by the tests and by bench.py to hand hand-built IR straight to the backend boundary
(``hj_kernel_get`` / ``hj_execute_graph``) — the same structures the host-side trace compiler
(csrc/trace.cpp) produces.
"""
from __future__ import annotations

import ctypes
import struct

import numpy as np

from . import _lib
from ._lib import check, lib

# hj_type_kind
VOID, BOOL, I8, U8, I16, U16, I32, U32, I64, U64, F16, F32, F64, VEC, ARRAY, MAT, STRUCT = range(17)
# hj_kernel_op
(OP_NOP, OP_SCATTER, OP_SCATTER_REDUCE, OP_SCATTER_ATOMIC, OP_ATOMIC_INC, OP_GATHER, OP_INDEX,
 OP_LITERAL, OP_EXTRACT, OP_DYN_EXTRACT, OP_CONSTRUCT, OP_SELECT, OP_LOOP_START, OP_LOOP_END,
 OP_IF_START, OP_IF_END, OP_TEX_LOOKUP, OP_TRACE_RAY, OP_BOP, OP_UOP, OP_FMA, OP_BUFFER_REF,
 OP_TEXTURE_REF, OP_ACCEL_REF) = range(24)
# hj_bop / hj_uop
(BOP_ADD, BOP_SUB, BOP_MUL, BOP_DIV, BOP_MODULUS, BOP_MIN, BOP_MAX, BOP_INNER, BOP_AND, BOP_OR,
 BOP_XOR, BOP_SHL, BOP_SHR, BOP_EQ, BOP_NEQ, BOP_LT, BOP_LE, BOP_GT, BOP_GE) = range(19)
(UOP_CAST, UOP_BITCAST, UOP_NEG, UOP_SQRT, UOP_ABS, UOP_SIN, UOP_COS, UOP_EXP2,
 UOP_LOG2) = range(9)

_PACK = {BOOL: "<B", I8: "<b", U8: "<B", I16: "<h", U16: "<H", I32: "<i", U32: "<I", I64: "<q",
         U64: "<Q", F16: "<e", F32: "<f", F64: "<d"}


def literal_bits(kind: int, value) -> int:
    """The u64 the reference stores for a literal: the value's bytes in the low bits
    (trace.rs:602-606)."""
    raw = struct.pack(_PACK[kind], bool(value) if kind == BOOL else value)
    return int.from_bytes(raw, "little")


class IRBuilder:
    def __init__(self):
        self.types: list[tuple] = []          # (kind, elem, num, cols, rows, first_field)
        self._type_ids: dict[tuple, int] = {}
        self.struct_fields: list[int] = []
        self.vars: list[tuple] = []           # (ty, op, arg, dep_start, dep_end, data)
        self.deps: list[int] = []
        self.n_buffers = 0
        self._keep = None

    # ---- types (interned like vartype.rs:19-85) ------------------------------------------
    def _intern(self, key, desc) -> int:
        if key not in self._type_ids:
            self._type_ids[key] = len(self.types)
            self.types.append(desc)
        return self._type_ids[key]

    def scalar(self, kind: int) -> int:
        return self._intern(("s", kind), (kind, 0, 0, 0, 0, 0))

    def vec(self, elem: int, num: int) -> int:
        return self._intern(("v", elem, num), (VEC, elem, num, 0, 0, 0))

    def array(self, elem: int, num: int) -> int:
        return self._intern(("a", elem, num), (ARRAY, elem, num, 0, 0, 0))

    def mat(self, elem: int, cols: int, rows: int) -> int:
        return self._intern(("m", elem, cols, rows), (MAT, elem, 0, cols, rows, 0))

    def struct(self, fields: list[int]) -> int:
        key = ("st", tuple(fields))
        if key not in self._type_ids:
            first = len(self.struct_fields)
            self.struct_fields.extend(fields)
            self._type_ids[key] = len(self.types)
            self.types.append((STRUCT, 0, len(fields), 0, 0, first))
        return self._type_ids[key]

    # ---- vars -----------------------------------------------------------------------------
    def push(self, op: int, ty: int, deps=(), data: int = 0, arg: int = 0) -> int:
        start = len(self.deps)
        self.deps.extend(int(d) for d in deps)
        self.vars.append((ty, op, arg, start, len(self.deps), data))
        return len(self.vars) - 1

    def buffer_ref(self, ty: int, slot: int | None = None) -> int:
        if slot is None:
            slot = self.n_buffers
        self.n_buffers = max(self.n_buffers, slot + 1)
        return self.push(OP_BUFFER_REF, ty, data=slot)

    def index(self) -> int:
        return self.push(OP_INDEX, self.scalar(U32))

    def literal(self, kind: int, value) -> int:
        return self.push(OP_LITERAL, self.scalar(kind), data=literal_bits(kind, value))

    def gather(self, ty: int, buf: int, idx: int, cond: int | None = None) -> int:
        return self.push(OP_GATHER, ty, [buf, idx] + ([cond] if cond is not None else []))

    def scatter(self, buf: int, src: int, idx: int, cond: int | None = None) -> int:
        return self.push(OP_SCATTER, self.scalar(VOID), [buf, src, idx] + ([cond] if cond is not None else []))

    def scatter_reduce(self, rop: int, buf: int, src: int, idx: int, cond: int | None = None) -> int:
        return self.push(OP_SCATTER_REDUCE, self.scalar(VOID),
                         [buf, src, idx] + ([cond] if cond is not None else []), arg=rop)

    def bop(self, bop: int, ty: int, a: int, b: int) -> int:
        return self.push(OP_BOP, ty, [a, b], arg=bop)

    def uop(self, uop: int, ty: int, a: int) -> int:
        return self.push(OP_UOP, ty, [a], arg=uop)

    def fma(self, ty: int, a: int, b: int, c: int) -> int:
        return self.push(OP_FMA, ty, [a, b, c])

    def select(self, ty: int, cond: int, t: int, f: int) -> int:
        return self.push(OP_SELECT, ty, [cond, t, f])

    # ---- C view ---------------------------------------------------------------------------
    def build(self) -> _lib.Ir:
        nv, nd, nt, nf = len(self.vars), len(self.deps), len(self.types), len(self.struct_fields)
        vars_ = (_lib.IrVar * max(nv, 1))()
        for i, (ty, op, arg, ds, de, data) in enumerate(self.vars):
            vars_[i] = _lib.IrVar(ty, op, arg, ds, de, 0, data)
        deps = (ctypes.c_uint32 * max(nd, 1))(*self.deps)
        types = (_lib.TypeDesc * max(nt, 1))()
        for i, t in enumerate(self.types):
            types[i] = _lib.TypeDesc(*t)
        fields = (ctypes.c_uint32 * max(nf, 1))(*self.struct_fields)
        ir = _lib.Ir(vars_, nv, deps, nd, types, nt, fields, nf, self.n_buffers)
        self._keep = (vars_, deps, types, fields)  # keep the arrays alive with the builder
        return ir


def ir_hash(ir: _lib.Ir) -> int:
    return int(lib.hj_ir_hash(ctypes.byref(ir)))


def codegen(ir: _lib.Ir) -> str:
    out = ctypes.c_void_p()
    check(lib.hj_ir_codegen(ctypes.byref(ir), ctypes.byref(out)))
    try:
        return ctypes.string_at(out).decode()
    finally:
        lib.hj_free_string(out)


def compile_cubin(ir: _lib.Ir) -> bytes:
    """NVRTC-compile to an sm_100a cubin; works without a GPU."""
    out, size = ctypes.c_void_p(), ctypes.c_size_t()
    check(lib.hj_ir_compile_cubin(ctypes.byref(ir), ctypes.byref(out), ctypes.byref(size)))
    try:
        return ctypes.string_at(out, size.value)
    finally:
        lib.hj_free_string(out)


def c2_chain_ir() -> IRBuilder:
    """BASELINE.json config C2: t = fma(x, 1.5, 0.25);
    y = select(x > 0, sin(t), exp2(t)) — the IR `Compiler::compile` (compiler.rs:20-65) emits
    for that trace: one input buffer, one output buffer, one kernel."""
    b = IRBuilder()
    f32, u32, boolt = b.scalar(F32), b.scalar(U32), b.scalar(BOOL)
    x_ref = b.buffer_ref(f32)
    idx = b.index()
    x = b.gather(f32, x_ref, idx)
    zero = b.literal(F32, 0.0)
    cond = b.bop(BOP_GT, boolt, x, zero)
    t = b.fma(f32, x, b.literal(F32, 1.5), b.literal(F32, 0.25))
    s = b.uop(UOP_SIN, f32, t)
    e = b.uop(UOP_EXP2, f32, t)
    y = b.select(f32, cond, s, e)
    y_ref = b.buffer_ref(f32)
    b.scatter(y_ref, y, idx)
    return b


def wavefront_step_passes(n: int, index_as_value: bool = False, conditional: bool = False, threshold: float = 0.1):
    """The wavefront step of the reference's `example` (jit/test.rs:1020-1062) as the pass list Graph::compile
    produces for it: Compress(index, count, mask), then two DynSize kernels sized by `count` —
    ``mask[idx] = a[idx] * 0.9 > threshold`` and ``a[idx] = a[idx] * 0.9`` with ``idx = index[Index]``.
    Resources: 0 a (f32), 1 mask (bool), 2 index (u32), 3 count (u32, one element).  ``index_as_value``: the
    lanes take their position in the compacted sequence instead; ``conditional``: the index is read under a
    condition (what a sharded launch must refuse).  Returns ``(passes, descs)``."""
    from . import PASS_COMPRESS, PASS_KERNEL

    def step(write_mask: bool) -> IRBuilder:
        b = IRBuilder()
        f32, u32, bl = b.scalar(F32), b.scalar(U32), b.scalar(BOOL)
        ra, rindex = b.buffer_ref(f32), b.buffer_ref(u32)
        i = b.index()
        active = b.bop(BOP_LT, bl, i, b.literal(U32, 7)) if conditional else b.literal(BOOL, 1)
        idx = b.gather(u32, rindex, i, active)
        v = b.bop(BOP_MUL, f32, b.gather(f32, ra, idx), b.literal(F32, 0.9))
        if index_as_value:
            v = b.uop(UOP_CAST, f32, i)
        if write_mask:
            b.scatter(b.buffer_ref(bl), b.bop(BOP_GT, bl, v, b.literal(F32, threshold)), idx)
        else:
            b.scatter(ra, v, idx)
        return b

    passes = [
        {"kind": PASS_COMPRESS, "resources": [2, 3, 1]},
        {"kind": PASS_KERNEL, "resources": [0, 2, 1], "ir": step(True), "size": n, "size_buffer": 3},
        {"kind": PASS_KERNEL, "resources": [0, 2], "ir": step(False), "size": n, "size_buffer": 3},
    ]
    descs = [(n, F32, 4), (n, BOOL, 1), (n, U32, 4), (1, U32, 4)]
    return passes, descs
