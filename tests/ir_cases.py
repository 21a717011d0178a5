"""Library of fused-kernel IR programs shared by the CPU tests (oracle interpreter + codegen/NVRTC
compile) and the GPU parity tests (NVRTC kernel vs interpreter).

Each case is a function returning ``Case(builder, size, buffers, expect, exact, ...)``:
  builder  — hephaestus-jit_b200.ir.IRBuilder holding the IR the trace compiler
             (hephaestus-jit/src/compiler.rs:20-230) emits for the traced program;
  buffers  — numpy arrays in memory layout, one per IR buffer slot (inputs AND outputs);
  expect   — {slot: expected array} taken from the reference's tests where they state one.
"""
from __future__ import annotations

import importlib
from dataclasses import dataclass, field

import numpy as np

irm = importlib.import_module("hephaestus-jit_b200.ir")
from oracle import ir_interp as I  # noqa: E402  (op / type constants are identical)

NP = I.NP


@dataclass
class Case:
    name: str
    builder: object
    size: int
    buffers: list
    expect: dict = field(default_factory=dict)   # slot -> exact expected contents (reference KAT)
    exact: bool = True                           # GPU == interpreter bit for bit
    rtol: float = 0.0
    atol: float = 0.0
    size_buf: object = None
    index_base: int = 0
    unordered_unique: tuple = ()                 # slots whose values only need to be unique
    check_slots: tuple = None                    # slots to compare (default: all)


def _out(b, kind_or_ty, src, idx, is_ty=False):
    ty = kind_or_ty if is_ty else b.scalar(kind_or_ty)
    ref = b.buffer_ref(ty)
    b.scatter(ref, src, idx)
    return ref


# ---- reference programs (hephaestus-jit/src/test.rs) -------------------------------------------

def select():
    """test.rs:393-409 + snapshot hephaestus_jit__test__select.snap."""
    b = irm.IRBuilder()
    boolt, u32, i32 = b.scalar(I.BOOL), b.scalar(I.U32), b.scalar(I.I32)
    v0 = b.buffer_ref(boolt)
    v1 = b.index()
    v2 = b.gather(boolt, v0, v1)
    v3 = b.literal(I.I32, 10)
    v4 = b.literal(I.I32, 5)
    v5 = b.select(i32, v2, v3, v4)
    v6 = b.buffer_ref(i32)
    b.scatter(v6, v5, v1)
    return Case("select", b, 2, [np.array([1, 0], np.uint8), np.zeros(2, np.int32)],
                expect={1: np.array([10, 5], np.int32)})


def conditional_scatter():
    """test.rs:348-372, second pass of the snapshot: Scatter(dst, 1, index, active)."""
    b = irm.IRBuilder()
    i32, boolt = b.scalar(I.I32), b.scalar(I.BOOL)
    v0 = b.buffer_ref(i32)
    v1 = b.literal(I.I32, 1)
    v2 = b.index()
    v3 = b.buffer_ref(boolt)
    v4 = b.gather(boolt, v3, v2)
    b.scatter(v0, v1, v2, v4)
    active = np.array([1, 1, 0, 0, 1, 0, 1, 0, 1, 0], np.uint8)
    return Case("conditional_scatter", b, 10, [np.zeros(10, np.int32), active],
                expect={0: active.astype(np.int32)})


def conditional_gather():
    """test.rs:373-392: dst = src.gather_if(index, active); src is a buffer of ones."""
    b = irm.IRBuilder()
    i32, boolt = b.scalar(I.I32), b.scalar(I.BOOL)
    src = b.buffer_ref(i32)
    idx = b.index()
    act_ref = b.buffer_ref(boolt)
    act = b.gather(boolt, act_ref, idx)
    g = b.gather(i32, src, idx, act)
    _out(b, I.I32, g, idx)
    active = np.array([1, 1, 0, 0, 1, 0, 1, 0, 1, 0], np.uint8)
    return Case("conditional_gather", b, 10, [np.ones(10, np.int32), active, np.full(10, -7, np.int32)],
                expect={2: active.astype(np.int32)})


def simple1_second_pass():
    """test.rs:62-92: (j + 1).scatter(i, j) over 5 lanes into the 10-element index buffer i."""
    b = irm.IRBuilder()
    u32 = b.scalar(I.U32)
    i_ref = b.buffer_ref(u32)
    j = b.index()
    one = b.literal(I.U32, 1)
    s = b.bop(I.BOP_ADD, u32, j, one)
    b.scatter(i_ref, s, j)
    _out(b, I.U32, j, j)
    return Case("simple1", b, 5, [np.arange(10, dtype=np.uint32), np.zeros(5, np.uint32)],
                expect={0: np.array([1, 2, 3, 4, 5, 5, 6, 7, 8, 9], np.uint32),
                        1: np.arange(5, dtype=np.uint32)})


def literal_fill(kind=I.U16, value=1, n=10):
    """test.rs:94-115 (simple_u16), :323-337 (conditionals: sized_literal(true, 100))."""
    b = irm.IRBuilder()
    lit = b.literal(kind, value)
    ref = b.buffer_ref(b.scalar(kind))
    b.scatter(ref, lit, b.index())
    mem = I.MEM[kind]
    return Case(f"literal_fill_{kind}", b, n, [np.zeros(n, mem)], expect={0: np.full(n, value, mem)})


def simple_f16():
    """test.rs:116-128: sized_index(10).cast(f16)."""
    b = irm.IRBuilder()
    idx = b.index()
    c = b.uop(I.UOP_CAST, b.scalar(I.F16), idx)
    _out(b, I.F16, c, idx)
    return Case("simple_f16", b, 10, [np.zeros(10, np.float16)], expect={0: np.arange(10, dtype=np.float16)})


def scatter_chain_add():
    """test.rs:130-154 last pass: b1 = b0 + 1 where b0 was scattered to 1."""
    b = irm.IRBuilder()
    i32 = b.scalar(I.I32)
    b0 = b.buffer_ref(i32)
    idx = b.index()
    x = b.gather(i32, b0, idx)
    y = b.bop(I.BOP_ADD, i32, x, b.literal(I.I32, 1))
    _out(b, I.I32, y, idx)
    return Case("scatter_chain1", b, 5, [np.ones(5, np.int32), np.zeros(5, np.int32)],
                expect={1: np.full(5, 2, np.int32)})


def uop_cos():
    """test.rs:804-828: cos of [0, 1, pi], abs eps 1e-3."""
    b = irm.IRBuilder()
    f32 = b.scalar(I.F32)
    ref = b.buffer_ref(f32)
    idx = b.index()
    x = b.gather(f32, ref, idx)
    _out(b, I.F32, b.uop(I.UOP_COS, f32, x), idx)
    x_host = np.array([0.0, 1.0, np.pi], np.float32)
    return Case("uop_cos", b, 3, [x_host, np.zeros(3, np.float32)], exact=False, atol=1e-3,
                expect={1: np.cos(x_host.astype(np.float64)).astype(np.float32)})


def scatter_reduce_kat():
    """test.rs:863-883: 16 lanes add 1u32 into bin 0 of [0, 0, 0]."""
    b = irm.IRBuilder()
    u32 = b.scalar(I.U32)
    dst = b.buffer_ref(u32)
    src = b.literal(I.U32, 1)
    idx = b.literal(I.I32, 0)  # tr::sized_literal(0, n) is an i32 literal in the reference test
    b.scatter_reduce(I.R_SUM, dst, src, idx)
    return Case("scatter_reduce", b, 16, [np.zeros(3, np.uint32)], expect={0: np.array([16, 0, 0], np.uint32)})


def scatter_atomic_u32():
    """test.rs:829-861: previous values returned by the atomic are unique, dst[0] == n."""
    b = irm.IRBuilder()
    u32 = b.scalar(I.U32)
    dst = b.buffer_ref(u32)
    src = b.literal(I.U32, 1)
    idx = b.literal(I.I32, 0)
    prev = b.push(I.OP_SCATTER_ATOMIC, u32, [dst, src, idx], arg=I.R_SUM)
    _out(b, I.U32, prev, b.index())
    return Case("scatter_atomic_u32", b, 16, [np.zeros(3, np.uint32), np.zeros(16, np.uint32)],
                expect={0: np.array([16, 0, 0], np.uint32)}, unordered_unique=(1,), check_slots=(0,))


def atomic_inc(n=1000, p=1.0, seed=0):
    """test.rs:1357-1379 / :1380-1409: ids handed out by atomic_inc are unique among active lanes."""
    b = irm.IRBuilder()
    u32, boolt = b.scalar(I.U32), b.scalar(I.BOOL)
    atomics = b.buffer_ref(u32)
    one = b.literal(I.U32, 1)
    act_ref = b.buffer_ref(boolt)
    idx = b.index()
    act = b.gather(boolt, act_ref, idx)
    ids = b.push(I.OP_ATOMIC_INC, u32, [atomics, one, act])
    out = b.buffer_ref(u32)
    b.scatter(out, ids, idx, act)
    rng = np.random.Generator(np.random.PCG64(seed))
    active = (rng.random(n) < p).astype(np.uint8)
    return Case("atomic_inc", b, n, [np.zeros(3, np.uint32), active, np.full(n, 0xFFFFFFFF, np.uint32)],
                expect={0: np.array([0, int(active.sum()), 0], np.uint32)}, unordered_unique=(2,),
                check_slots=(0,))


def _loop_ir(kind_is_loop=True, with_side_effect=False):
    """Shape of loop_record1 / if_record1 (test.rs:1411-1462): state = (cond: bool, i: i32)."""
    b = irm.IRBuilder()
    boolt, i32 = b.scalar(I.BOOL), b.scalar(I.I32)
    st = b.struct([boolt, i32])
    return b, boolt, i32, st


def loop_record1():
    """test.rs:1436-1462: i=[0,1]; while c { i += 1; c &= i < 2 } -> [2, 2]."""
    b, boolt, i32, st = _loop_ir()
    c0 = b.literal(I.BOOL, True)
    i_ref = b.buffer_ref(i32)
    idx = b.index()
    i0 = b.gather(i32, i_ref, idx)
    s0 = b.push(I.OP_CONSTRUCT, st, [c0, i0])
    ls = b.push(I.OP_LOOP_START, st, [s0])
    c1 = b.push(I.OP_EXTRACT, boolt, [ls], arg=0)
    i1 = b.push(I.OP_EXTRACT, i32, [ls], arg=1)
    i2 = b.bop(I.BOP_ADD, i32, i1, b.literal(I.I32, 1))
    lt = b.bop(I.BOP_LT, boolt, i2, b.literal(I.I32, 2))
    c2 = b.bop(I.BOP_AND, boolt, c1, lt)
    s1 = b.push(I.OP_CONSTRUCT, st, [c2, i2])
    le = b.push(I.OP_LOOP_END, st, [ls, s1])
    i_out = b.push(I.OP_EXTRACT, i32, [le], arg=1)
    c_out = b.push(I.OP_EXTRACT, boolt, [le], arg=0)
    _out(b, I.I32, i_out, idx)
    _out(b, I.BOOL, c_out, idx)
    return Case("loop_record", b, 2, [np.array([0, 1], np.int32), np.zeros(2, np.int32), np.ones(2, np.uint8)],
                expect={1: np.array([2, 2], np.int32), 2: np.array([0, 0], np.uint8)})


def if_record1():
    """test.rs:1411-1434: if c { i += 1 } with c=[true,false], i=[0,0] -> [1, 0]."""
    b, boolt, i32, st = _loop_ir()
    c_ref = b.buffer_ref(boolt)
    idx = b.index()
    c0 = b.gather(boolt, c_ref, idx)
    i_ref = b.buffer_ref(i32)
    i0 = b.gather(i32, i_ref, idx)
    s0 = b.push(I.OP_CONSTRUCT, st, [c0, i0])
    fs = b.push(I.OP_IF_START, st, [s0])
    c1 = b.push(I.OP_EXTRACT, boolt, [fs], arg=0)
    i1 = b.push(I.OP_EXTRACT, i32, [fs], arg=1)
    i2 = b.bop(I.BOP_ADD, i32, i1, b.literal(I.I32, 1))
    s1 = b.push(I.OP_CONSTRUCT, st, [c1, i2])
    fe = b.push(I.OP_LOOP_END, st, [fs, s1])  # if_end emits LoopEnd (trace.rs:510)
    i_out = b.push(I.OP_EXTRACT, i32, [fe], arg=1)
    _out(b, I.I32, i_out, idx)
    return Case("if_record1", b, 2, [np.array([1, 0], np.uint8), np.zeros(2, np.int32), np.zeros(2, np.int32)],
                expect={2: np.array([1, 0], np.int32)})


def loop_side_effect():
    """test.rs:1488-1510: one lane loops i = 0..3 scattering 1 into dst[i] -> [1,1,1,1,0,...]."""
    b, boolt, i32, st = _loop_ir()
    c0 = b.literal(I.BOOL, True)
    i0 = b.literal(I.I32, 0)
    s0 = b.push(I.OP_CONSTRUCT, st, [c0, i0])
    ls = b.push(I.OP_LOOP_START, st, [s0])
    c1 = b.push(I.OP_EXTRACT, boolt, [ls], arg=0)
    i1 = b.push(I.OP_EXTRACT, i32, [ls], arg=1)
    i2 = b.bop(I.BOP_ADD, i32, i1, b.literal(I.I32, 1))
    lt = b.bop(I.BOP_LT, boolt, i2, b.literal(I.I32, 4))
    c2 = b.bop(I.BOP_AND, boolt, c1, lt)
    s1 = b.push(I.OP_CONSTRUCT, st, [c2, i2])
    dst = b.buffer_ref(i32)
    sc = b.scatter(dst, b.literal(I.I32, 1), i1)  # side effect attached to the loop end
    le = b.push(I.OP_LOOP_END, st, [ls, s1, sc])
    i_out = b.push(I.OP_EXTRACT, i32, [le], arg=1)
    _out(b, I.I32, i_out, b.index())
    return Case("loop_side_effect", b, 1, [np.zeros(10, np.int32), np.zeros(1, np.int32)],
                expect={0: np.array([1, 1, 1, 1, 0, 0, 0, 0, 0, 0], np.int32), 1: np.array([4], np.int32)})


def vec3_memory_layout():
    """test.rs:1322-1336: vec3(1,2,3) stored packed -> [1,2,3,1,2,3]."""
    b = irm.IRBuilder()
    i32 = b.scalar(I.I32)
    v3 = b.vec(i32, 3)
    parts = [b.literal(I.I32, k) for k in (1, 2, 3)]
    v = b.push(I.OP_CONSTRUCT, v3, parts)
    ref = b.buffer_ref(v3)
    b.scatter(ref, v, b.index())
    return Case("vec3_memory_layout", b, 2, [np.zeros((2, 3), np.int32)],
                expect={0: np.array([[1, 2, 3], [1, 2, 3]], np.int32)})


def cast_array_vec():
    """test.rs:1337-1356: arr(1f,2f,3f) -> vec3<f32> -> array<i32,3>."""
    b = irm.IRBuilder()
    f32, i32 = b.scalar(I.F32), b.scalar(I.I32)
    a = b.push(I.OP_CONSTRUCT, b.array(f32, 3), [b.literal(I.F32, x) for x in (1.0, 2.0, 3.0)])
    v = b.uop(I.UOP_CAST, b.vec(f32, 3), a)
    ai = b.uop(I.UOP_CAST, b.array(i32, 3), v)
    idx = b.index()
    r0 = b.buffer_ref(b.vec(f32, 3))
    b.scatter(r0, v, idx)
    r1 = b.buffer_ref(b.array(i32, 3))
    b.scatter(r1, ai, idx)
    return Case("cast_array_vec", b, 2, [np.zeros((2, 3), np.float32), np.zeros((2, 3), np.int32)],
                expect={1: np.array([[1, 2, 3], [1, 2, 3]], np.int32)})


def array_dyn_extract():
    """test.rs:1288-1321: arr(1,2,3).extract_dyn(index)."""
    b = irm.IRBuilder()
    i32 = b.scalar(I.I32)
    a = b.push(I.OP_CONSTRUCT, b.array(i32, 3), [b.literal(I.I32, k) for k in (1, 2, 3)])
    idx = b.index()
    e = b.push(I.OP_DYN_EXTRACT, i32, [a, idx])
    _out(b, I.I32, e, idx)
    return Case("dyn_extract", b, 3, [np.zeros(3, np.int32)], expect={0: np.array([1, 2, 3], np.int32)})


def struct_roundtrip():
    """test.rs:214-236 (test_struct): composite(u8, u32) constructed, extracted."""
    b = irm.IRBuilder()
    u8, u32 = b.scalar(I.U8), b.scalar(I.U32)
    st = b.struct([u8, u32])
    s = b.push(I.OP_CONSTRUCT, st, [b.literal(I.U8, 1), b.literal(I.U32, 2)])
    idx = b.index()
    _out(b, I.U8, b.push(I.OP_EXTRACT, u8, [s], arg=0), idx)
    _out(b, I.U32, b.push(I.OP_EXTRACT, u32, [s], arg=1), idx)
    return Case("struct_roundtrip", b, 10, [np.zeros(10, np.uint8), np.zeros(10, np.uint32)],
                expect={0: np.ones(10, np.uint8), 1: np.full(10, 2, np.uint32)})


REFERENCE_CASES = [select, conditional_scatter, conditional_gather, simple1_second_pass, literal_fill,
                   lambda: literal_fill(I.BOOL, True, 100), simple_f16, scatter_chain_add, uop_cos,
                   scatter_reduce_kat, scatter_atomic_u32, atomic_inc, lambda: atomic_inc(1000, 0.5, 1),
                   loop_record1, if_record1, loop_side_effect, vec3_memory_layout, cast_array_vec,
                   array_dyn_extract, struct_roundtrip]


# ---- synthetic coverage: every scalar op on every type ----------------------------------------

INT_KINDS = [I.I8, I.U8, I.I16, I.U16, I.I32, I.U32, I.I64, I.U64]
FLOAT_KINDS = [I.F32, I.F64]
ARITH_BOPS = [I.BOP_ADD, I.BOP_SUB, I.BOP_MUL, I.BOP_MIN, I.BOP_MAX]
CMP_BOPS = [I.BOP_EQ, I.BOP_NEQ, I.BOP_LT, I.BOP_LE, I.BOP_GT, I.BOP_GE]


def _rand(rng, kind, n, nonzero=False, small=False):
    dt = NP[kind]
    if kind in (I.F16, I.F32, I.F64):
        x = (rng.random(n) * 8 - 4).astype(dt)
        if nonzero:
            x = np.where(np.abs(x) < 0.25, dt(1.5), x).astype(dt)
        return x
    if kind == I.BOOL:
        return rng.integers(0, 2, size=n).astype(np.uint8)
    info = np.iinfo(dt)
    lo, hi = (0, 8) if small else (info.min, int(info.max) + 1)
    x = rng.integers(lo, hi, size=n, dtype=np.int64 if info.min < 0 else np.uint64).astype(dt)
    if nonzero:
        x = np.where(x == 0, dt(3), x).astype(dt)
    return x


def binary_ops(kind, n=4099, seed=0):
    """All binary ops of one scalar type in a single kernel: two inputs, one output per op."""
    b = irm.IRBuilder()
    t, boolt = b.scalar(kind), b.scalar(I.BOOL)
    ra, rb = b.buffer_ref(t), b.buffer_ref(t)
    idx = b.index()
    a, c = b.gather(t, ra, idx), b.gather(t, rb, idx)
    rng = np.random.Generator(np.random.PCG64(seed + kind))
    bufs = [_rand(rng, kind, n), _rand(rng, kind, n, nonzero=True)]
    ops = list(ARITH_BOPS)
    if kind != I.BOOL:
        ops += [I.BOP_DIV]
    if kind in INT_KINDS:
        ops += [I.BOP_MODULUS, I.BOP_AND, I.BOP_OR, I.BOP_XOR]
    for op in ops:
        _out(b, kind, b.bop(op, t, a, c), idx)
        bufs.append(np.zeros(n, I.MEM[kind]))
    for op in CMP_BOPS:
        _out(b, I.BOOL, b.bop(op, boolt, a, c), idx)
        bufs.append(np.zeros(n, np.uint8))
    if kind in INT_KINDS:  # shifts by a small amount (shift >= width is undefined in GLSL and C)
        rs = b.buffer_ref(t)
        sh = b.gather(t, rs, idx)
        bits = np.dtype(NP[kind]).itemsize * 8
        bufs.append((rng.integers(0, bits, size=n)).astype(NP[kind]))
        for op in (I.BOP_SHL, I.BOP_SHR):
            _out(b, kind, b.bop(op, t, a, sh), idx)
            bufs.append(np.zeros(n, NP[kind]))
    is_f = kind in (I.F32, I.F64, I.F16)
    return Case(f"binary_ops_{kind}", b, n, bufs, exact=not is_f, rtol=1e-6 if kind == I.F32 else 1e-14)


def bool_ops(n=1000):
    b = irm.IRBuilder()
    boolt = b.scalar(I.BOOL)
    ra, rb = b.buffer_ref(boolt), b.buffer_ref(boolt)
    idx = b.index()
    a, c = b.gather(boolt, ra, idx), b.gather(boolt, rb, idx)
    rng = np.random.Generator(np.random.PCG64(5))
    bufs = [_rand(rng, I.BOOL, n), _rand(rng, I.BOOL, n)]
    for op in (I.BOP_AND, I.BOP_OR, I.BOP_XOR, I.BOP_EQ, I.BOP_NEQ):
        _out(b, I.BOOL, b.bop(op, boolt, a, c), idx)
        bufs.append(np.zeros(n, np.uint8))
    _out(b, I.BOOL, b.uop(I.UOP_NEG, boolt, a), idx)
    bufs.append(np.zeros(n, np.uint8))
    return Case("bool_ops", b, n, bufs)


def unary_ops(kind, n=4099):
    b = irm.IRBuilder()
    t = b.scalar(kind)
    ra = b.buffer_ref(t)
    idx = b.index()
    a = b.gather(t, ra, idx)
    rng = np.random.Generator(np.random.PCG64(kind))
    x = _rand(rng, kind, n)
    bufs = [x]
    ops = [I.UOP_NEG, I.UOP_ABS]
    if kind in FLOAT_KINDS:
        ops += [I.UOP_SIN, I.UOP_COS, I.UOP_EXP2]
    for op in ops:
        _out(b, kind, b.uop(op, t, a), idx)
        bufs.append(np.zeros(n, NP[kind]))
    if kind in FLOAT_KINDS:  # sqrt / log2 of |a| + 0.5
        pos = b.bop(I.BOP_ADD, t, b.uop(I.UOP_ABS, t, a), b.literal(kind, 0.5))
        for op in (I.UOP_SQRT, I.UOP_LOG2):
            _out(b, kind, b.uop(op, t, pos), idx)
            bufs.append(np.zeros(n, NP[kind]))
        fm = b.fma(t, a, b.literal(kind, 1.5), b.literal(kind, 0.25))
        _out(b, kind, fm, idx)
        bufs.append(np.zeros(n, NP[kind]))
    else:
        fm = b.fma(t, a, b.literal(kind, 3), b.literal(kind, 7))
        _out(b, kind, fm, idx)
        bufs.append(np.zeros(n, NP[kind]))
    is_f = kind in FLOAT_KINDS
    # CUDA's sinf/cosf/exp2f/log2f are within 2 ulp; the Vulkan precision table the reference
    # runs under is far looser (sin/cos abs 2^-11, exp2 3+2|x| ulp, log2 3 ulp)
    return Case(f"unary_ops_{kind}", b, n, bufs, exact=not is_f, rtol=4e-7 if kind == I.F32 else 1e-15,
                atol=2e-7 if kind == I.F32 else 1e-15)


def casts(src_kind, n=2053):
    """Cast one source type to every scalar type, plus same-size bitcasts."""
    b = irm.IRBuilder()
    t = b.scalar(src_kind)
    ra = b.buffer_ref(t)
    idx = b.index()
    a = b.gather(t, ra, idx)
    rng = np.random.Generator(np.random.PCG64(100 + src_kind))
    if src_kind in (I.F32, I.F64, I.F16):
        x = (rng.random(n) * 200 - 100).astype(NP[src_kind])  # in range for every int >= 8 bits? no: clip
        x = np.clip(x, -100, 100).astype(NP[src_kind])
    else:
        x = _rand(rng, src_kind, n)
    bufs = [x]
    for dk in [I.BOOL] + INT_KINDS + [I.F16, I.F32, I.F64]:
        if src_kind in (I.F32, I.F64, I.F16) and dk in (I.U8, I.U16, I.U32, I.U64, I.I8):
            continue  # negative / out-of-range float -> narrow or unsigned int is undefined in GLSL
        _out(b, dk, b.uop(I.UOP_CAST, b.scalar(dk), a), idx)
        bufs.append(np.zeros(n, I.MEM[dk]))
    same = {I.I32: [I.U32, I.F32], I.U32: [I.I32, I.F32], I.F32: [I.I32, I.U32], I.I64: [I.U64, I.F64],
            I.U64: [I.I64, I.F64], I.F64: [I.I64, I.U64], I.I16: [I.U16], I.U16: [I.I16], I.I8: [I.U8],
            I.U8: [I.I8]}.get(src_kind, [])
    for dk in same:
        _out(b, dk, b.uop(I.UOP_BITCAST, b.scalar(dk), a), idx)
        bufs.append(np.zeros(n, I.MEM[dk]))
    return Case(f"casts_{src_kind}", b, n, bufs)


def gather_scatter_permute(n=10007, seed=3):
    """out[perm[i]] = src[perm2[i]] * 2 — computed-index gather and scatter (no conflicts)."""
    b = irm.IRBuilder()
    f32, u32 = b.scalar(I.F32), b.scalar(I.U32)
    src, p_ref, q_ref, dst = b.buffer_ref(f32), b.buffer_ref(u32), b.buffer_ref(u32), b.buffer_ref(f32)
    idx = b.index()
    p = b.gather(u32, p_ref, idx)
    q = b.gather(u32, q_ref, idx)
    v = b.gather(f32, src, q)
    v2 = b.bop(I.BOP_MUL, f32, v, b.literal(I.F32, 2.0))
    b.scatter(dst, v2, p)
    rng = np.random.Generator(np.random.PCG64(seed))
    return Case("gather_scatter_permute", b, n,
                [rng.random(n, dtype=np.float32), rng.permutation(n).astype(np.uint32),
                 rng.integers(0, n, size=n).astype(np.uint32), np.zeros(n, np.float32)])


def scatter_reduce_ops(kind=I.U32, rop=I.R_SUM, n=50021, n_bins=257, seed=4, cond=False):
    b = irm.IRBuilder()
    t, u32, boolt = b.scalar(kind), b.scalar(I.U32), b.scalar(I.BOOL)
    dst, k_ref, v_ref = b.buffer_ref(t), b.buffer_ref(u32), b.buffer_ref(t)
    idx = b.index()
    k = b.gather(u32, k_ref, idx)
    v = b.gather(t, v_ref, idx)
    c = None
    rng = np.random.Generator(np.random.PCG64(seed + rop))
    bufs = [_rand(rng, kind, n_bins), rng.integers(0, n_bins, size=n).astype(np.uint32), _rand(rng, kind, n)]
    if cond:
        c_ref = b.buffer_ref(boolt)
        c = b.gather(boolt, c_ref, idx)
        bufs.append(_rand(rng, I.BOOL, n))
    b.scatter_reduce(rop, dst, v, k, c)
    is_f = kind in (I.F32, I.F64)
    # float sums: the atomics add in an arbitrary order, ~200 values of magnitude <= 4 and mixed sign per bin — a
    # bin that sums to almost zero still carries an absolute error of ~200 * 4 * 2^-24, hence the absolute term
    return Case(f"scatter_reduce_{kind}_{rop}", b, n, bufs, exact=not (is_f and rop == I.R_SUM), rtol=1e-4, atol=1e-3)


def dyn_size(n=1000, live=137):
    """Kernel sized by a device-resident count (Extent::DynSize, graph.rs:503-508)."""
    b = irm.IRBuilder()
    u32 = b.scalar(I.U32)
    idx = b.index()
    _out(b, I.U32, b.bop(I.BOP_ADD, u32, idx, b.literal(I.U32, 1)), idx)
    exp = np.zeros(n, np.uint32)
    exp[:live] = np.arange(1, live + 1)
    return Case("dyn_size", b, n, [np.zeros(n, np.uint32)], expect={0: exp},
                size_buf=np.array([live], np.uint32))


def index_base(n=5000, base=123456):
    """Sharded launch: Index yields base + local index, memory is addressed locally."""
    b = irm.IRBuilder()
    u32 = b.scalar(I.U32)
    ref = b.buffer_ref(u32)
    idx = b.index()
    x = b.gather(u32, ref, idx)
    _out(b, I.U32, b.bop(I.BOP_ADD, u32, x, idx), idx)
    src = np.arange(n, dtype=np.uint32) * 3
    return Case("index_base", b, n, [src, np.zeros(n, np.uint32)], index_base=base,
                expect={1: src + np.arange(n, dtype=np.uint32) + np.uint32(base)})


def c2_chain(n=100003, seed=0):
    """BASELINE.json config C2."""
    b = irm.c2_chain_ir()
    rng = np.random.Generator(np.random.PCG64(seed))
    x = (rng.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
    return Case("c2_chain", b, n, [x, np.zeros(n, np.float32)], exact=False, rtol=4e-7, atol=1e-7)


def mixed_width(n=4100):
    """f64 + u8 + bool buffers in one kernel: exercises the vector-geometry choice (VEC = 2)."""
    b = irm.IRBuilder()
    f64, u8, boolt = b.scalar(I.F64), b.scalar(I.U8), b.scalar(I.BOOL)
    ra, rb = b.buffer_ref(f64), b.buffer_ref(u8)
    idx = b.index()
    a, c = b.gather(f64, ra, idx), b.gather(u8, rb, idx)
    s = b.bop(I.BOP_ADD, f64, a, b.uop(I.UOP_CAST, f64, c))
    _out(b, I.F64, s, idx)
    _out(b, I.BOOL, b.bop(I.BOP_GT, boolt, s, b.literal(I.F64, 100.0)), idx)
    rng = np.random.Generator(np.random.PCG64(8))
    return Case("mixed_width", b, n, [rng.random(n) * 200, rng.integers(0, 256, size=n).astype(np.uint8),
                                      np.zeros(n), np.zeros(n, np.uint8)])


def in_place_update(n=3000):
    """record_test (test.rs:1063-1083): a = a + 1 scattered back into the same buffer."""
    b = irm.IRBuilder()
    i32 = b.scalar(I.I32)
    ref = b.buffer_ref(i32)
    idx = b.index()
    x = b.gather(i32, ref, idx)
    b.scatter(ref, b.bop(I.BOP_ADD, i32, x, b.literal(I.I32, 1)), idx)
    src = np.arange(n, dtype=np.int32)
    return Case("in_place_update", b, n, [src.copy()], expect={0: src + 1})


def vector_ops(n=1025):
    """Componentwise vec3<f32> arithmetic, dot (Inner), sqrt and FMA on vectors."""
    b = irm.IRBuilder()
    f32 = b.scalar(I.F32)
    v3 = b.vec(f32, 3)
    ra, rb = b.buffer_ref(v3), b.buffer_ref(v3)
    idx = b.index()
    a, c = b.gather(v3, ra, idx), b.gather(v3, rb, idx)
    s = b.bop(I.BOP_ADD, v3, a, c)
    m = b.bop(I.BOP_MUL, v3, s, c)
    d = b.bop(I.BOP_INNER, f32, a, c)
    q = b.uop(I.UOP_SQRT, v3, b.uop(I.UOP_ABS, v3, m))
    f = b.fma(v3, a, c, q)
    r0 = b.buffer_ref(v3)
    b.scatter(r0, f, idx)
    _out(b, I.F32, d, idx)
    rng = np.random.Generator(np.random.PCG64(9))
    return Case("vector_ops", b, n, [rng.random((n, 3), dtype=np.float32), rng.random((n, 3), dtype=np.float32),
                                     np.zeros((n, 3), np.float32), np.zeros(n, np.float32)],
                exact=False, rtol=1e-6, atol=1e-6)


def matrix_product(n=257):
    """test.rs matrix_times_matrix: `m0.mul(&m1)` on matrices is GLSL's linear-algebra product
    (column-major).  Lane 0..: A = [[1,2],[3,4]] + index, B = [[5,6],[7,8]]; lane 0 is the reference's
    program, whose product [[19,22],[43,50]] is stored column by column."""
    b = irm.IRBuilder()
    f32 = b.scalar(I.F32)
    v2, m2 = b.vec(f32, 2), b.mat(f32, 2, 2)
    idx = b.index()
    fi = b.uop(I.UOP_CAST, f32, idx)
    lit = lambda x: b.bop(I.BOP_ADD, f32, b.literal(I.F32, x), fi)
    a0 = b.push(I.OP_CONSTRUCT, v2, [lit(1.0), lit(3.0)])
    a1 = b.push(I.OP_CONSTRUCT, v2, [lit(2.0), lit(4.0)])
    ma = b.push(I.OP_CONSTRUCT, m2, [a0, a1])
    b0 = b.push(I.OP_CONSTRUCT, v2, [b.literal(I.F32, 5.0), b.literal(I.F32, 7.0)])
    b1 = b.push(I.OP_CONSTRUCT, v2, [b.literal(I.F32, 6.0), b.literal(I.F32, 8.0)])
    mb = b.push(I.OP_CONSTRUCT, m2, [b0, b1])
    prod = b.bop(I.BOP_MUL, m2, ma, mb)
    r0 = b.buffer_ref(m2)
    b.scatter(r0, prod, idx)
    i = np.arange(n, dtype=np.float32)
    A = np.stack([np.stack([1 + i, 2 + i], -1), np.stack([3 + i, 4 + i], -1)], -2)  # (n, row, col)
    B = np.array([[5, 6], [7, 8]], np.float32)
    want = np.einsum("nij,jk->nik", A, B).transpose(0, 2, 1).reshape(n, 4).astype(np.float32)  # column-major
    assert want[0].tolist() == [19.0, 43.0, 22.0, 50.0]
    return Case("matrix_product", b, n, [np.zeros((n, 4), np.float32)], expect={0: want})


def rmw_through_index_list(n=3 * 4096 + 517, seed=11):
    """a[p[i]] = a[p[i]] * 3 + 1 and flags[p[i]] = a[p[i]] > t, p a permutation: the wavefront update
    (test.rs:1020-1062).  The slot is gathered AND scattered through a computed index: the vector entry
    runs the gathers of all its elements first (codegen.cpp: split_point); full tiles and a ragged tail."""
    b = irm.IRBuilder()
    u32, boolt = b.scalar(I.U32), b.scalar(I.BOOL)
    ra, rp = b.buffer_ref(u32), b.buffer_ref(u32)
    idx = b.index()
    p = b.gather(u32, rp, idx)
    v = b.bop(I.BOP_ADD, u32, b.bop(I.BOP_MUL, u32, b.gather(u32, ra, p), b.literal(I.U32, 3)), b.literal(I.U32, 1))
    b.scatter(ra, v, p)
    b.scatter(b.buffer_ref(boolt), b.bop(I.BOP_GT, boolt, v, b.literal(I.U32, 1 << 31)), p)
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    perm = rng.permutation(n).astype(np.uint32)
    want = a * np.uint32(3) + np.uint32(1)
    return Case("rmw_through_index_list", b, n, [a.copy(), perm, np.zeros(n, np.uint8)],
                expect={0: want, 2: (want > np.uint32(1 << 31)).astype(np.uint8)})


def rmw_then_loop(n=2 * 4096 + 33, seed=12):
    """A vec3 value and a conditional scatter behind the cut, a recorded loop after it:
    w = (a[p[i]], a[p[i]] + 1, 2); if a[p[i]] is even: a[p[i]] = w.x + w.y; then k = 0; while k < 3: k += 1;
    out[i] = k + w.z.  Everything up to the first scatter runs for all elements of a thread first."""
    b = irm.IRBuilder()
    i32, boolt = b.scalar(I.I32), b.scalar(I.BOOL)
    u32 = b.scalar(I.U32)
    v3 = b.vec(i32, 3)
    st = b.struct([boolt, i32])
    ra, rp = b.buffer_ref(i32), b.buffer_ref(u32)
    idx = b.index()
    p = b.gather(u32, rp, idx)
    x = b.gather(i32, ra, p)
    w = b.push(I.OP_CONSTRUCT, v3, [x, b.bop(I.BOP_ADD, i32, x, b.literal(I.I32, 1)), b.literal(I.I32, 2)])
    even = b.bop(I.BOP_EQ, boolt, b.bop(I.BOP_AND, i32, x, b.literal(I.I32, 1)), b.literal(I.I32, 0))
    wx, wy = b.push(I.OP_EXTRACT, i32, [w], arg=0), b.push(I.OP_EXTRACT, i32, [w], arg=1)
    b.scatter(ra, b.bop(I.BOP_ADD, i32, wx, wy), p, even)
    s0 = b.push(I.OP_CONSTRUCT, st, [b.literal(I.BOOL, True), b.literal(I.I32, 0)])
    ls = b.push(I.OP_LOOP_START, st, [s0])
    k1 = b.push(I.OP_EXTRACT, i32, [ls], arg=1)
    k2 = b.bop(I.BOP_ADD, i32, k1, b.literal(I.I32, 1))
    s1 = b.push(I.OP_CONSTRUCT, st, [b.bop(I.BOP_LT, boolt, k2, b.literal(I.I32, 3)), k2])
    le = b.push(I.OP_LOOP_END, st, [ls, s1])
    k_out = b.push(I.OP_EXTRACT, i32, [le], arg=1)
    _out(b, I.I32, b.bop(I.BOP_ADD, i32, k_out, b.push(I.OP_EXTRACT, i32, [w], arg=2)), idx)
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(-1000, 1000, size=n).astype(np.int32)
    perm = rng.permutation(n).astype(np.uint32)
    want = np.where(a % 2 == 0, 2 * a + 1, a).astype(np.int32)
    return Case("rmw_then_loop", b, n, [a.copy(), perm, np.zeros(n, np.int32)],
                expect={0: want, 2: np.full(n, 5, np.int32)})


SYNTHETIC_CASES = (
    [lambda k=k: binary_ops(k) for k in INT_KINDS + FLOAT_KINDS]
    + [bool_ops]
    + [lambda k=k: unary_ops(k) for k in [I.I8, I.I16, I.I32, I.I64, I.U32, I.F32, I.F64]]
    + [lambda k=k: casts(k) for k in [I.BOOL] + INT_KINDS + FLOAT_KINDS]
    + [gather_scatter_permute]
    + [lambda r=r: scatter_reduce_ops(I.U32, r) for r in (I.R_MAX, I.R_MIN, I.R_SUM, I.R_OR, I.R_AND, I.R_XOR)]
    + [lambda: scatter_reduce_ops(I.I32, I.R_MIN), lambda: scatter_reduce_ops(I.U64, I.R_SUM),
       lambda: scatter_reduce_ops(I.I64, I.R_MAX), lambda: scatter_reduce_ops(I.F32, I.R_SUM),
       lambda: scatter_reduce_ops(I.F32, I.R_MAX), lambda: scatter_reduce_ops(I.U32, I.R_SUM, cond=True)]
    + [dyn_size, index_base, c2_chain, mixed_width, in_place_update, vector_ops, matrix_product]
    + [rmw_through_index_list, rmw_then_loop]
)

ALL_CASES = REFERENCE_CASES + SYNTHETIC_CASES


def case_id(fn):
    try:
        return fn().name
    except Exception:  # pragma: no cover
        return getattr(fn, "__name__", "case")
