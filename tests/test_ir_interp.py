"""Pins the numpy IR interpreter (oracle/ir_interp.py) against the reference's own elementwise /
scatter / gather / loop known-answer tests, and checks the product's code generator on the CPU:
every case lowers to CUDA C++ and NVRTC compiles it to an sm_100a cubin (no GPU needed)."""
import copy
import importlib

import numpy as np
import pytest

import ir_cases
from oracle import ir_interp
from oracle import ir_interp as I  # op / type constants

irm = importlib.import_module("hephaestus-jit_b200.ir")


def run_interp(case):
    bufs = [np.array(b, copy=True) for b in case.buffers]
    ir_interp.run_ir(case.builder, case.size, bufs, size_buf=case.size_buf, index_base=case.index_base)
    return bufs


@pytest.mark.parametrize("make", ir_cases.REFERENCE_CASES, ids=ir_cases.case_id)
def test_interpreter_reproduces_reference_kats(make):
    case = make()
    out = run_interp(case)
    assert case.expect, "reference cases must state an expected output"
    for slot, want in case.expect.items():
        got = out[slot].reshape(want.shape)
        if case.exact or np.issubdtype(want.dtype, np.integer):
            assert np.array_equal(got, want), (case.name, slot, got, want)
        else:
            assert np.allclose(got, want, rtol=case.rtol, atol=case.atol)
    for slot in case.unordered_unique:
        vals = out[slot][out[slot] != 0xFFFFFFFF]
        assert len(set(vals.tolist())) == len(vals)


@pytest.mark.parametrize("make", ir_cases.SYNTHETIC_CASES, ids=ir_cases.case_id)
def test_interpreter_runs_synthetic_cases(make):
    case = make()
    out = run_interp(case)
    for slot, want in case.expect.items():
        assert np.array_equal(out[slot].reshape(want.shape), want)


def test_interpreter_c2_equals_c_oracle():
    import oracle
    case = ir_cases.c2_chain()
    out = run_interp(case)
    assert np.array_equal(out[1], oracle.c2_chain(case.buffers[0]))


def test_interpreter_integer_division_truncates():
    b = irm.IRBuilder()
    i32 = b.scalar(irm.I32)
    ra, rb = b.buffer_ref(i32), b.buffer_ref(i32)
    idx = b.index()
    a, c = b.gather(i32, ra, idx), b.gather(i32, rb, idx)
    for op in (irm.BOP_DIV, irm.BOP_MODULUS):
        r = b.buffer_ref(i32)
        b.scatter(r, b.bop(op, i32, a, c), idx)
    x = np.array([7, -7, 7, -7], np.int32)
    y = np.array([2, 2, -2, -2], np.int32)
    bufs = [x, y, np.zeros(4, np.int32), np.zeros(4, np.int32)]
    ir_interp.run_ir(b, 4, bufs)
    assert bufs[2].tolist() == [3, -3, -3, 3] and bufs[3].tolist() == [1, -1, 1, -1]


# ---- code generator / NVRTC (CPU) ------------------------------------------------------------

@pytest.mark.parametrize("make", ir_cases.ALL_CASES, ids=ir_cases.case_id)
def test_codegen_compiles_to_sm100a_cubin(make, tmp_path, monkeypatch):
    monkeypatch.setenv("HJ_CACHE_DIR", str(tmp_path))
    case = make()
    ir = case.builder.build()
    src = irm.codegen(ir)
    assert "hj_kernel_scalar" in src
    cubin = irm.compile_cubin(ir)
    assert cubin[:4] == b"\x7fELF" and len(cubin) > 1000
    assert len(list(tmp_path.glob("*.cubin"))) == 1  # published to the on-disk cache
    assert irm.compile_cubin(ir) == cubin            # second call is served from it


def test_vector_entry_only_when_something_is_staged():
    src = irm.codegen(ir_cases.c2_chain().builder.build())
    assert "hj_kernel_vec" in src and "hj_load_vec<f32, 4>" in src and "hj_store_vec<f32, 4>" in src
    # a buffer that is read and written in place is not staged -> no vector entry at all
    assert "hj_kernel_vec" not in irm.codegen(ir_cases.in_place_update().builder.build())
    # 8-byte elements: 2 per 128-bit access, and the 1-byte companions move as 2-byte accesses
    src = irm.codegen(ir_cases.mixed_width().builder.build())
    assert "hj_load_vec<f64, 2>" in src and "hj_load_vec<u8, 2>" in src


def test_ir_hash_is_content_hash():
    a = ir_cases.c2_chain().builder.build()
    b = ir_cases.c2_chain().builder.build()
    c = ir_cases.select().builder.build()
    assert irm.ir_hash(a) == irm.ir_hash(b) != irm.ir_hash(c)
    other = irm.c2_chain_ir()
    ty, op, arg, ds, de, data = other.vars[5]
    other.vars[5] = (ty, op, arg, ds, de, data ^ 1)  # flip one literal bit
    assert irm.ir_hash(other.build()) != irm.ir_hash(a)


def test_malformed_ir_is_rejected():
    hj = importlib.import_module("hephaestus-jit_b200")
    b = irm.IRBuilder()
    f32 = b.scalar(irm.F32)
    b.push(irm.OP_BOP, f32, [5, 6], arg=irm.BOP_ADD)  # forward references
    with pytest.raises(hj.HjError) as e:
        irm.codegen(b.build())
    assert e.value.status == 1
    b = irm.IRBuilder()
    b.push(irm.OP_TEX_LOOKUP, b.scalar(irm.F32), [])
    with pytest.raises(hj.HjError):
        irm.codegen(b.build())
    b = irm.IRBuilder()  # ScatterReduce(Prod) is todo!() in the reference
    u32 = b.scalar(irm.U32)
    r = b.buffer_ref(u32)
    b.scatter_reduce(3, r, b.literal(irm.U32, 1), b.index())
    with pytest.raises(hj.HjError):
        irm.codegen(b.build())


# ---- execute_graph's decision to leave the index zero-fill of a Compress pass to the compaction ------------
def _zero_fill_match(builder, slot):
    import ctypes
    L = importlib.import_module("hephaestus-jit_b200._lib")
    var, only = ctypes.c_uint32(0xFFFFFFFF), ctypes.c_int32(-1)
    assert "hj_kernel" in irm.codegen(builder.build())  # a well-formed kernel: a rejection below is the matcher's
    hit = L.lib.hj_ir_index_zero_fill(ctypes.byref(builder.build()), slot, ctypes.byref(var), ctypes.byref(only))
    return (hit, var.value, only.value)


def test_index_zero_fill_matcher():
    """The store may only be dropped when dropping it cannot be observed: a plain, unconditional,
    top-level `index[i] = 0u32` at the bare Index into a buffer nothing else in the kernel touches."""
    def fill_kernel(lit=0, kind=I.U32, with_mask=False, cond=False, idx_expr=False, read_back=False, in_loop=False):
        b = irm.IRBuilder()
        zero = b.literal(kind, lit)
        ref = b.buffer_ref(b.scalar(kind))
        idx = b.index()
        where = b.bop(I.BOP_ADD, b.scalar(I.U32), idx, b.literal(I.U32, 1)) if idx_expr else idx
        c = b.bop(I.BOP_LT, b.scalar(I.BOOL), idx, b.literal(I.U32, 5)) if cond else None
        if in_loop:
            state = b.push(I.OP_CONSTRUCT, b.struct([b.scalar(I.BOOL)]), [b.literal(I.BOOL, 0)])
            start = b.push(I.OP_LOOP_START, b.struct([b.scalar(I.BOOL)]), [state])
        sc = b.scatter(ref, zero, where, c)
        if in_loop:
            b.push(I.OP_LOOP_END, b.struct([b.scalar(I.BOOL)]), [start, state])
        if with_mask:  # the usual shape: the mask of the compaction is computed by the same kernel
            m = b.bop(I.BOP_LT, b.scalar(I.BOOL), idx, b.literal(I.U32, 10))
            b.scatter(b.buffer_ref(b.scalar(I.BOOL)), m, idx)
        if read_back:
            v = b.gather(b.scalar(kind), ref, idx)
            b.scatter(b.buffer_ref(b.scalar(kind)), v, idx)
        return b, sc

    b, sc = fill_kernel()
    assert _zero_fill_match(b, 0) == (1, sc, 1)                    # nothing else in the kernel: skip the pass
    b, sc = fill_kernel(with_mask=True)
    assert _zero_fill_match(b, 0) == (1, sc, 0)                    # the mask store stays, the fill goes
    assert _zero_fill_match(b, 1)[0] == 0                          # the mask buffer is not a zero fill
    assert _zero_fill_match(fill_kernel(lit=7)[0], 0)[0] == 0      # not zero
    assert _zero_fill_match(fill_kernel(kind=I.F32)[0], 0)[0] == 0  # not the u32 index type
    assert _zero_fill_match(fill_kernel(cond=True)[0], 0)[0] == 0  # conditional store
    assert _zero_fill_match(fill_kernel(idx_expr=True)[0], 0)[0] == 0  # not at the bare Index
    assert _zero_fill_match(fill_kernel(read_back=True)[0], 0)[0] == 0  # the kernel reads the buffer too
    assert _zero_fill_match(fill_kernel(in_loop=True)[0], 0)[0] == 0    # inside a recorded loop
    assert _zero_fill_match(fill_kernel()[0], 3)[0] == 0           # no such slot


def test_codegen_rejects_or_lowers_mutated_ir_without_crashing():
    """hj_ir_codegen is the boundary a foreign host hands IR to: a damaged IR (wild type ids, deps,
    op tags, slot counts) must be refused by validate_ir or lowered, never crash the library."""
    import random
    hj = importlib.import_module("hephaestus-jit_b200")
    rnd = random.Random(99)
    cases = [c() for c in ir_cases.ALL_CASES]
    lowered = rejected = 0
    for _ in range(400):
        b = copy.deepcopy(rnd.choice(cases).builder)
        b._keep = None
        for _ in range(rnd.choice((1, 1, 2, 3))):
            what = rnd.randrange(4)
            if what == 0:
                i = rnd.randrange(len(b.vars))
                v = list(b.vars[i])
                v[rnd.randrange(6)] = rnd.choice((0, 1, 2, 3, 5, 7, 20, 23, 24, 0xFFFFFFFF, rnd.randrange(64)))
                b.vars[i] = tuple(v)
            elif what == 1 and b.deps:
                b.deps[rnd.randrange(len(b.deps))] = rnd.choice((0, len(b.vars) - 1, len(b.vars), 0xFFFFFFFF, rnd.randrange(64)))
            elif what == 2:
                i = rnd.randrange(len(b.types))
                t = list(b.types[i])
                t[rnd.randrange(6)] = rnd.choice((0, 1, 3, 13, 14, 15, 16, 17, 255, 0xFFFFFFFF))
                b.types[i] = tuple(t)
            else:
                b.n_buffers = rnd.choice((0, 1, 2, 100))
        try:
            irm.codegen(b.build())
            lowered += 1
        except hj.HjError:
            rejected += 1
    assert lowered + rejected == 400 and rejected > 50
