"""Pins the CPU oracle (oracle/hj_oracle.c) against every known-answer vector the reference's
own tests hold for the device ops (tests/golden/reference_kats.json, transcribed from
hephaestus-jit/src/test.rs and benches/vulkan.rs), and against independent numpy folds on
seeded inputs.  CPU only."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import (AND, BOOL, F32, F64, I8, I16, I32, I64, MAX, MIN, OR, PROD, SUM, U8, U16,
                    U32, U64, XOR)

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_kats.json")) as f:
    KATS = json.load(f)

TY = {"Bool": BOOL, "I8": I8, "U8": U8, "I16": I16, "U16": U16, "I32": I32, "U32": U32,
      "I64": I64, "U64": U64, "F32": F32, "F64": F64}
OP = {"max": MAX, "min": MIN, "sum": SUM, "prod": PROD, "or": OR, "and": AND, "xor": XOR}


def host_fold(op, x):
    """The fold the reference's tests compare against (wrapping for ints)."""
    dt = x.dtype
    with np.errstate(over="ignore"):
        if op == "max":
            return x.max()
        if op == "min":
            return x.min()
        if op == "sum":
            return np.add.reduce(x, dtype=dt)
        if op == "prod":
            return np.multiply.reduce(x, dtype=dt)
        if op == "and":
            return np.bitwise_and.reduce(x)
        if op == "or":
            return np.bitwise_or.reduce(x)
        if op == "xor":
            return np.bitwise_xor.reduce(x)
    raise ValueError(op)


@pytest.mark.parametrize("case", KATS["reduce"], ids=lambda c: f"{c['op']}-{c['ty']}")
def test_reduce_kat(case):
    ty = TY[case["ty"]]
    lo, hi = case["range"]
    x = np.arange(lo, hi).astype(oracle.NP_DTYPE[ty])
    got = oracle.reduce(OP[case["op"]], ty, x)[0]
    assert got == case["expect"]


@pytest.mark.parametrize("case", KATS["reduce_random"], ids=lambda c: f"{c['op']}-{c['ty']}")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_reduce_random_vs_host_fold(case, seed):
    ty = TY[case["ty"]]
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.integers(case["lo"], case["hi"], size=case["n"]).astype(oracle.NP_DTYPE[ty])
    got = oracle.reduce(OP[case["op"]], ty, x)[0]
    assert got == host_fold(case["op"], x)


def test_reduce_prod_f32_kat():
    c = KATS["reduce_prod_f32"]
    rng = np.random.Generator(np.random.PCG64(0))
    k = rng.integers(c["k_lo"], c["k_hi"], size=c["n"]).astype(np.float32)
    x = np.log2(k * np.float32(0.01)).astype(np.float32)
    got = oracle.reduce(PROD, F32, x)[0]
    want = np.float32(1)
    for v in x:  # the reference folds left to right on the host
        want = np.float32(want * v)
    assert abs(float(got) - float(want)) <= c["abs_eps"] or (np.isnan(got) and np.isnan(want))


def test_reduce_table_matches_reference_support_matrix():
    """(op, type) pairs reduce.rs:84-166 implements vs todo!()."""
    arith = [I8, U8, I16, U16, I32, U32, I64, U64, F32, F64]
    bitw = [BOOL, U8, U16, U32, U64]
    for ty in range(1, 13):
        for op in (MAX, MIN, SUM, PROD):
            assert oracle.reduce_supported(op, ty) == (ty in arith), (op, ty)
        for op in (AND, OR, XOR):
            assert oracle.reduce_supported(op, ty) == (ty in bitw), (op, ty)
    with pytest.raises(NotImplementedError):
        oracle.reduce(SUM, oracle.F16, np.zeros(4, np.float16))


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 1000, 1024, 1025, 32768, 32769, 100003])
def test_reduce_sizes_and_tree_order(n):
    """Integer folds at pass-count boundaries (32^k, 32^k + 1); f32 sum equals an independent
    numpy restatement of the radix-32 stride-halving tree (reduce.glsl:40-47)."""
    rng = np.random.Generator(np.random.PCG64(n))
    x = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    assert oracle.reduce(SUM, U32, x)[0] == host_fold("sum", x)
    assert oracle.reduce(MAX, U32, x)[0] == x.max()
    assert oracle.reduce(XOR, U32, x)[0] == host_fold("xor", x)
    f = rng.random(n, dtype=np.float32)
    got = oracle.reduce(SUM, F32, f)[0]
    if n == 1:
        assert got == f[0]
        return
    n_passes = 0
    v = n - 1
    while v > 0:
        n_passes += 1
        v //= 32
    buf = np.zeros(32 ** n_passes, dtype=np.float32)
    buf[:n] = f
    while buf.size > 1:
        g = buf.reshape(-1, 32)
        for s in (16, 8, 4, 2, 1):
            g = (g[:, :s] + g[:, s:2 * s]).astype(np.float32)
        buf = g.reshape(-1)
    assert got == buf[0]
    assert abs(float(got) - float(f.astype(np.float64).sum())) <= 1e-5 * n


def test_reduce_faithful_copy_same_result():
    rng = np.random.Generator(np.random.PCG64(7))
    f = rng.random(50000, dtype=np.float32)
    assert oracle.reduce(SUM, F32, f)[0] == oracle.reduce(SUM, F32, f, faithful_copy=True)[0]


def test_prefix_sum_kat_u64():
    c = KATS["prefix_sum"]
    x = np.arange(c["n"], dtype=np.uint64)
    got = oracle.prefix_sum(U64, x, inclusive=True)
    assert np.array_equal(got, np.cumsum(x, dtype=np.uint64))
    assert int(got[-1]) == c["expect_last"]


def test_prefix_sum_bench_invariant_ref_compat():
    """benches/vulkan.rs:130-137: the reference's prefix_sum(false) of n ones ends in n (D10)."""
    c = KATS["prefix_sum_bench_invariant"]
    for n in c["n"]:
        x = np.ones(n, dtype=np.uint32)
        got = oracle.prefix_sum(U32, x, inclusive=False, ref_compat=True)
        assert int(got[-1]) == n
        real_excl = oracle.prefix_sum(U32, x, inclusive=False)
        assert int(real_excl[-1]) == n - 1 and int(real_excl[0]) == 0


@pytest.mark.parametrize("ty", [U8, I8, U16, I16, U32, I32, U64, I64])
@pytest.mark.parametrize("n", [1, 15, 16, 17, 2047, 2048, 2049, 8195, 70001])
@pytest.mark.parametrize("inclusive", [True, False])
def test_prefix_sum_ints_wrap_exact(ty, n, inclusive):
    dt = oracle.NP_DTYPE[ty]
    rng = np.random.Generator(np.random.PCG64(n + ty))
    info = np.iinfo(dt)
    x = rng.integers(info.min, int(info.max) + 1, size=n, dtype=np.int64 if info.min < 0 else np.uint64).astype(dt)
    got = oracle.prefix_sum(ty, x, inclusive)
    with np.errstate(over="ignore"):
        inc = np.cumsum(x, dtype=dt)
    want = inc if inclusive else np.concatenate([np.zeros(1, dt), inc[:-1]])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("inclusive", [True, False])
def test_prefix_sum_u32_mt_equals_serial(inclusive):
    rng = np.random.Generator(np.random.PCG64(3))
    x = rng.integers(0, 2**32, size=300001, dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(oracle.prefix_sum_u32_mt(x, inclusive),
                          oracle.prefix_sum(U32, x, inclusive))


@pytest.mark.parametrize("ty,tol", [(F32, 1e-5), (F64, 1e-13)])
def test_prefix_sum_float_close_to_f64(ty, tol):
    rng = np.random.Generator(np.random.PCG64(5))
    x = rng.random(50000).astype(oracle.NP_DTYPE[ty])
    got = oracle.prefix_sum(ty, x, True).astype(np.float64)
    want = np.cumsum(x.astype(np.float64))
    assert np.max(np.abs(got - want) / want) < tol


def test_compress_all_true_kat():
    n = KATS["compress_all_true"]["n"]
    count, idx = oracle.compress(np.ones(n, dtype=np.uint8))
    assert count == n
    assert np.array_equal(idx, np.arange(n, dtype=np.uint32))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_compress_small_kat(seed):
    n = KATS["compress_small"]["n"]
    rng = np.random.Generator(np.random.PCG64(seed))
    mask = rng.integers(0, 2, size=n).astype(np.uint8)
    count, idx = oracle.compress(mask)
    ref = np.flatnonzero(mask).astype(np.uint32)
    assert count == ref.size
    assert np.array_equal(idx[:count], ref)
    assert not idx[count:].any()  # tail keeps the zero fill


def test_compress_bench_invariant():
    for n in KATS["compress_bench_invariant"]["n"]:
        count, _ = oracle.compress(np.ones(n, dtype=np.uint8))
        assert count == n


@pytest.mark.parametrize("n", [1, 15, 16, 17, 2047, 2048, 2049, 4111, 100003])
@pytest.mark.parametrize("p", [0.0, 0.01, 0.5, 0.99, 1.0])
def test_compress_random_masks_ragged(n, p):
    """The case the reference's own test avoids (test.rs:924-926, defect D4): random masks at
    sizes not divisible by 16."""
    rng = np.random.Generator(np.random.PCG64(n))
    mask = (rng.random(n) < p).astype(np.uint8)
    sentinel = np.full(n, 0xDEADBEEF, dtype=np.uint32)
    count, idx = oracle.compress(mask, index_out=sentinel.copy(), index_base=7)
    ref = np.flatnonzero(mask).astype(np.uint32) + 7
    assert count == ref.size
    assert np.array_equal(idx[:count], ref)
    assert np.all(idx[count:] == 0xDEADBEEF)  # untouched beyond count
    count_mt, idx_mt = oracle.compress(mask, index_out=sentinel.copy(), index_base=7, mt=True)
    assert count_mt == count and np.array_equal(idx_mt, idx)


def test_scatter_reduce_kat():
    c = KATS["scatter_reduce"]
    dst = np.array(c["dst"], dtype=np.uint32)
    idx = np.full(c["n"], c["idx"], dtype=np.uint32)
    oracle.scatter_reduce(SUM, U32, idx, c["value"], dst)
    assert dst.tolist() == c["expect"]


def test_scatter_reduce_histogram_vs_bincount():
    rng = np.random.Generator(np.random.PCG64(11))
    idx = rng.integers(0, 1 << 16, size=1 << 18).astype(np.uint32)
    dst = np.zeros(1 << 16, dtype=np.uint32)
    oracle.scatter_reduce(SUM, U32, idx, 1, dst)
    want = np.bincount(idx, minlength=1 << 16).astype(np.uint32)
    assert np.array_equal(dst, want)
    assert np.array_equal(oracle.histogram_u32_mt(idx, 1 << 16), want)
    # min / max / or with a value buffer
    vals = rng.integers(0, 2**32, size=idx.size, dtype=np.uint64).astype(np.uint32)
    dmax = np.zeros(1 << 16, dtype=np.uint32)
    oracle.scatter_reduce(MAX, U32, idx, vals, dmax)
    want_max = np.zeros(1 << 16, dtype=np.uint32)
    np.maximum.at(want_max, idx, vals)
    assert np.array_equal(dmax, want_max)
    with pytest.raises(NotImplementedError):  # Prod is todo!() in the reference
        oracle.scatter_reduce(PROD, U32, idx, 1, dst)


def test_gather():
    rng = np.random.Generator(np.random.PCG64(13))
    src = rng.random(1000, dtype=np.float32)
    idx = rng.integers(0, 1000, size=5000).astype(np.uint32)
    assert np.array_equal(oracle.gather(src, idx), src[idx])


def test_dynamic_index_kat_via_oracle_ops():
    """dynamic_index (test.rs:976-1019) expressed with the oracle's device ops."""
    c = KATS["dynamic_index"]
    rng = np.random.Generator(np.random.PCG64(0))
    src = rng.integers(c["lo"], c["hi"], size=c["n"]).astype(np.int32)
    cnt, idx = oracle.compress((src < c["max"]).astype(np.uint8))
    vals = oracle.gather(src, idx[:cnt])
    cnt2, idx2 = oracle.compress((vals > c["min"]).astype(np.uint8))
    vals2 = oracle.gather(vals, idx2[:cnt2])
    ref = src[(src > c["min"]) & (src < c["max"])]
    assert np.array_equal(vals2, ref)


def test_c2_chain_matches_numpy_f64():
    rng = np.random.Generator(np.random.PCG64(0))
    x = (rng.random(100000, dtype=np.float32) * 8 - 4).astype(np.float32)
    y = oracle.c2_chain(x)
    t = (x.astype(np.float64) * 1.5 + 0.25).astype(np.float32).astype(np.float64)
    want = np.where(x > 0, np.sin(t), np.exp2(t)).astype(np.float32)
    assert np.array_equal(y, want)
    yf = oracle.c2_chain(x, fast=True)
    assert np.max(np.abs(yf.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-30)) < 1e-6
