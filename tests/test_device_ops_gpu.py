"""GPU parity tests: the hand-written sm_100a kernels, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Integer / index / byte results must be bit-exact; float
reductions and scans must be within the stated relative tolerance of the f64 answer."""
import importlib
import json
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

hj = importlib.import_module("hephaestus-jit_b200")

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_kats.json")) as f:
    KATS = json.load(f)

TY = {"Bool": hj.BOOL, "I8": hj.I8, "U8": hj.U8, "I16": hj.I16, "U16": hj.U16, "I32": hj.I32,
      "U32": hj.U32, "I64": hj.I64, "U64": hj.U64, "F32": hj.F32, "F64": hj.F64}
OP = {"max": hj.MAX, "min": hj.MIN, "sum": hj.SUM, "prod": hj.PROD, "or": hj.OR, "and": hj.AND,
      "xor": hj.XOR}
F32_SUM_RTOL = 1e-5  # BASELINE.json north_star: "1e-5 for f32 sums, accounting for reordering"


@pytest.fixture(scope="module")
def dev():
    return hj.Device.cuda(0)


def gpu_reduce(dev, op, ty, x):
    src = dev.create_buffer_from_slice(x)
    dst = dev.create_buffer(hj.TYPE_SIZE[ty])
    dev.reduce(op, ty, x.size, src, dst)
    return dst.to_host(oracle.NP_DTYPE[ty])[0]


def rand_array(rng, ty, n, small=False):
    dt = oracle.NP_DTYPE[ty]
    if ty in (hj.F32, hj.F64):
        return rng.random(n).astype(dt)
    if ty == hj.BOOL:
        return rng.integers(0, 2, size=n).astype(np.uint8)
    info = np.iinfo(dt)
    if small:
        return rng.integers(0, 4, size=n).astype(dt)
    return rng.integers(info.min, int(info.max) + 1, size=n,
                        dtype=np.int64 if info.min < 0 else np.uint64).astype(dt)


# ---- reduce ------------------------------------------------------------------------------

@pytest.mark.parametrize("case", KATS["reduce"], ids=lambda c: f"{c['op']}-{c['ty']}")
def test_reduce_reference_kats(dev, case):
    ty = TY[case["ty"]]
    lo, hi = case["range"]
    x = np.arange(lo, hi).astype(oracle.NP_DTYPE[ty])
    assert gpu_reduce(dev, OP[case["op"]], ty, x) == case["expect"]


@pytest.mark.parametrize("case", KATS["reduce_random"], ids=lambda c: f"{c['op']}-{c['ty']}")
def test_reduce_reference_random_cases(dev, case):
    ty = TY[case["ty"]]
    rng = np.random.Generator(np.random.PCG64(0))
    x = rng.integers(case["lo"], case["hi"], size=case["n"]).astype(oracle.NP_DTYPE[ty])
    assert gpu_reduce(dev, OP[case["op"]], ty, x) == oracle.reduce(OP[case["op"]], ty, x)[0]


INT_TYPES = [hj.I8, hj.U8, hj.I16, hj.U16, hj.I32, hj.U32, hj.I64, hj.U64]


@pytest.mark.parametrize("ty", INT_TYPES + [hj.BOOL])
@pytest.mark.parametrize("n", [1, 2, 31, 33, 1000, 4097, 65536 + 3, (1 << 20) + 17])
def test_reduce_ints_bit_exact_all_ops(dev, ty, n):
    rng = np.random.Generator(np.random.PCG64(n * 31 + ty))
    x = rand_array(rng, ty, n)
    ops = [hj.AND, hj.OR, hj.XOR] if ty == hj.BOOL else [hj.MAX, hj.MIN, hj.SUM, hj.PROD]
    if ty in (hj.U8, hj.U16, hj.U32, hj.U64):
        ops += [hj.AND, hj.OR, hj.XOR]
    for op in ops:
        if op == hj.PROD:  # keep products non-trivial: odd values never collapse to 0
            xx = (x | 1).astype(x.dtype)
        elif op == hj.AND and ty != hj.BOOL:
            xx = (x | np.array(0x55, dtype=x.dtype)).astype(x.dtype)
        else:
            xx = x
        want = oracle.reduce(op, ty, xx)[0]
        got = gpu_reduce(dev, op, ty, xx)
        assert got == want, (oracle.OP_NAME[op], oracle.TYPE_NAME[ty], n)


@pytest.mark.parametrize("ty,rtol", [(hj.F32, F32_SUM_RTOL), (hj.F64, 1e-12)])
@pytest.mark.parametrize("n", [1, 33, 1000, (1 << 20) + 17, 1 << 24])
def test_reduce_floats(dev, ty, rtol, n):
    rng = np.random.Generator(np.random.PCG64(n))
    x = rng.random(n).astype(oracle.NP_DTYPE[ty])
    exact = float(np.sum(x.astype(np.float64)))
    got = float(gpu_reduce(dev, hj.SUM, ty, x))
    assert abs(got - exact) <= rtol * abs(exact)
    # the reference's own radix-32 tree answer is also within tolerance of ours
    tree = float(oracle.reduce(hj.SUM, ty, x)[0])
    assert abs(got - tree) <= 2 * rtol * abs(exact)
    assert gpu_reduce(dev, hj.MAX, ty, x) == x.max()
    assert gpu_reduce(dev, hj.MIN, ty, x) == x.min()
    xs = (1.0 + (x[: min(n, 4096)] - 0.5) * 1e-3).astype(x.dtype)
    gp = float(gpu_reduce(dev, hj.PROD, ty, xs))
    wp = float(np.prod(xs.astype(np.float64)))
    assert abs(gp - wp) <= 1e-4 * abs(wp)


def test_reduce_f32_integer_valued_is_exact(dev):
    """test.rs:580 — f32 sum over integer-valued data compares with assert_eq."""
    rng = np.random.Generator(np.random.PCG64(0))
    x = rng.integers(0, 100, size=1000).astype(np.float32)
    assert gpu_reduce(dev, hj.SUM, hj.F32, x) == np.float32(x.sum(dtype=np.float64))


def test_reduce_unsupported_pairs_error(dev):
    x = np.zeros(16, np.float16)
    with pytest.raises(hj.HjError) as e:
        gpu_reduce(dev, hj.SUM, hj.F16, x)
    assert e.value.status == 2
    with pytest.raises(hj.HjError):
        gpu_reduce(dev, hj.AND, hj.I32, np.zeros(16, np.int32))
    with pytest.raises(hj.HjError):
        gpu_reduce(dev, hj.SUM, hj.BOOL, np.zeros(16, np.uint8))


def test_reduce_deterministic(dev):
    rng = np.random.Generator(np.random.PCG64(5))
    x = rng.random(1 << 22, dtype=np.float32)
    a = gpu_reduce(dev, hj.SUM, hj.F32, x)
    for _ in range(3):
        assert gpu_reduce(dev, hj.SUM, hj.F32, x) == a


def test_reduce_unaligned_wrapped_buffer(dev):
    """A wrapped, non-16-byte-aligned view exercises the scalar head/tail path."""
    rng = np.random.Generator(np.random.PCG64(6))
    x = rng.integers(0, 2**32, size=100003, dtype=np.uint64).astype(np.uint32)
    whole = dev.create_buffer_from_slice(x)
    view = dev.wrap(whole.ptr + 4, (x.size - 1) * 4)
    dst = dev.create_buffer(4)
    dev.reduce(hj.SUM, hj.U32, x.size - 1, view, dst)
    assert dst.to_host(np.uint32)[0] == oracle.reduce(hj.SUM, hj.U32, x[1:])[0]


# ---- prefix sum --------------------------------------------------------------------------

def gpu_scan(dev, ty, x, inclusive, seed=None):
    src = dev.create_buffer_from_slice(x)
    dst = dev.create_buffer(x.nbytes)
    sb = dev.create_buffer_from_slice(np.array([seed], dtype=x.dtype)) if seed is not None else None
    dev.prefix_sum(ty, x.size, inclusive, src, dst, sb)
    return dst.to_host(x.dtype)


def test_prefix_sum_reference_kat(dev):
    c = KATS["prefix_sum"]
    x = np.arange(c["n"], dtype=np.uint64)
    got = gpu_scan(dev, hj.U64, x, True)
    assert np.array_equal(got, oracle.prefix_sum(hj.U64, x, True))
    assert int(got[-1]) == c["expect_last"]


@pytest.mark.parametrize("ty", INT_TYPES)
@pytest.mark.parametrize("n", [1, 3, 4095, 4096, 4097, 8195, 100003, (1 << 20) + 5])
@pytest.mark.parametrize("inclusive", [True, False])
def test_prefix_sum_ints_bit_exact(dev, ty, n, inclusive):
    rng = np.random.Generator(np.random.PCG64(n + ty))
    x = rand_array(rng, ty, n)
    got = gpu_scan(dev, ty, x, inclusive)
    assert np.array_equal(got, oracle.prefix_sum(ty, x, inclusive))


def test_prefix_sum_large_u32_and_seed(dev):
    n = (1 << 24) + 123
    rng = np.random.Generator(np.random.PCG64(1))
    x = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    want = oracle.prefix_sum_u32_mt(x, True)
    assert np.array_equal(gpu_scan(dev, hj.U32, x, True), want)
    seeded = gpu_scan(dev, hj.U32, x, False, seed=12345)
    with np.errstate(over="ignore"):
        assert np.array_equal(seeded, oracle.prefix_sum_u32_mt(x, False) + np.uint32(12345))


def test_prefix_sum_repeated_launches_epoch(dev):
    """Back-to-back scans of different sizes share the epoch-tagged scratch."""
    rng = np.random.Generator(np.random.PCG64(2))
    for n in [5000, 1 << 18, 777, (1 << 20) + 1, 4096, 1 << 18]:
        x = rng.integers(0, 4, size=n).astype(np.uint32)
        assert np.array_equal(gpu_scan(dev, hj.U32, x, True), np.cumsum(x, dtype=np.uint32))


@pytest.mark.parametrize("ty,rtol", [(hj.F32, 1e-5), (hj.F64, 1e-12)])
@pytest.mark.parametrize("inclusive", [True, False])
def test_prefix_sum_floats(dev, ty, rtol, inclusive):
    n = (1 << 20) + 7
    rng = np.random.Generator(np.random.PCG64(3))
    x = rng.random(n).astype(oracle.NP_DTYPE[ty])
    got = gpu_scan(dev, ty, x, inclusive).astype(np.float64)
    inc = np.cumsum(x.astype(np.float64))
    want = inc if inclusive else np.concatenate([[0.0], inc[:-1]])
    err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    assert err.max() < rtol


def test_prefix_sum_ref_compat_bench_invariant(dev, monkeypatch):
    """benches/vulkan.rs:130-137 under HJ_REF_COMPAT=1 (reference defect D10)."""
    x = np.ones(1 << 16, dtype=np.uint32)
    assert int(gpu_scan(dev, hj.U32, x, False)[-1]) == x.size - 1
    monkeypatch.setenv("HJ_REF_COMPAT", "1")
    assert int(gpu_scan(dev, hj.U32, x, False)[-1]) == x.size


# ---- compress ----------------------------------------------------------------------------

def gpu_compress(dev, mask, index_base=0, size=None, sentinel=0):
    n = mask.size
    src = dev.create_buffer_from_slice(mask)
    idx = dev.create_buffer_from_slice(np.full(n, sentinel, dtype=np.uint32))
    cnt = dev.create_buffer_from_slice(np.zeros(1, np.uint32))
    sb = dev.create_buffer_from_slice(np.array([size], np.uint32)) if size is not None else None
    dev.compress(n, cnt, src, idx, size_buf=sb, index_base=index_base)
    return int(cnt.to_host(np.uint32)[0]), idx.to_host(np.uint32)


def test_compress_reference_kats(dev):
    n = KATS["compress_all_true"]["n"]
    count, idx = gpu_compress(dev, np.ones(n, np.uint8))
    assert count == n and np.array_equal(idx, np.arange(n, dtype=np.uint32))
    rng = np.random.Generator(np.random.PCG64(0))
    m = rng.integers(0, 2, size=KATS["compress_small"]["n"]).astype(np.uint8)
    count, idx = gpu_compress(dev, m)
    wc, wi = oracle.compress(m)
    assert count == wc and np.array_equal(idx, wi)


@pytest.mark.parametrize("n", [1, 15, 17, 2049, 4111, 8191, 8192, 8193, 100003, (1 << 20) + 9])
@pytest.mark.parametrize("p", [0.0, 0.01, 0.5, 0.99, 1.0])
def test_compress_bit_exact(dev, n, p):
    rng = np.random.Generator(np.random.PCG64(n))
    mask = (rng.random(n) < p).astype(np.uint8)
    count, idx = gpu_compress(dev, mask, index_base=5, sentinel=0xDEADBEEF)
    wc, wi = oracle.compress(mask, index_out=np.full(n, 0xDEADBEEF, np.uint32), index_base=5)
    assert count == wc
    assert np.array_equal(idx, wi)  # includes the untouched tail


def test_compress_mixed_density(dev):
    # density changes every 1000 elements so sparse, medium and dense rows (and slices) meet in one
    # launch of the ring kernel
    n = (1 << 22) + 1234
    rng = np.random.Generator(np.random.PCG64(11))
    dens = rng.choice([0.0, 0.002, 0.02, 0.1, 0.13, 0.3, 0.8, 0.82, 0.97, 1.0], size=n // 1000 + 1)
    mask = (rng.random(n) < np.repeat(dens, 1000)[:n]).astype(np.uint8)
    count, idx = gpu_compress(dev, mask, index_base=3, sentinel=0xDEADBEEF)
    wc, wi = oracle.compress(mask, index_out=np.full(n, 0xDEADBEEF, np.uint32), index_base=3, mt=True)
    assert count == wc and np.array_equal(idx, wi)
    # Outside the reference's contract (its rank is the running sum of the mask BYTES, so it only
    # defines 0/1 masks, compress_large.glsl:126-132): this backend selects every non-zero byte.
    vals = rng.choice(np.array([1, 1, 1, 2, 128, 255], np.uint8), size=n)
    wide = np.where(mask != 0, vals, 0).astype(np.uint8)
    count2, idx2 = gpu_compress(dev, wide, index_base=3, sentinel=0xDEADBEEF)
    assert count2 == wc and np.array_equal(idx2, wi)


def test_compress_large_and_dynsize(dev):
    n = (1 << 24) + 77
    rng = np.random.Generator(np.random.PCG64(4))
    mask = (rng.random(n) < 0.5).astype(np.uint8)
    count, idx = gpu_compress(dev, mask)
    wc, wi = oracle.compress(mask, mt=True)
    assert count == wc and np.array_equal(idx, wi)
    # DynSize: only the first `size` elements take part
    size = 1_000_003
    count, idx = gpu_compress(dev, mask, size=size)
    wc, wi = oracle.compress(mask[:size], index_out=np.zeros(n, np.uint32))
    assert count == wc and np.array_equal(idx, wi)


# ---- scatter-reduce / gather ----------------------------------------------------------------

def test_scatter_reduce_reference_kat(dev):
    c = KATS["scatter_reduce"]
    dst = dev.create_buffer_from_slice(np.array(c["dst"], np.uint32))
    idx = dev.create_buffer_from_slice(np.full(c["n"], c["idx"], np.uint32))
    dev.scatter_reduce(hj.SUM, hj.U32, c["n"], idx, None, c["value"], dst, 3)
    assert dst.to_host(np.uint32).tolist() == c["expect"]


@pytest.mark.parametrize("n_bins", [16, 1 << 12, 1 << 16])
def test_histogram_bit_exact(dev, n_bins):
    n = (1 << 22) + 3
    rng = np.random.Generator(np.random.PCG64(n_bins))
    keys = rng.integers(0, n_bins, size=n).astype(np.uint32)
    dst = dev.create_buffer_from_slice(np.zeros(n_bins, np.uint32))
    idx = dev.create_buffer_from_slice(keys)
    dev.scatter_reduce(hj.SUM, hj.U32, n, idx, None, 1, dst, n_bins)
    got = dst.to_host(np.uint32)
    assert np.array_equal(got, oracle.histogram_u32_mt(keys, n_bins))
    assert int(got.sum(dtype=np.uint64)) == n


@pytest.mark.parametrize("n_bins,literal,dist", [
    (1 << 16, 1, "hot"),       # packed 16-bit counters: hot bins cross 0x8000 and are cashed into HBM
    (1 << 16, 3, "uniform"),   # literal != 1: u32 bins in two windows
    (100_000, 1, "uniform"),   # more than 2^16 bins: four windows, ragged last window
    (40_000, 1, "oob"),        # keys >= n_dst are ignored
    (1 << 15, 1, "uniform"),   # one window holding every bin
    (40_001, 1, "oob"),        # odd bin count, packed counters: the unused upper half of the last word takes key == n_dst
    (50_001, 1, "edge"),       # every key is n_dst - 1 or n_dst: the last real counter and its out-of-range neighbour
    (99_999, 5, "oob"),        # odd windows, literal != 1, half the keys out of range (dummy words behind each window)
    (1000, 1, "oob"),          # 32 copies of every bin (one per lane), dummy row for the keys out of range
    (1024, 3, "hot"),          # 32 copies, 1024 bins + dummy row exactly fill the window
    (3000, 1, "edge"),         # 8 copies
    (16_383, 2, "uniform"),    # 2 copies
])
def test_histogram_ring_paths_bit_exact(dev, n_bins, literal, dist):
    n = (1 << 22) + 12345
    rng = np.random.Generator(np.random.PCG64(n_bins + literal))
    if dist == "hot":
        keys = np.where(rng.random(n) < 0.6, rng.integers(0, 4, size=n), rng.integers(0, n_bins, size=n)).astype(np.uint32)
    elif dist == "oob":
        keys = rng.integers(0, 2 * n_bins, size=n).astype(np.uint32)
        keys[::1001] = 0xFFFFFFFF
    elif dist == "edge":
        keys = np.where(rng.random(n) < 0.5, n_bins - 1, n_bins).astype(np.uint32)
    else:
        keys = rng.integers(0, n_bins, size=n).astype(np.uint32)
    init = rng.integers(0, 1000, size=n_bins).astype(np.uint32)
    dst = dev.create_buffer_from_slice(init)
    dev.scatter_reduce(hj.SUM, hj.U32, n, dev.create_buffer_from_slice(keys), None, literal, dst, n_bins)
    inside = keys[keys < n_bins]
    want = init + (oracle.histogram_u32_mt(inside, n_bins) * np.uint32(literal))
    assert np.array_equal(dst.to_host(np.uint32), want)


def test_histogram_single_bin_worst_case(dev):
    # every key hits one bin (and its 16-bit neighbour in the same word): each CTA pushes 2^25 / 148
    # adds through ONE packed counter, several sweep periods' worth — no counter may wrap or carry
    n, n_bins = 1 << 25, 1 << 16
    keys = np.full(n, 40001, np.uint32)
    keys[1::3] = 40000
    init = np.arange(n_bins, dtype=np.uint32)
    dst = dev.create_buffer_from_slice(init)
    dev.scatter_reduce(hj.SUM, hj.U32, n, dev.create_buffer_from_slice(keys), None, 1, dst, n_bins)
    want = init.copy()
    want[40000] += np.uint32(np.count_nonzero(keys == 40000))
    want[40001] += np.uint32(np.count_nonzero(keys == 40001))
    assert np.array_equal(dst.to_host(np.uint32), want)


@pytest.mark.parametrize("n_bins,literal,dist", [
    (1 << 18, 1, "uniform"),     # 8 windows of 32 Ki bins: groups of 8 CTAs walk the same key tiles
    (300_001, 3, "oob"),         # beyond 8 windows: one L2 atomic per key; literal != 1, half the keys out of range
    (1 << 20, 1, "hot"),         # 60 % of the keys in four bins: hot addresses in L2
    (5_000_000, 1, "uniform"),
    (200_000, 1, "edge"),        # every key is n_dst - 1 or n_dst
])
def test_histogram_many_bins(dev, n_bins, literal, dist):
    n = (1 << 22) + 12345
    rng = np.random.Generator(np.random.PCG64(n_bins + literal))
    if dist == "hot":
        keys = np.where(rng.random(n) < 0.6, rng.integers(0, 4, size=n), rng.integers(0, n_bins, size=n)).astype(np.uint32)
    elif dist == "oob":
        keys = rng.integers(0, 2 * n_bins, size=n).astype(np.uint32)
        keys[::1001] = 0xFFFFFFFF
    elif dist == "edge":
        keys = np.where(rng.random(n) < 0.5, n_bins - 1, n_bins).astype(np.uint32)
    else:
        keys = rng.integers(0, n_bins, size=n).astype(np.uint32)
    init = rng.integers(0, 1000, size=n_bins).astype(np.uint32)
    for rep in range(2):
        dst = dev.create_buffer_from_slice(init)
        dev.scatter_reduce(hj.SUM, hj.U32, n, dev.create_buffer_from_slice(keys), None, literal, dst, n_bins)
        inside = keys[keys < n_bins]
        want = init + (np.bincount(inside, minlength=n_bins).astype(np.uint32) * np.uint32(literal))
        assert np.array_equal(dst.to_host(np.uint32), want), rep


@pytest.mark.parametrize("dist", ["uniform", "one_bin", "oob", "hot_tail"])
def test_histogram_adaptive_sweeping(dev, dist):
    """The packed-16 histogram decides after its first 64512 keys per CTA whether the 16-bit counters need
    their overflow sweeps, and a CTA that ran without them proves afterwards that no counter wrapped
    (sum of its counters == keys applied) or walks its tiles again with the sweeps on.
    uniform: no sweeps, the proof holds; one_bin: the sweeps are switched on; hot_tail: the projection is
    wrong (the keys turn hot late), counters wrap, the second walk must give the exact result; oob: most
    keys are out of range and land on the lanes' dummy counters (which may wrap too: a false alarm)."""
    n, n_bins = (1 << 26) + 4099, 1 << 16
    rng = np.random.Generator(np.random.PCG64(len(dist)))
    if dist == "uniform":
        keys = rng.integers(0, n_bins, size=n, dtype=np.uint32)
    elif dist == "one_bin":
        keys = np.full(n, 40001, np.uint32)
        keys[1::3] = 40000
    elif dist == "oob":
        keys = rng.integers(0, 16 * n_bins, size=n, dtype=np.uint32)
    else:  # uniform, except the last third of the keys: > 65535 late keys per CTA in a single bin
        keys = rng.integers(0, n_bins, size=n, dtype=np.uint32)
        keys[-(n // 3):] = 7
    init = rng.integers(0, 1000, size=n_bins).astype(np.uint32)
    for rep in range(2):
        dst = dev.create_buffer_from_slice(init)
        dev.scatter_reduce(hj.SUM, hj.U32, n, dev.create_buffer_from_slice(keys), None, 1, dst, n_bins)
        want = init + np.bincount(keys[keys < n_bins], minlength=n_bins).astype(np.uint32)
        assert np.array_equal(dst.to_host(np.uint32), want), (dist, rep)


@pytest.mark.parametrize("op", [hj.MAX, hj.MIN, hj.OR, hj.AND, hj.XOR, hj.SUM])
def test_scatter_reduce_values_bit_exact(dev, op):
    n, n_bins = 200003, 1000
    rng = np.random.Generator(np.random.PCG64(op))
    keys = rng.integers(0, n_bins, size=n).astype(np.uint32)
    vals = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    init = rng.integers(0, 2**32, size=n_bins, dtype=np.uint64).astype(np.uint32)
    dst = dev.create_buffer_from_slice(init)
    dev.scatter_reduce(op, hj.U32, n, dev.create_buffer_from_slice(keys),
                       dev.create_buffer_from_slice(vals), 0, dst, n_bins)
    want = oracle.scatter_reduce(op, hj.U32, keys, vals, init.copy())
    assert np.array_equal(dst.to_host(np.uint32), want)


def test_scatter_reduce_f32_sum_tolerance(dev):
    n, n_bins = 1 << 20, 64
    rng = np.random.Generator(np.random.PCG64(9))
    keys = rng.integers(0, n_bins, size=n).astype(np.uint32)
    vals = rng.random(n, dtype=np.float32)
    dst = dev.create_buffer_from_slice(np.zeros(n_bins, np.float32))
    dev.scatter_reduce(hj.SUM, hj.F32, n, dev.create_buffer_from_slice(keys),
                       dev.create_buffer_from_slice(vals), 0.0, dst, n_bins)
    want = np.zeros(n_bins, np.float64)
    np.add.at(want, keys, vals.astype(np.float64))
    assert np.allclose(dst.to_host(np.float32), want, rtol=1e-4)


def test_gather_bit_exact(dev):
    rng = np.random.Generator(np.random.PCG64(10))
    for dt in (np.uint8, np.uint16, np.float32, np.float64):
        src = (rng.random(5000) * 200).astype(dt)
        idx = rng.integers(0, src.size, size=100001).astype(np.uint32)
        out = dev.create_buffer(idx.size * src.itemsize)
        dev.gather(src.itemsize, idx.size, dev.create_buffer_from_slice(src),
                   dev.create_buffer_from_slice(idx), out)
        assert np.array_equal(out.to_host(dt), oracle.gather(src, idx))


# ---- runtime -------------------------------------------------------------------------------

def test_buffer_roundtrip_and_ranges(dev):
    x = np.arange(1000, dtype=np.int32)
    b = dev.create_buffer_from_slice(x)
    assert np.array_equal(b.to_host(np.int32), x)
    assert np.array_equal(b.to_host(np.int32, 10, 20), x[10:20])
    with pytest.raises(hj.HjError):
        b.to_host(np.int32, 0, 1001)
    assert dev.info()["sm_count"] > 0
    assert dev.launch_count() > 0


# ---- BASELINE.json config 1, exactly as SURVEY.md §8d states it --------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_baseline_config_c1(dev, seed):
    """C1: u32 exclusive prefix sum + f32 sum-reduce over 2^20 elements (the reference's own
    CPU-runnable case).  Inputs per SURVEY §8d: numpy PCG64(seed); u32 uniform in [0, 16) for the scan
    (bit-exact); f32 uniform [0, 1) for the sum (rel 1e-5 vs f64) plus an integer-valued f32 set in
    [0, 100) that must be exact (mirrors test.rs:580)."""
    n = 1 << 20
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.integers(0, 16, size=n).astype(np.uint32)
    src, dst = dev.create_buffer_from_slice(u), dev.create_buffer(4 * n)
    dev.prefix_sum(hj.U32, n, False, src, dst)
    assert np.array_equal(dst.to_host(np.uint32), oracle.prefix_sum(oracle.U32, u, False))
    f = rng.random(n, dtype=np.float32)
    got = gpu_reduce(dev, hj.SUM, hj.F32, f)
    exact = float(f.astype(np.float64).sum())
    assert abs(float(got) - exact) <= F32_SUM_RTOL * exact
    assert abs(float(oracle.reduce(oracle.SUM, oracle.F32, f)[0]) - exact) <= F32_SUM_RTOL * exact  # the oracle's tree order too
    fi = rng.integers(0, 100, size=n).astype(np.float32)   # sum < 2^27 is not exactly representable step by step ...
    got_i = gpu_reduce(dev, hj.SUM, hj.F32, fi[:1000])       # ... so the exact case keeps the reference's 1000 elements
    assert float(got_i) == float(fi[:1000].astype(np.float64).sum())


# ---- the device ops over HOST arrays (hj_reduce_host / hj_prefix_sum_host / hj_compress_host) ---------------
@pytest.mark.parametrize("n,chunk", [(1, 0), (1000, 64), ((1 << 20) + 7, 1 << 16), ((1 << 22) + 3, 0), (5_000_001, 1 << 19)])
def test_host_streamed_ops_match_oracle(dev, n, chunk):
    rng = np.random.Generator(np.random.PCG64(n))
    u = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    f = rng.random(n, dtype=np.float32)
    # reduce: chunk partials folded on the device
    for op in (hj.SUM, hj.MAX, hj.XOR):
        assert dev.reduce_host(op, hj.U32, n, u, chunk) == oracle.reduce(op, oracle.U32, u)[0]
    exact = float(f.astype(np.float64).sum())
    assert abs(float(dev.reduce_host(hj.SUM, hj.F32, n, f, chunk)) - exact) <= 1e-5 * max(exact, 1.0)
    assert dev.reduce_host(hj.MIN, hj.F32, n, f, chunk) == f.min()
    # scan: the running total is the seed of the next chunk
    out = np.empty(n, np.uint32)
    for inclusive in (True, False):
        dev.prefix_sum_host(hj.U32, n, inclusive, u, out, chunk)
        assert np.array_equal(out, oracle.prefix_sum(oracle.U32, u, inclusive))
    u64 = rng.integers(0, 2**63, size=n, dtype=np.uint64)
    out64 = np.empty(n, np.uint64)
    dev.prefix_sum_host(hj.U64, n, True, u64, out64, chunk)
    assert np.array_equal(out64, np.cumsum(u64, dtype=np.uint64))
    # compress: indices of every chunk behind those of the chunks before it; the tail is not touched
    for p in (0.5, 0.02, 1.0, 0.0):
        mask = (rng.random(n) < p).astype(np.uint8)
        idx = np.full(n, 0xDEADBEEF, np.uint32)
        cnt = dev.compress_host(n, mask, idx, 0, chunk)
        want_cnt, want = oracle.compress(mask)
        assert cnt == want_cnt
        assert np.array_equal(idx[:cnt], want[:cnt]) and (idx[cnt:] == 0xDEADBEEF).all()
    idx = np.zeros(n, np.uint32)
    mask = (rng.random(n) < 0.3).astype(np.uint8)
    cnt = dev.compress_host(n, mask, idx, 1000, chunk)
    assert np.array_equal(idx[:cnt], np.flatnonzero(mask).astype(np.uint32) + 1000)


def test_host_streamed_ops_from_pinned_memory(dev):
    """The same through pinned staging memory (hj_host_alloc), the configuration bench.py times."""
    import ctypes
    L = importlib.import_module("hephaestus-jit_b200._lib")
    n = (1 << 23) + 11
    src, dst = ctypes.c_void_p(), ctypes.c_void_p()
    L.check(L.lib.hj_host_alloc(4 * n, ctypes.byref(src)))
    L.check(L.lib.hj_host_alloc(4 * n, ctypes.byref(dst)))
    try:
        a = np.ctypeslib.as_array(ctypes.cast(src, ctypes.POINTER(ctypes.c_uint32)), shape=(n,))
        b = np.ctypeslib.as_array(ctypes.cast(dst, ctypes.POINTER(ctypes.c_uint32)), shape=(n,))
        a[:] = np.random.Generator(np.random.PCG64(1)).integers(0, 4, size=n).astype(np.uint32)
        dev.prefix_sum_host(hj.U32, n, True, src.value, dst.value)
        assert np.array_equal(b, np.cumsum(a, dtype=np.uint32))
        assert dev.reduce_host(hj.SUM, hj.U32, n, src.value) == np.uint32(a.sum(dtype=np.uint64) & 0xFFFFFFFF)
    finally:
        L.lib.hj_host_free(src)
        L.lib.hj_host_free(dst)


def test_host_streamed_ops_from_concurrent_threads(dev):
    """Four host threads stream four different arrays at once (the library only holds its device lock
    while a chunk is being enqueued): every result must be what the op gives alone."""
    import threading
    irm = importlib.import_module("hephaestus-jit_b200.ir")
    n = (1 << 22) + 123
    rng = np.random.Generator(np.random.PCG64(9))
    u = rng.integers(0, 1 << 16, size=n).astype(np.uint32)
    f = rng.random(n, dtype=np.float32)
    x = (rng.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
    mask = (rng.random(n) < 0.4).astype(np.uint8)
    scan, idx, y = np.empty(n, np.uint32), np.zeros(n, np.uint32), np.empty(n, np.float32)
    kern = dev.kernel(irm.c2_chain_ir())
    res, errs = {}, []

    def guarded(fn):
        def run():
            try:
                for _ in range(3):
                    fn()
            except BaseException as exc:  # noqa: BLE001
                errs.append(repr(exc))
        return run

    jobs = [lambda: dev.prefix_sum_host(hj.U32, n, True, u, scan, 1 << 18),
            lambda: res.__setitem__("cnt", dev.compress_host(n, mask, idx, 0, 1 << 18)),
            lambda: dev.map_host(kern, n, [x, y], 1 << 18),
            lambda: res.__setitem__("sum", dev.reduce_host(hj.SUM, hj.F32, n, f, 1 << 18))]
    ts = [threading.Thread(target=guarded(j)) for j in jobs]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errs, errs
    assert np.array_equal(scan, np.cumsum(u, dtype=np.uint32))
    want_cnt, want_idx = oracle.compress(mask)
    assert res["cnt"] == want_cnt and np.array_equal(idx[:want_cnt], want_idx[:want_cnt])
    assert np.allclose(y, oracle.c2_chain(x), rtol=4e-7, atol=1e-7)
    exact = float(f.astype(np.float64).sum())
    assert abs(float(res["sum"]) - exact) <= 1e-5 * exact
