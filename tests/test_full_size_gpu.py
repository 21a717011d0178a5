"""GPU tests at BASELINE.json's full sizes (2^30 / 2^28 elements) through size-independent
properties — the oracle cannot re-compute these sizes in seconds, so each test checks an invariant
the domain offers (the reference's own bench asserts, benches/vulkan.rs:107-137,190, are of this
kind): scan of ones ends in n, inclusive/exclusive differ by the input, linearity of reductions,
compaction of a periodic mask is an arithmetic progression, a histogram sums to the key count."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

hj = importlib.import_module("hephaestus-jit_b200")
torch = pytest.importorskip("torch")

N30 = 1 << 30
N28 = 1 << 28


@pytest.fixture(scope="module")
def dev():
    torch.cuda.set_device(0)
    d = hj.Device.cuda(0)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    d.set_stream(s.cuda_stream)  # torch kernels and library kernels share one stream
    yield d
    torch.cuda.synchronize()
    d.set_stream(None)


def wrap(dev, t):
    return dev.wrap(t.data_ptr(), t.numel() * t.element_size())


def test_scan_of_ones_2p30(dev):  # benches/vulkan.rs:130-137: pfs[n-1] == n
    x = torch.ones(N30, device="cuda", dtype=torch.int32)
    y = torch.empty_like(x)
    dev.prefix_sum(hj.U32, N30, True, wrap(dev, x), wrap(dev, y))
    torch.cuda.synchronize()
    assert int(y[-1]) == N30 and int(y[0]) == 1  # 2^30 fits i32 as a positive value
    # every element: y[i] == i + 1, checked on device
    assert bool((y == torch.arange(1, N30 + 1, device="cuda", dtype=torch.int32)).all())
    dev.prefix_sum(hj.U32, N30, False, wrap(dev, x), wrap(dev, y))
    torch.cuda.synchronize()
    assert int(y[0]) == 0 and int(y[-1]) == N30 - 1


def test_scan_inclusive_minus_exclusive_is_input_2p30(dev):
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randint(-(1 << 31), (1 << 31) - 1, (N30,), device="cuda", generator=g, dtype=torch.int32)  # wraps
    inc = torch.empty_like(x)
    exc = torch.empty_like(x)
    dev.prefix_sum(hj.U32, N30, True, wrap(dev, x), wrap(dev, inc))
    dev.prefix_sum(hj.U32, N30, False, wrap(dev, x), wrap(dev, exc))
    torch.cuda.synchronize()
    assert bool(((inc - exc) == x).all())           # wrapping arithmetic on both sides
    assert bool((exc[1:] == inc[:-1]).all())
    # total against a 64-bit reduction of the same data
    total = int(x.to(torch.int64).sum().item()) & 0xFFFFFFFF
    assert (int(inc[-1].item()) & 0xFFFFFFFF) == total


def test_reduce_linearity_and_extrema_2p30(dev):
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randint(0, 1 << 31, (N30,), device="cuda", generator=g, dtype=torch.int32)
    out = torch.zeros(4, device="cuda", dtype=torch.int32)
    bo = wrap(dev, out)
    half = N30 // 2
    parts = []
    for lo, n in ((0, N30), (0, half), (half, half)):
        dev.reduce(hj.SUM, hj.U32, n, wrap(dev, x[lo:lo + n]), bo)
        torch.cuda.synchronize()
        parts.append(int(out[0].item()) & 0xFFFFFFFF)
    assert parts[0] == (parts[1] + parts[2]) & 0xFFFFFFFF            # sum of halves, mod 2^32
    assert parts[0] == int(x.to(torch.int64).sum().item()) & 0xFFFFFFFF
    x[123456789] = (1 << 31) - 1
    x[987654321] = 0
    dev.reduce(hj.MAX, hj.U32, N30, wrap(dev, x), bo)
    torch.cuda.synchronize()
    assert int(out[0].item()) == (1 << 31) - 1
    dev.reduce(hj.MIN, hj.U32, N30, wrap(dev, x), bo)
    torch.cuda.synchronize()
    assert int(out[0].item()) == 0


def test_reduce_f32_sum_tolerance_2p30(dev):
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(N30, device="cuda", generator=g, dtype=torch.float32)
    out = torch.zeros(4, device="cuda", dtype=torch.float32)
    dev.reduce(hj.SUM, hj.F32, N30, wrap(dev, x), wrap(dev, out))
    torch.cuda.synchronize()
    exact = float(x.to(torch.float64).sum().item())
    assert abs(float(out[0].item()) - exact) <= 1e-5 * exact  # BASELINE north star: rel 1e-5


def test_compress_periodic_mask_2p30(dev):
    # every third element selected: the indices are the arithmetic progression 0, 3, 6, ...
    idx_all = torch.arange(N30, device="cuda", dtype=torch.int32)
    mask = (idx_all % 3 == 0).to(torch.uint8)
    del idx_all
    out = torch.zeros(N30, device="cuda", dtype=torch.int32)
    cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
    dev.compress(N30, wrap(dev, cnt), wrap(dev, mask), wrap(dev, out))
    torch.cuda.synchronize()
    c = int(cnt.item())
    assert c == (N30 + 2) // 3
    want = torch.arange(0, c, device="cuda", dtype=torch.int32) * 3
    assert bool((out[:c] == want).all()) and bool((out[c:] == 0).all())
    # all true: count == n (benches/vulkan.rs:107-114) and the identity permutation
    mask.fill_(1)
    dev.compress(N30, wrap(dev, cnt), wrap(dev, mask), wrap(dev, out))
    torch.cuda.synchronize()
    assert int(cnt.item()) == N30 and int(out[-1].item()) == N30 - 1 and int(out[N30 // 2].item()) == N30 // 2


def test_compress_random_mask_matches_nonzero_2p28(dev):
    g = torch.Generator(device="cuda").manual_seed(8)
    for p in (0.5, 0.01):
        mask = (torch.rand(N28, device="cuda", generator=g) < p).to(torch.uint8)
        out = torch.zeros(N28, device="cuda", dtype=torch.int32)
        cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
        dev.compress(N28, wrap(dev, cnt), wrap(dev, mask), wrap(dev, out))
        torch.cuda.synchronize()
        want = torch.nonzero(mask).flatten().to(torch.int32)
        c = int(cnt.item())
        assert c == want.numel() and bool((out[:c] == want).all()) and bool((out[c:] == 0).all())
        del mask, out, want


def test_histogram_2p28_keys_2p16_bins(dev):
    g = torch.Generator(device="cuda").manual_seed(9)
    keys = torch.randint(0, 1 << 16, (N28,), device="cuda", generator=g, dtype=torch.int32)
    keys[: N28 // 8] = 4242  # one hot bin: 2^25 hits cross the 16-bit counters 3 times per CTA
    hist = torch.zeros(1 << 16, device="cuda", dtype=torch.int32)
    dev.scatter_reduce(hj.SUM, hj.U32, N28, wrap(dev, keys), None, 1, wrap(dev, hist), 1 << 16)
    torch.cuda.synchronize()
    assert int(hist.to(torch.int64).sum().item()) == N28
    want = torch.bincount(keys.to(torch.int64), minlength=1 << 16)
    assert bool((hist.to(torch.int64) == want).all())


def test_fused_chain_2p28_matches_torch(dev):
    irm = importlib.import_module("hephaestus-jit_b200.ir")
    g = torch.Generator(device="cuda").manual_seed(10)
    x = torch.rand(N28, device="cuda", generator=g, dtype=torch.float32) * 8 - 4
    y = torch.empty_like(x)
    dev.execute_graph([{"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": irm.c2_chain_ir(), "size": N28}],
                      [wrap(dev, x), wrap(dev, y)], [(N28, hj.F32, 4), (N28, hj.F32, 4)])
    torch.cuda.synchronize()
    t = torch.addcmul(torch.full_like(x, 0.25), x, torch.full_like(x, 1.5))  # not fused: compare loosely
    want = torch.where(x > 0, torch.sin(t), torch.exp2(t))
    # torch evaluates x*1.5+0.25 with two roundings; fma has one: |t| <= 6.25 -> |dt| <= 2^-22
    assert float((y - want).abs().max().item()) <= 6e-6
