"""The largest extents the u32 index type allows (trace.rs:552-562: sizes must fit u32): n = 2^32 - 1, which is also
odd (every vector / tile path ends ragged), and n just above 2^31 (where a signed 32-bit index would wrap).
Checked through invariants, like tests/test_full_size_gpu.py; 2^32 itself must be rejected, not truncated."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

hj = importlib.import_module("hephaestus-jit_b200")
irm = importlib.import_module("hephaestus-jit_b200.ir")
torch = pytest.importorskip("torch")

NMAX = (1 << 32) - 1
NBIG = (1 << 31) + 12345


@pytest.fixture(scope="module")
def dev():
    torch.cuda.set_device(0)
    free, _total = torch.cuda.mem_get_info()
    if free < 64 * (1 << 30):
        pytest.skip("needs 64 GiB of free device memory")
    d = hj.Device.cuda(0)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    d.set_stream(s.cuda_stream)
    yield d
    torch.cuda.synchronize()
    d.set_stream(None)
    torch.cuda.empty_cache()


@pytest.fixture(autouse=True)
def free_memory():
    yield
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def wrap(dev, t):
    return dev.wrap(t.data_ptr(), t.numel() * t.element_size())


def u32(t):  # torch has no uint32 arithmetic: compare as int64
    return t.to(torch.int64) & 0xFFFFFFFF


def as_i32(v):  # int64 values in [0, 2^32) -> the int32 with the same bit pattern
    return torch.where(v >= (1 << 31), v - (1 << 32), v).to(torch.int32)


def check_ramp(y, first, n, chunk=1 << 28):
    """y[i] == (first + i) mod 2^32 for all i < n, chunk by chunk (an arange of 2^32 int64 would be 32 GiB)."""
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        want = (torch.arange(lo, hi, device="cuda", dtype=torch.int64) + first) & 0xFFFFFFFF
        assert bool((u32(y[lo:hi]) == want).all()), f"mismatch in [{lo}, {hi})"


@pytest.mark.parametrize("n", [NMAX, NBIG])
def test_scan_of_ones(dev, n):
    x = torch.ones(n, device="cuda", dtype=torch.int32)
    y = torch.empty_like(x)
    dev.prefix_sum(hj.U32, n, True, wrap(dev, x), wrap(dev, y))
    torch.cuda.synchronize()
    assert int(u32(y[-1:]).item()) == n & 0xFFFFFFFF
    check_ramp(y, 1, n)
    dev.prefix_sum(hj.U32, n, False, wrap(dev, x), wrap(dev, y))
    torch.cuda.synchronize()
    check_ramp(y, 0, n)


@pytest.mark.parametrize("n", [NMAX, NBIG])
def test_reduce_sum_and_extrema(dev, n):
    x = torch.ones(n, device="cuda", dtype=torch.int32)
    out = torch.zeros(4, device="cuda", dtype=torch.int32)
    dev.reduce(hj.SUM, hj.U32, n, wrap(dev, x), wrap(dev, out))
    assert int(u32(out[:1]).item()) == n & 0xFFFFFFFF
    x[n - 1] = 77          # the very last element must take part
    x[(1 << 31) + 1] = -5  # as u32: 4294967291
    dev.reduce(hj.MAX, hj.U32, n, wrap(dev, x), wrap(dev, out))
    assert int(u32(out[:1]).item()) == 4294967291
    x[(1 << 31) + 1] = 1
    dev.reduce(hj.MAX, hj.U32, n, wrap(dev, x), wrap(dev, out))
    assert int(u32(out[:1]).item()) == 77
    x[n - 1] = 0
    dev.reduce(hj.MIN, hj.U32, n, wrap(dev, x), wrap(dev, out))
    assert int(u32(out[:1]).item()) == 0


@pytest.mark.parametrize("n", [NMAX, NBIG])
def test_compress_every_third_and_all(dev, n):
    mask = torch.zeros(n, device="cuda", dtype=torch.uint8)
    mask[::3] = 1
    out = torch.empty(n, device="cuda", dtype=torch.int32)
    cnt = torch.zeros(4, device="cuda", dtype=torch.int32)
    dev.compress(n, wrap(dev, cnt), wrap(dev, mask), wrap(dev, out))
    torch.cuda.synchronize()
    c = int(u32(cnt[:1]).item())
    assert c == (n + 2) // 3
    for lo in range(0, c, 1 << 28):  # out[k] == 3k, also beyond 2^31
        hi = min(c, lo + (1 << 28))
        assert bool((u32(out[lo:hi]) == 3 * torch.arange(lo, hi, device="cuda", dtype=torch.int64)).all())
    mask.fill_(1)
    dev.compress(n, wrap(dev, cnt), wrap(dev, mask), wrap(dev, out))
    torch.cuda.synchronize()
    assert int(u32(cnt[:1]).item()) == n
    check_ramp(out, 0, n)
    mask.zero_()
    mask[n - 1] = 1  # only the last element
    out.fill_(-1)
    dev.compress(n, wrap(dev, cnt), wrap(dev, mask), wrap(dev, out))
    torch.cuda.synchronize()
    assert int(cnt[0].item()) == 1 and int(u32(out[:1]).item()) == n - 1 and int(out[1].item()) == -1


def test_fused_kernel_sees_every_index_at_the_largest_extent(dev):
    """dst[i] = src[i] + Index over 2^32 - 1 elements: the last index is 2^32 - 2."""
    n = NMAX
    b = irm.IRBuilder()
    t = b.scalar(hj.U32)
    src = b.buffer_ref(t)
    idx = b.index()
    v = b.bop(irm.BOP_ADD, t, b.gather(t, src, idx), idx)
    dst = b.buffer_ref(t)
    b.scatter(dst, v, idx)
    x = torch.full((n,), 5, device="cuda", dtype=torch.int32)
    y = torch.empty_like(x)
    dev.execute_graph([{"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": b, "size": n}], [wrap(dev, x), wrap(dev, y)],
                      [(n, hj.U32, 4), (n, hj.U32, 4)])
    torch.cuda.synchronize()
    check_ramp(y, 5, n)


def test_histogram_of_every_key_at_the_largest_extent(dev):
    n, bins = NMAX, 1 << 16
    keys = torch.empty(n, device="cuda", dtype=torch.int32)
    for lo in range(0, n, 1 << 28):  # keys[i] = i mod 2^16
        hi = min(n, lo + (1 << 28))
        keys[lo:hi] = (torch.arange(lo, hi, device="cuda", dtype=torch.int64) & (bins - 1)).to(torch.int32)
    hist = torch.zeros(bins, device="cuda", dtype=torch.int32)
    dev.scatter_reduce(hj.SUM, hj.U32, n, wrap(dev, keys), None, 1, wrap(dev, hist), bins)
    torch.cuda.synchronize()
    want = torch.full((bins,), n // bins, device="cuda", dtype=torch.int64)
    want[: n % bins] += 1
    assert bool((u32(hist) == want).all())


def test_gather_over_the_largest_extent(dev):
    """dst[i] = table[idx[i]] for 2^32 - 1 indices into a 2^20-entry table: idx[i] = (n - 1 - i) mod 2^20."""
    n, tn = NMAX, 1 << 20
    table = torch.arange(tn, device="cuda", dtype=torch.int32) * 7 + 3
    idx = torch.empty(n, device="cuda", dtype=torch.int32)
    for lo in range(0, n, 1 << 28):
        hi = min(n, lo + (1 << 28))
        idx[lo:hi] = ((n - 1 - torch.arange(lo, hi, device="cuda", dtype=torch.int64)) & (tn - 1)).to(torch.int32)
    dst = torch.empty(n, device="cuda", dtype=torch.int32)
    dev.gather(4, n, wrap(dev, table), wrap(dev, idx), wrap(dev, dst))
    torch.cuda.synchronize()
    for lo in range(0, n, 1 << 28):
        hi = min(n, lo + (1 << 28))
        want = ((n - 1 - torch.arange(lo, hi, device="cuda", dtype=torch.int64)) & (tn - 1)) * 7 + 3
        assert bool((dst[lo:hi].to(torch.int64) == want).all()), f"mismatch in [{lo}, {hi})"


def test_2p32_elements_are_rejected_not_truncated(dev):
    one = dev.create_buffer(16)
    with pytest.raises(hj.HjError):
        dev.execute_graph([{"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": irm.c2_chain_ir(), "size": 1 << 32}],
                          [one, one], [(1 << 32, hj.F32, 4), (1 << 32, hj.F32, 4)])
    with pytest.raises(hj.HjError):
        dev.compress(1 << 32, one, one, one)
