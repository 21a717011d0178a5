"""Host logic of the asynchronous arrays that needs no GPU: the chunk schedule (hj_async_chunk_schedule) every
upload, streamed kernel pass and chunk-wise download shares.  A wrong schedule — a gap, an overlap, a misaligned
boundary — would corrupt data silently, so the invariants are checked over many sizes."""
import ctypes
import importlib

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

L = importlib.import_module("hephaestus-jit_b200._lib")


def schedule(n, chunk):
    cap = 1 << 17
    first = (ctypes.c_uint64 * cap)()
    count = (ctypes.c_uint64 * cap)()
    k = ctypes.c_uint32()
    L.check(L.lib.hj_async_chunk_schedule(n, chunk, first, count, cap, ctypes.byref(k)))
    assert k.value <= cap
    return np.array(first[: k.value], dtype=np.uint64), np.array(count[: k.value], dtype=np.uint64)


def check_invariants(n, chunk):
    first, count = schedule(n, chunk)
    if n == 0:
        assert len(first) == 0
        return first, count
    assert first[0] == 0 and int(first[-1] + count[-1]) == n          # covers [0, n)
    assert np.array_equal(first[1:], (first + count)[:-1])           # contiguous, no gap, no overlap
    assert (count > 0).all()
    assert (first % 4096 == 0).all()                                  # 16-byte aligned for every element size
    full = max(4096, -(-chunk // 4096) * 4096)
    assert int(count.max()) < full + 4096                            # the ragged end joins the last chunk
    return first, count


@pytest.mark.parametrize("n", [0, 1, 4095, 4096, 4097, 65536, 1_000_003, (1 << 28), (1 << 32) - 1])
@pytest.mark.parametrize("chunk", [1, 4096, 65536, 1 << 24])
def test_schedule_invariants(n, chunk):
    if n >= (1 << 28) and chunk < 65536:
        pytest.skip("more chunks than the test's capacity")
    first, count = check_invariants(n, chunk)
    if n % 4096 == 0 and n:
        assert (count % 4096 == 0).all()


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 1 << 26), st.integers(1, 1 << 22))
def test_schedule_invariants_random(n, chunk):
    check_invariants(n, chunk)


def test_long_arrays_ramp_up_and_down():
    first, count = check_invariants(1 << 28, 1 << 24)
    c = count.tolist()
    assert c[:3] == [1 << 21, 1 << 22, 1 << 23] and c[-3:] == [1 << 23, 1 << 22, 1 << 21]
    assert set(c[3:-4]) == {1 << 24} and 0 < c[-4] <= 1 << 24  # full chunks in the middle (the last one may be short)
    ragged, rc = check_invariants((1 << 28) + 5, 1 << 24)
    assert rc.tolist()[-1] == (1 << 21) + 5 and rc.tolist()[:-1] == c[:-1]
    short, _ = check_invariants(3 << 24, 1 << 24)  # fewer than four full chunks: no ramp
    assert len(short) == 3


def test_null_arguments_are_rejected():
    k = ctypes.c_uint32()
    assert L.lib.hj_async_chunk_schedule(10, 0, None, None, 4, ctypes.byref(k)) != 0
    assert L.lib.hj_async_chunk_schedule(10, 0, None, None, 0, ctypes.byref(k)) == 0 and k.value == 1
