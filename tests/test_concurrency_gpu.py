"""Concurrent execute_graph calls from several threads (Device is Send + Sync; the
reference serialises submission behind a fence mutex, vulkan_core/device.rs:301-332).  Each thread
launches pass lists with its own buffers; the kernels are DISTINCT IRs, so every thread misses the
kernel cache and compiles — outside the cache lock (jit.cpp: in-flight set + condition variable) —
while the others keep hitting it; a second round has all threads ask for the SAME new IR at once."""
import importlib
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
hj = importlib.import_module("hephaestus-jit_b200")
irm = importlib.import_module("hephaestus-jit_b200.ir")


def _affine_ir(mul: int, add: int):
    b = irm.IRBuilder()
    u32 = b.scalar(hj.U32)
    src, idx = b.buffer_ref(u32), b.index()
    v = b.bop(irm.BOP_ADD, u32, b.bop(irm.BOP_MUL, u32, b.gather(u32, src, idx), b.literal(hj.U32, mul)), b.literal(hj.U32, add))
    b.scatter(b.buffer_ref(u32), v, idx)
    return b


def test_threads_compile_distinct_kernels_and_launch_concurrently():
    dev = hj.Device.cuda(0)
    n, n_threads, rounds = (1 << 18) + 3, 4, 6
    stats0 = dev.kernel_cache_stats()
    errors, barrier = [], threading.Barrier(n_threads)
    salt = int(time.time()) & 0xFFFF   # fresh constants: nothing comes out of the on-disk cubin cache

    def worker(t):
        try:
            rng = np.random.Generator(np.random.PCG64(t))
            x = rng.integers(0, 1 << 20, size=n).astype(np.uint32)
            bx, by, bs, bscan = dev.create_buffer_from_slice(x), dev.create_buffer(4 * n), dev.create_buffer(4), dev.create_buffer(4 * n)
            descs = [(n, hj.U32, 4), (n, hj.U32, 4), (1, hj.U32, 4), (n, hj.U32, 4)]
            barrier.wait()
            for r in range(rounds):
                mul, add = 3 + 2 * t + 100 * r + salt, 7 * t + r     # a new IR every round and thread
                passes = [{"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": _affine_ir(mul, add), "size": n},
                          {"kind": hj.PASS_REDUCE, "arg": hj.SUM, "resources": [2, 1]},
                          {"kind": hj.PASS_PREFIX_SUM, "arg": 1, "resources": [3, 1]}]
                dev.execute_graph(passes, [bx, by, bs, bscan], descs)
                want = x * np.uint32(mul) + np.uint32(add)
                assert np.array_equal(by.to_host(np.uint32), want), (t, r)
                assert bs.to_host(np.uint32)[0] == np.uint32(want.sum(dtype=np.uint64) & 0xFFFFFFFF), (t, r)
                assert np.array_equal(bscan.to_host(np.uint32), np.cumsum(want, dtype=np.uint32)), (t, r)
            # everybody asks for the same, new IR at the same moment: one compiles, the others wait for it
            barrier.wait()
            shared = _affine_ir(999983 + salt, 17)
            dev.execute_graph([{"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": shared, "size": n}], [bx, by], descs[:2])
            assert np.array_equal(by.to_host(np.uint32), x * np.uint32(999983 + salt) + np.uint32(17)), t
        except BaseException as exc:  # noqa: BLE001
            errors.append((t, repr(exc)))
            try:
                barrier.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=600)
    assert not errors, errors
    stats = dev.kernel_cache_stats()
    compiled = stats["compiled"] + stats["disk_hits"] - stats0["compiled"] - stats0["disk_hits"]
    assert compiled == n_threads * rounds + 1, (stats0, stats)   # the shared IR was compiled exactly once
