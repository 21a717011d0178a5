"""GPU tests of hj_execute_graph_cached: a pass list launched again and again (FCache::call,
hephaestus-jit/src/record.rs:120-210) is captured into ONE CUDA graph on its second launch with
the same buffers and replayed afterwards.  Every replay must give what the pass-by-pass path gives
— in particular the scan / compress kernels, whose status-word epoch now lives on the device so that
no launch parameter changes between replays."""
import importlib

import numpy as np
import pytest

import ir_cases
import oracle

pytestmark = pytest.mark.gpu

hj = importlib.import_module("hephaestus-jit_b200")
irm = importlib.import_module("hephaestus-jit_b200.ir")
tr = importlib.import_module("hephaestus-jit_b200.tr")
record = importlib.import_module("hephaestus-jit_b200.record")


@pytest.fixture(scope="module")
def dev():
    return hj.Device.cuda(0)


def make_env(dev, n, rng):
    x = (rng.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
    u = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    mask = (rng.random(n) < 0.3).astype(np.uint8)
    # resources: 0 x, 1 y = chain(x), 2 max(y), 3 u, 4 inclusive scan(u), 5 mask, 6 index, 7 count
    env = [dev.create_buffer_from_slice(x), dev.create_buffer(4 * n), dev.create_buffer(4),
           dev.create_buffer_from_slice(u), dev.create_buffer(4 * n), dev.create_buffer_from_slice(mask),
           dev.create_buffer(4 * n), dev.create_buffer(4)]
    return env, (x, u, mask)


def pass_list(n):
    passes = [
        {"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": irm.c2_chain_ir(), "size": n},
        {"kind": hj.PASS_REDUCE, "arg": hj.MAX, "resources": [2, 1]},
        {"kind": hj.PASS_PREFIX_SUM, "arg": 1, "resources": [4, 3]},
        {"kind": hj.PASS_COMPRESS, "resources": [6, 7, 5]},
    ]
    descs = [(n, hj.F32, 4), (n, hj.F32, 4), (1, hj.F32, 4), (n, hj.U32, 4), (n, hj.U32, 4), (n, hj.BOOL, 1),
             (n, hj.U32, 4), (1, hj.U32, 4)]
    return passes, descs


def check_env(env, host):
    x, u, mask = host
    y = env[1].to_host(np.float32)
    assert np.allclose(y, oracle.c2_chain(x), rtol=4e-7, atol=1e-7)
    assert env[2].to_host(np.float32)[0] == y.max()
    assert np.array_equal(env[4].to_host(np.uint32), oracle.prefix_sum(oracle.U32, u, True))
    cnt, idx = oracle.compress(mask, mt=True)
    assert int(env[7].to_host(np.uint32)[0]) == cnt
    assert np.array_equal(env[6].to_host(np.uint32)[:cnt], idx[:cnt])


@pytest.mark.parametrize("n", [5000, (1 << 21) + 77])  # look-back kernels / ring kernels
def test_capture_then_replay_matches_oracle(dev, n):
    rng = np.random.Generator(np.random.PCG64(n))
    env, host = make_env(dev, n, rng)
    passes, descs = pass_list(n)
    key = 0xC0FFEE00 + n
    c0, r0, p0 = dev.graph_cache_stats()
    hows = []
    for it in range(6):
        # new CONTENTS in the same buffers every launch: a replay must read them, not stale data
        x = (rng.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
        u = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        mask = (rng.random(n) < (0.02 if it % 2 else 0.6)).astype(np.uint8)
        env[0].upload(x); env[3].upload(u); env[5].upload(mask)
        hows.append(dev.execute_graph(passes, env, descs, graph_key=key))
        check_env(env, (x, u, mask))
    assert hows == [0, 1, 2, 2, 2, 2]
    c1, r1, p1 = dev.graph_cache_stats()
    assert (c1 - c0, r1 - r0, p1 - p0) == (1, 4, 1)


def test_other_buffers_take_their_own_instance(dev):
    n = 300001
    rng = np.random.Generator(np.random.PCG64(5))
    passes, descs = pass_list(n)
    key = 0xABCD
    env_a, host_a = make_env(dev, n, rng)
    env_b, host_b = make_env(dev, n, rng)
    seq = []
    for env, host in ((env_a, host_a), (env_b, host_b), (env_a, host_a), (env_b, host_b), (env_a, host_a)):
        seq.append(dev.execute_graph(passes, env, descs, graph_key=key))
        check_env(env, host)
    assert seq == [0, 0, 1, 1, 2]
    # device ops launched outside the graph in between must not disturb the replays (shared scratch, epochs)
    big = rng.integers(0, 4, size=1 << 22).astype(np.uint32)
    out = dev.create_buffer(4 * big.size)
    dev.prefix_sum(hj.U32, big.size, True, dev.create_buffer_from_slice(big), out)
    assert np.array_equal(out.to_host(np.uint32), np.cumsum(big, dtype=np.uint32))
    assert dev.execute_graph(passes, env_b, descs, graph_key=key) == 2
    check_env(env_b, host_b)


def test_recorded_function_relaunch(dev):
    """record(f) relaunches its cached Graph (record.rs:120-210); with CUDA-graph replay the
    results stay those of the first launch's code path."""
    def f(x):
        y = x.mul(tr.literal(2.0, hj.F32)).add(tr.literal(1.0, hj.F32))
        return y, y.reduce_sum(), y.prefix_sum(True)

    rf = record.record(f)
    n = 70001
    rng = np.random.Generator(np.random.PCG64(2))
    before = dev.graph_cache_stats()
    for _ in range(5):
        x = rng.integers(0, 8, size=n).astype(np.float32)
        (y, s, ps), _ = rf(dev, tr.array(x, dev))
        want = x * 2 + 1
        assert np.array_equal(y.to_vec(), want)
        assert float(s.item()) == float(want.sum(dtype=np.float64))  # small integers: exact in f32
        assert np.array_equal(ps.to_vec(), np.cumsum(want.astype(np.float64)).astype(np.float32))
    after = dev.graph_cache_stats()
    assert sum(after) - sum(before) >= 5  # every launch went through hj_execute_graph_cached
