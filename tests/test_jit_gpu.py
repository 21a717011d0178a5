"""GPU parity tests for the fused-kernel path: IR -> CUDA C++ -> NVRTC -> sm_100a cubin -> launch,
compared with the numpy interpreter of the reference's per-op semantics on the same inputs, and
with the reference's own known answers where its tests state them."""
import importlib

import numpy as np
import pytest

import ir_cases
import oracle
from oracle import ir_interp

pytestmark = pytest.mark.gpu

hj = importlib.import_module("hephaestus-jit_b200")
irm = importlib.import_module("hephaestus-jit_b200.ir")


@pytest.fixture(scope="module")
def dev():
    return hj.Device.cuda(0)


def run_gpu(dev, case, force_scalar=False, monkeypatch=None):
    bufs = [dev.create_buffer_from_slice(b) for b in case.buffers]
    sb = dev.create_buffer_from_slice(case.size_buf) if case.size_buf is not None else None
    k = dev.kernel(case.builder)
    dev.launch(k, case.size, bufs, size_buf=sb, index_base=case.index_base)
    return [b.to_host(h.dtype).reshape(h.shape) for b, h in zip(bufs, case.buffers)]


def compare(case, got, want):
    slots = case.check_slots if case.check_slots is not None else range(len(want))
    for s in slots:
        if s in case.unordered_unique:
            continue
        g, w = got[s], want[s]
        if case.exact or not np.issubdtype(w.dtype, np.floating):
            # bitwise: bitcasts of random integers produce NaN payloads that must survive too
            assert g.tobytes() == w.tobytes(), (case.name, s, g.ravel()[:8], w.ravel()[:8])
        else:
            assert np.allclose(g, w, rtol=case.rtol, atol=case.atol, equal_nan=True), \
                (case.name, s, np.max(np.abs(g.astype(np.float64) - w.astype(np.float64))))
    for s in case.unordered_unique:
        vals = got[s][got[s] != 0xFFFFFFFF]
        assert len(set(vals.tolist())) == len(vals), "atomic results must be unique"


@pytest.mark.parametrize("make", ir_cases.ALL_CASES, ids=ir_cases.case_id)
def test_fused_kernel_matches_interpreter(dev, make):
    case = make()
    want = [np.array(b, copy=True) for b in case.buffers]
    ir_interp.run_ir(case.builder, case.size, want, size_buf=case.size_buf, index_base=case.index_base)
    got = run_gpu(dev, case)
    compare(case, got, want)
    for slot, exp in case.expect.items():  # the reference's own stated answer
        if case.exact or np.issubdtype(exp.dtype, np.integer):
            assert np.array_equal(got[slot].reshape(exp.shape), exp)
        else:
            assert np.allclose(got[slot].reshape(exp.shape), exp, rtol=case.rtol, atol=case.atol)


@pytest.mark.parametrize("make", [ir_cases.c2_chain, ir_cases.mixed_width, ir_cases.index_base,
                                  ir_cases.dyn_size, lambda: ir_cases.binary_ops(ir_cases.I.U32)],
                         ids=["c2", "mixed_width", "index_base", "dyn_size", "binary_u32"])
def test_scalar_and_vector_entry_points_agree(dev, make, monkeypatch):
    case = make()
    vec = run_gpu(dev, case)
    monkeypatch.setenv("HJ_JIT_SCALAR", "1")
    scalar = run_gpu(dev, case)
    for a, b in zip(vec, scalar):
        assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("n", [1, 7, 2047, 2048, 2049, 1 << 20, (1 << 22) + 5])
def test_c2_chain_sizes_vs_oracle(dev, n):
    """BASELINE C2 against the C oracle; tolerance: 2 ulp (CUDA sinf/exp2f) — the reference's own
    bound is the Vulkan precision table (sin: abs 2^-11)."""
    case = ir_cases.c2_chain(n=n, seed=n)
    got = run_gpu(dev, case)[1]
    want = oracle.c2_chain(case.buffers[0])
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
    assert np.all(err <= 2 * ulp + 1e-9)
    assert np.max(err) < 2.0 ** -11


def test_kernel_cache_hits(dev):
    before = dev.kernel_cache_stats()
    b1 = ir_cases.in_place_update().builder
    dev.kernel(b1)
    dev.kernel(ir_cases.in_place_update().builder)
    after = dev.kernel_cache_stats()
    assert after["hits"] >= before["hits"] + 1
    assert after["compiled"] + after["disk_hits"] <= before["compiled"] + before["disk_hits"] + 1


def test_unaligned_buffers_use_scalar_entry(dev):
    case = ir_cases.c2_chain(n=5000)
    x = case.buffers[0]
    whole = dev.create_buffer_from_slice(np.concatenate([np.zeros(1, np.float32), x]))
    view = dev.wrap(whole.ptr + 4, x.nbytes)
    out = dev.create_buffer(x.nbytes)
    dev.launch(dev.kernel(case.builder), x.size, [view, out])
    ref = run_gpu(dev, case)[1]
    assert np.array_equal(out.to_host(np.float32), ref)


# ---- execute_graph ----------------------------------------------------------------------------

def test_execute_graph_fill_compress_gather_reduce(dev):
    """The pass list `mask.compress()` + gather + reduce_sum produces (trace.rs:1595-1620):
    zero-fill kernel, Compress device op, DynSize gather kernel, Reduce device op."""
    n = 100003
    rng = np.random.Generator(np.random.PCG64(0))
    vals = rng.integers(0, 10, size=n).astype(np.int32)
    I = ir_cases.I
    # pass 0: mask = vals < 7 ; index = 0 ; count = 0  (three kernels in the reference; fused here
    # by extent as graph.rs:463-495 would: mask+index share size n, count has size 1)
    b0 = irm.IRBuilder()
    i32, boolt, u32 = b0.scalar(I.I32), b0.scalar(I.BOOL), b0.scalar(I.U32)
    vref = b0.buffer_ref(i32)
    idx = b0.index()
    v = b0.gather(i32, vref, idx)
    m = b0.bop(I.BOP_LT, boolt, v, b0.literal(I.I32, 7))
    mref = b0.buffer_ref(boolt)
    b0.scatter(mref, m, idx)
    iref = b0.buffer_ref(u32)
    b0.scatter(iref, b0.literal(I.U32, 0), idx)
    b1 = irm.IRBuilder()
    cref = b1.buffer_ref(b1.scalar(I.U32))
    b1.scatter(cref, b1.literal(I.U32, 0), b1.index())
    # pass 3: out[i] = vals[index[i]] for i < count (DynSize)
    b3 = irm.IRBuilder()
    i32b, u32b = b3.scalar(I.I32), b3.scalar(I.U32)
    ir_ = b3.buffer_ref(u32b)
    idx3 = b3.index()
    k = b3.gather(u32b, ir_, idx3)
    vr = b3.buffer_ref(i32b)
    g = b3.gather(i32b, vr, k)
    orf = b3.buffer_ref(i32b)
    b3.scatter(orf, g, idx3)
    # resources: 0 vals, 1 mask, 2 index, 3 count, 4 gathered, 5 sum
    env = [dev.create_buffer_from_slice(vals), dev.create_buffer(n), dev.create_buffer(4 * n),
           dev.create_buffer(4), dev.create_buffer_from_slice(np.zeros(n, np.int32)), dev.create_buffer(4)]
    descs = [(n, hj.I32, 4), (n, hj.BOOL, 1), (n, hj.U32, 4), (1, hj.U32, 4), (n, hj.I32, 4), (1, hj.I32, 4)]
    passes = [
        {"kind": hj.PASS_KERNEL, "resources": [0, 1, 2], "ir": b0, "size": n},
        {"kind": hj.PASS_KERNEL, "resources": [3], "ir": b1, "size": 1},
        {"kind": hj.PASS_COMPRESS, "resources": [2, 3, 1]},
        {"kind": hj.PASS_KERNEL, "resources": [2, 0, 4], "ir": b3, "size": n, "size_buffer": 3},
        {"kind": hj.PASS_REDUCE, "arg": hj.SUM, "resources": [5, 4]},
    ]
    report = dev.execute_graph(passes, env, descs, timed=True)
    assert [r[0].split(" [")[0] for r in report] == ["JIT Kernel 0", "JIT Kernel 1", "Compress Large",
                                                     "JIT Kernel 3", "Reduce"]
    assert all(r[2] > 0 for r in report)
    sel = vals[vals < 7]
    assert int(env[3].to_host(np.uint32)[0]) == sel.size
    assert np.array_equal(env[4].to_host(np.int32)[: sel.size], sel)
    assert int(env[5].to_host(np.int32)[0]) == int(sel.sum())
    # untimed (asynchronous) execution gives the same answer
    env[5].fill_zero()
    dev.execute_graph(passes, env, descs)
    assert int(env[5].to_host(np.int32)[0]) == int(sel.sum())


def test_execute_graph_prefix_sum_pass_and_errors(dev):
    n = 50000
    x = np.arange(n, dtype=np.uint64)
    env = [dev.create_buffer(8 * n), dev.create_buffer_from_slice(x)]
    descs = [(n, hj.U64, 8), (n, hj.U64, 8)]
    dev.execute_graph([{"kind": hj.PASS_PREFIX_SUM, "arg": 1, "resources": [0, 1]}], env, descs)
    assert np.array_equal(env[0].to_host(np.uint64), np.cumsum(x, dtype=np.uint64))
    with pytest.raises(hj.HjError):  # MatMul etc. are out of scope
        dev.execute_graph([{"kind": 7, "resources": [0, 1]}], env, descs)
    with pytest.raises(hj.HjError):  # empty resource slot
        dev.execute_graph([{"kind": hj.PASS_PREFIX_SUM, "arg": 1, "resources": [0, 1]}], [env[0], None], descs)


def test_map_host_streams_chunks_with_global_index(dev):
    """hj_kernel_map_host: host arrays in, host arrays out, chunked through three streams; Index
    keeps its global value across chunks and the ragged last chunk is masked."""
    import ctypes
    irm = importlib.import_module("hephaestus-jit_b200.ir")
    n = (1 << 20) + 12345
    rng = np.random.Generator(np.random.PCG64(3))
    x = (rng.random(n, dtype=np.float32) * 8 - 4).astype(np.float32)
    y = np.zeros(n, np.float32)
    k = dev.kernel(irm.c2_chain_ir())
    dev.map_host(k, n, [x, y], chunk_elems=1 << 17)  # 9 chunks, pageable host memory is allowed
    assert np.allclose(y, oracle.c2_chain(x), rtol=4e-7, atol=1e-7)
    # y[i] = x[i] + f32(Index): the index must be global, not chunk-local
    b = irm.IRBuilder()
    f32, u32 = b.scalar(irm.F32), b.scalar(irm.U32)
    src = b.buffer_ref(f32, 0)
    idx = b.index()
    val = b.gather(f32, src, idx)
    s = b.bop(irm.BOP_ADD, f32, val, b.uop(irm.UOP_CAST, f32, idx))
    b.scatter(b.buffer_ref(f32, 1), s, idx)
    k2 = dev.kernel(b)
    dev.map_host(k2, n, [x, y], chunk_elems=1 << 18)
    assert np.array_equal(y, x + np.arange(n, dtype=np.uint32).astype(np.float32))
