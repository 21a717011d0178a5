"""`tr.array_async -> graph.launch -> to_vec`: the reference's call sequence (trace.rs:647-663, graph.rs:315-323,
trace.rs:1404-1438) with the three steps overlapped chunk by chunk (hj_buffer_create_from_host_async,
jit.cpp: kernel_launch_streamed, chunk-wise hj_buffer_to_host).

Every test runs the same traced program twice — inputs uploaded with the blocking `tr.array` and with
`tr.array_async` — and demands bit-identical results, plus the numpy statement of the program where it is exact.
The chunk size is forced down to 2^16 elements so that a million elements already give a ramped schedule of
~20 chunks; the launch counter proves which launches went chunk-wise."""
import gc
import importlib
import os

os.environ.setdefault("HJ_ASYNC_CHUNK_ELEMS", "65536")  # read once, at the first asynchronous upload of the process

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

hj = importlib.import_module("hephaestus-jit_b200")
tr = importlib.import_module("hephaestus-jit_b200.tr")

U32, F32, U8, BOOL = hj.U32, hj.F32, hj.U8, hj.BOOL
N = 1_000_003  # ragged on purpose


@pytest.fixture(scope="module")
def device():
    return hj.Device.cuda(0)


@pytest.fixture(autouse=True)
def clean_trace():
    hj.lib.hj_tr_reset_schedule()
    gc.collect()
    base = tr.n_live()
    yield
    hj.lib.hj_tr_reset_schedule()
    gc.collect()
    assert tr.n_live() == base


def pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy()


def chain(x):  # BASELINE C2: fma -> sin / exp2 -> select
    y = x.fma(tr.literal(1.0009765625, F32), tr.literal(0.5, F32))
    return y.sin().select(y.gt(tr.literal(0.75, F32)), y.exp2())


def run(make_inputs, program, device):
    """Traces `program` over `make_inputs()`, launches, returns (outputs as numpy, kernel launches of the launch)."""
    ins = make_inputs()
    outs = program(*ins)
    outs = outs if isinstance(outs, (list, tuple)) else [outs]
    for o in outs:
        o.schedule()
    g = tr.compile()
    before = device.launch_count()
    g.launch(device)
    launches = device.launch_count() - before
    res = [o.to_vec() for o in outs]
    return res, launches


def test_chain_matches_blocking_upload_and_is_streamed(device):
    rng = np.random.default_rng(1)
    x = pinned(rng.random(N, dtype=np.float32))
    ref, n_ref = run(lambda: [tr.array(x, device)], chain, device)
    got, n_got = run(lambda: [tr.array_async(x, device)], chain, device)
    assert n_ref == 1
    assert n_got > 8, "the launch did not go chunk by chunk"
    assert np.array_equal(ref[0].view(np.uint32), got[0].view(np.uint32))


def test_pageable_source_and_single_chunk(device):
    x = np.arange(1000, dtype=np.float32)  # pageable, one chunk
    got, n = run(lambda: [tr.array_async(x, device)], lambda a: a.mul(tr.literal(2.0, F32)), device)
    assert n == 1 and np.array_equal(got[0], x * 2)


def test_two_async_inputs_of_different_width_and_two_outputs(device):
    rng = np.random.default_rng(2)
    a = pinned(rng.integers(0, 1 << 30, N, dtype=np.uint32))
    m = pinned((rng.random(N) < 0.5).astype(np.uint8))

    def prog(va, vm):
        keep = vm.neq(tr.literal(0, U8))
        return [va.select(keep, tr.literal(7, U32)), va.add(tr.index())]

    ref, _ = run(lambda: [tr.array(a, device), tr.array(m, device)], prog, device)
    got, n = run(lambda: [tr.array_async(a, device), tr.array_async(m, device)], prog, device)
    assert n > 8
    assert np.array_equal(got[0], np.where(m != 0, a, 7).astype(np.uint32))
    assert np.array_equal(got[1], a + np.arange(N, dtype=np.uint32))
    assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])


def test_async_and_resident_input_mixed(device):
    rng = np.random.default_rng(3)
    a = pinned(rng.integers(0, 1000, N, dtype=np.uint32))
    b = rng.integers(0, 1000, N, dtype=np.uint32)
    got, n = run(lambda: [tr.array_async(a, device), tr.array(b, device)], lambda x, y: x.mul(y), device)
    assert n > 8 and np.array_equal(got[0], a * b)


def test_to_vec_ranges_of_a_streamed_result_and_of_the_input(device):
    a = pinned(np.arange(N, dtype=np.uint32))
    x = tr.array_async(a, device)
    assert np.array_equal(x.to_vec(start=70_000, end=200_001), a[70_000:200_001])  # the upload itself, chunk-wise
    y = x.add(tr.literal(5, U32))
    y.schedule()
    tr.compile().launch(device)
    assert np.array_equal(y.to_vec(start=65_535, end=65_537), a[65_535:65_537] + 5)
    out = pinned(np.empty(N, dtype=np.uint32))
    y.to_vec(out=out.view(np.uint8))
    assert np.array_equal(out, a + 5)
    assert np.array_equal(y.to_vec(), a + 5)


def test_consumers_that_cannot_stream_wait_for_the_upload(device):
    rng = np.random.default_rng(4)
    a = pinned(rng.integers(0, 100, N, dtype=np.uint32))
    # device op straight on the arriving buffer
    x = tr.array_async(a, device)
    s = x.reduce_sum()
    s.schedule()
    tr.compile().launch(device)
    assert int(s.item()) == int(a.sum(dtype=np.uint64) & 0xffffffff)
    # gather through a computed index (not the bare Index)
    x = tr.array_async(a, device)
    idx = tr.sized_index(N)
    rev = tr.literal(N - 1, U32).sub(idx)
    y = x.gather(rev)
    y.schedule()
    before = device.launch_count()
    tr.compile().launch(device)
    assert device.launch_count() - before == 1
    assert np.array_equal(y.to_vec(), a[::-1])
    # kernel pass + scan in one graph
    x = tr.array_async(a, device)
    p = x.add(tr.literal(1, U32)).prefix_sum(True)
    p.schedule()
    tr.compile().launch(device)
    assert np.array_equal(p.to_vec(), np.cumsum(a + 1, dtype=np.uint32))


@pytest.mark.parametrize("inclusive", [True, False])
def test_scan_of_an_arriving_array_is_streamed_and_bit_identical(device, inclusive):
    rng = np.random.default_rng(7)
    a = pinned(rng.integers(0, 1 << 32, N, dtype=np.uint32))  # wraps many times
    ref, n_ref = run(lambda: [tr.array(a, device)], lambda x: x.prefix_sum(inclusive), device)
    got, n_got = run(lambda: [tr.array_async(a, device)], lambda x: x.prefix_sum(inclusive), device)
    assert n_got > 2 * n_ref + 8, "the scan did not go chunk by chunk"
    want = np.cumsum(a, dtype=np.uint32)
    if not inclusive:
        want = np.concatenate([[0], want[:-1]]).astype(np.uint32)
    assert np.array_equal(got[0], want) and np.array_equal(ref[0], got[0])


def test_float_scan_of_an_arriving_array_is_not_cut_differently(device):
    rng = np.random.default_rng(8)
    a = pinned(rng.random(N, dtype=np.float32))
    ref, n_ref = run(lambda: [tr.array(a, device)], lambda x: x.prefix_sum(True), device)
    got, n_got = run(lambda: [tr.array_async(a, device)], lambda x: x.prefix_sum(True), device)
    assert n_got == n_ref  # waits for the whole upload: same kernel, same rounding
    assert np.array_equal(ref[0].view(np.uint32), got[0].view(np.uint32))


def test_streamed_result_feeds_later_launches_and_inputs_are_reusable(device):
    rng = np.random.default_rng(5)
    a = pinned(rng.integers(0, 1 << 20, N, dtype=np.uint32))
    x = tr.array_async(a, device)
    y = x.mul(tr.literal(3, U32))
    y.schedule()
    tr.compile().launch(device)          # streamed
    z = y.add(x)                          # both operands carry progress of the streamed launch: settle, then launch
    z.schedule()
    tr.compile().launch(device)
    w = z.reduce_max()
    w.schedule()
    tr.compile().launch(device)
    assert np.array_equal(z.to_vec(), a * 4)
    assert int(w.item()) == int((a * 4).max())
    assert np.array_equal(y.to_vec(), a * 3)
    assert np.array_equal(x.to_vec(), a)


def test_dropping_an_arriving_array_and_many_in_flight(device):
    a = pinned(np.arange(N, dtype=np.uint32))
    for _ in range(8):
        tr.array_async(a, device)  # released while the upload may still run: the free is ordered behind it
    xs = [tr.array_async(a, device) for _ in range(4)]
    ys = [x.add(tr.literal(i, U32)) for i, x in enumerate(xs)]
    for y in ys:
        y.schedule()
    tr.compile().launch(device)  # one kernel, four arriving inputs with the same schedule
    for i, y in enumerate(ys):
        assert np.array_equal(y.to_vec(), a + i)
    device.sync()


@pytest.mark.parametrize("seed", range(24))
def test_random_programs_do_not_depend_on_how_the_inputs_arrive(device, seed):
    """A random traced program — elementwise chains over up to three inputs of mixed width, comparisons and
    selects, the global Index, optionally ended by a scan, a reduction or a gather through a computed index,
    one to three scheduled results — run with blocking uploads and with any subset of the inputs arriving
    asynchronously: every result bit-identical.  Streamed or not is the library's choice; the values are not."""
    import random
    rnd = random.Random(9000 + seed)
    rng = np.random.default_rng(9000 + seed)
    n = rnd.choice([N, 300_001, 65_536 * 3, 70_000])
    host = [pinned(rng.integers(0, 1 << 16, n, dtype=np.uint32)), pinned(rng.random(n, dtype=np.float32)),
            pinned((rng.random(n) < 0.4).astype(np.uint8))]
    arriving = [rnd.random() < 0.7 for _ in host]
    if not any(arriving):
        arriving[rnd.randrange(3)] = True
    script = [(rnd.randrange(8), rnd.randrange(100), rnd.randrange(100), rnd.randrange(1, 9)) for _ in range(rnd.randrange(2, 9))]
    ending = rnd.randrange(5)
    n_out = rnd.randrange(1, 4)

    def program(asynchronous):
        make = lambda a, late: (tr.array_async if (asynchronous and late) else tr.array)(a, device)
        vu, vf, vm = (make(a, late) for a, late in zip(host, arriving))
        ints, floats = [vu, tr.sized_index(n)], [vf]
        keep = vm.neq(tr.literal(0, U8))
        for op, i, j, c in script:
            a, b = ints[i % len(ints)], ints[j % len(ints)]
            x, y = floats[i % len(floats)], floats[j % len(floats)]
            if op == 0:
                ints.append(a.add(b))
            elif op == 1:
                ints.append(a.mul(tr.literal(c, U32)).xor(b))
            elif op == 2:
                ints.append(a.select(keep, b))
            elif op == 3:
                ints.append(a.min(b).or_(tr.literal(c, U32)))
            elif op == 4:
                floats.append(x.fma(tr.literal(1.25, F32), y))
            elif op == 5:
                floats.append(x.sin().select(x.gt(y), y.exp2()))
            elif op == 6:
                ints.append(x.mul(tr.literal(1000.0, F32)).cast(U32).add(a))
            else:
                floats.append(a.and_(tr.literal(1023, U32)).cast(F32).add(x))
        outs = [ints[-1], floats[-1], ints[len(ints) // 2]][:n_out]
        if ending == 0:
            outs.append(ints[-1].prefix_sum(True))
        elif ending == 1:
            outs.append(ints[-1].reduce_sum())
        elif ending == 2:
            outs.append(ints[-1].gather(tr.literal(n - 1, U32).sub(tr.sized_index(n))))
        elif ending == 3:
            outs.append(vu.prefix_sum(False))  # a scan straight on an (arriving) input
        for o in outs:
            o.schedule()
        tr.compile().launch(device)
        return [o.to_vec() for o in outs]

    want = program(False)
    got = program(True)
    assert len(want) == len(got)
    for w, g in zip(want, got):
        assert w.dtype == g.dtype and np.array_equal(w.view(np.uint8), g.view(np.uint8))


def test_two_host_threads_stream_their_own_arrays(device):
    """Device: Send + Sync.  The side streams are per device: two threads' chunks interleave on them,
    ordered by their own events only.  Through the buffer-level API (the trace and its schedule are per thread)."""
    import ctypes
    import threading
    irm = importlib.import_module("hephaestus-jit_b200.ir")

    def ir_add(c):  # dst[i] = src[i] + c
        b = irm.IRBuilder()
        t = b.scalar(U32)
        src = b.buffer_ref(t)
        idx = b.index()
        v = b.bop(irm.BOP_ADD, t, b.gather(t, src, idx), b.literal(U32, c))
        dst = b.buffer_ref(t)
        b.scatter(dst, v, idx)
        return b

    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for it in range(3):
                a = pinned(rng.integers(0, 1 << 30, N, dtype=np.uint32))
                out = ctypes.c_void_p()
                hj.check(hj.lib.hj_buffer_create_from_host_async(device.handle, a.ctypes.data_as(ctypes.c_void_p), a.nbytes, 4,
                                                                 ctypes.byref(out)))
                bx = hj.Buffer(out.value, device)
                by = device.create_buffer(4 * N)
                device.execute_graph([{"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": ir_add(seed + it), "size": N}],
                                     [bx, by], [(N, U32, 4), (N, U32, 4)])
                got = by.to_host(np.uint32)
                if not np.array_equal(got, a + np.uint32(seed + it)):
                    errors.append(f"thread {seed} iteration {it}: wrong result")
        except BaseException as exc:  # noqa: BLE001
            errors.append(repr(exc))

    ts = [threading.Thread(target=worker, args=(s,)) for s in (11, 23)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_buffer_level_api_through_execute_graph(device):
    """The same without the trace layer: hj_buffer_create_from_host_async + one Kernel pass + to_host."""
    irm = importlib.import_module("hephaestus-jit_b200.ir")
    rng = np.random.default_rng(6)
    x = pinned(rng.random(N, dtype=np.float32))
    import ctypes
    out = ctypes.c_void_p()
    hj.check(hj.lib.hj_buffer_create_from_host_async(device.handle, x.ctypes.data_as(ctypes.c_void_p), x.nbytes, 4,
                                                     ctypes.byref(out)))
    bx = hj.Buffer(out.value, device)
    by = device.create_buffer(4 * N)
    k = device.kernel(irm.c2_chain_ir())
    before = device.launch_count()
    device.execute_graph([{"kind": hj.PASS_KERNEL, "resources": [0, 1], "ir": irm.c2_chain_ir(), "size": N}], [bx, by],
                         [(N, F32, 4), (N, F32, 4)])
    assert device.launch_count() - before > 8
    got = by.to_host(np.float32)
    bx2 = device.create_buffer_from_slice(x)
    by2 = device.create_buffer(4 * N)
    device.launch(k, N, [bx2, by2])
    assert np.array_equal(got.view(np.uint32), by2.to_host(np.float32).view(np.uint32))
