"""CPU execution of a whole compiled graph — TEST INFRASTRUCTURE (like everything under oracle/): only
tests/ may import it; the product never does.

Restates the pass interpreter of the reference's backend (`VulkanDevice::execute_graph`,
hephaestus-jit/src/backend/vulkan/mod.rs:151-383) on numpy arrays: kernel passes run through the IR
interpreter (oracle/ir_interp.py), device-op passes through the C restatement of the reference's
builtins (oracle/hj_oracle.c) with the reference's resource order (Reduce / PrefixSum: [dst, src];
Compress: [index_out, out_count, mask]).  The graph comes in through its wire format
(csrc/tgraph_io.cpp), so this also checks that a serialised graph means what the traced program means.

Used to check the trace -> schedule -> graph layers end to end WITHOUT a GPU: a traced program is
compiled, executed here, and compared with the plain numpy statement of the same program.
"""
from __future__ import annotations

import struct
from types import SimpleNamespace

import numpy as np

from . import ir_interp
from . import compress as _compress, prefix_sum as _prefix_sum, reduce as _reduce

MAGIC = b"HJGRAPH1"
INPUT, CAPTURED, INTERNAL = 0, 1, 2
DOP_REDUCE, DOP_PREFIX_SUM, DOP_COMPRESS = 0, 1, 2
POISON = 0xAB  # what an internal buffer holds before anything writes it (the pool hands out stale memory)


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.p = data, 0

    def u32(self) -> int:
        v = struct.unpack_from("<I", self.d, self.p)[0]
        self.p += 4
        return v

    def u64(self) -> int:
        v = struct.unpack_from("<Q", self.d, self.p)[0]
        self.p += 8
        return v

    def u32s(self) -> list:
        n = self.u32()
        v = list(struct.unpack_from(f"<{n}I", self.d, self.p))
        self.p += 4 * n
        return v

    def raw(self, n: int) -> bytes:
        v = self.d[self.p:self.p + n]
        self.p += n
        return v


def parse(blob: bytes) -> SimpleNamespace:
    """The wire format of csrc/tgraph_io.cpp as plain Python data."""
    assert blob[:8] == MAGIC, "not a serialised graph"
    r = _Reader(blob)
    r.p = 8
    r.u32()  # abi version
    types = []
    for _ in range(r.u32()):
        kind, elem, num, cols, rows = (r.u32() for _ in range(5))
        types.append((kind, elem, num, cols, rows, r.u32s()))
    resources = []
    for _ in range(r.u32()):
        kind, size, ty = r.u32(), r.u64(), r.u32()
        data = r.raw(r.u64()) if kind == CAPTURED else None
        resources.append(SimpleNamespace(kind=kind, size=size, ty=ty, data=data))
    inputs, outputs = r.u32s(), r.u32s()
    passes = []
    for _ in range(r.u32()):
        res = r.u32s()
        size_buffer = struct.unpack("<i", struct.pack("<I", r.u32()))[0]
        is_kernel, size, code, arg = r.u32(), r.u64(), r.u32(), r.u32()
        ir = None
        if is_kernel:
            n = r.u32()
            ir_vars = []
            for _ in range(n):
                ty, op, a, ds, de, _pad, data = struct.unpack_from("<6IQ", r.d, r.p)
                r.p += 32
                ir_vars.append((ty, op, a, ds, de, data))
            deps = r.u32s()
            ir_types = []
            for _ in range(r.u32()):
                ir_types.append(struct.unpack_from("<6I", r.d, r.p))
                r.p += 24
            fields = r.u32s()
            ir = SimpleNamespace(vars=ir_vars, deps=deps, types=ir_types, struct_fields=fields, n_buffers=r.u32())
        passes.append(SimpleNamespace(resources=res, size_buffer=size_buffer, is_kernel=bool(is_kernel), size=size,
                                      code=code, arg=arg, ir=ir))
    assert r.p == len(blob) - 8, "trailing bytes"
    return SimpleNamespace(types=types, resources=resources, inputs=inputs, outputs=outputs, passes=passes)


def _dtype(g, ty: int):
    kind = g.types[ty][0]
    assert kind <= ir_interp.F64, "graph_exec handles scalar-typed resources only"
    return ir_interp.NP[kind], kind


def execute(blob: bytes, inputs=()) -> list:
    """Run the graph; returns the arrays of its `outputs` resources (graph.rs:332-393)."""
    g = parse(blob)
    res = [None] * len(g.resources)
    for slot, arr in zip(g.inputs, inputs):
        dt, _ = _dtype(g, g.resources[slot].ty)
        assert arr.size == g.resources[slot].size, "Resource does not match variable type!"
        res[slot] = np.array(arr, dtype=np.uint8 if dt == np.bool_ else dt, copy=True)
    for i, rs in enumerate(g.resources):
        if res[i] is not None:
            continue
        dt, _ = _dtype(g, rs.ty)
        store = np.uint8 if dt == np.bool_ else dt  # bool buffers are bytes (glsl/mod.rs:256-270)
        if rs.kind == CAPTURED:
            res[i] = np.frombuffer(rs.data, dtype=store).copy()
        else:
            res[i] = np.frombuffer(bytes([POISON]) * (rs.size * np.dtype(store).itemsize), dtype=store).copy()
    for p in g.passes:
        bufs = [res[r] for r in p.resources]
        size_buf = res[p.size_buffer] if p.size_buffer >= 0 else None
        if p.is_kernel:
            ir_interp.run_ir(p.ir, p.size, bufs, size_buf=size_buf)
        elif p.code == DOP_REDUCE:
            _, kind = _dtype(g, g.resources[p.resources[0]].ty)
            bufs[0][:1] = _reduce(p.arg, kind, bufs[1])
        elif p.code == DOP_PREFIX_SUM:
            _, kind = _dtype(g, g.resources[p.resources[0]].ty)
            bufs[0][:] = _prefix_sum(kind, bufs[1], bool(p.arg))
        elif p.code == DOP_COMPRESS:
            index_out, out_count, mask = bufs
            n = mask.size if size_buf is None else min(mask.size, int(size_buf[0]))
            count, _ = _compress(mask[:n], index_out)  # entries at and beyond count stay untouched
            out_count[0] = count
        else:
            raise NotImplementedError(f"device op {p.code}")
    return [res[o] for o in g.outputs]
