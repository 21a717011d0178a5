"""``record`` — host-side mirror of hephaestus-jit/src/record.rs: functions whose trace is compiled
once per input signature and re-launched with new inputs afterwards.

``record(f)`` returns ``g(device, *inputs) -> (output, report)`` like the reference's
``record(f)(&device, inputs)`` (record.rs:93-109); ``@recorded`` is the ``#[recorded]`` attribute
(hephaestus-macros/src/attributes.rs:7-69).  Inputs and outputs may be ``VarRef``s or nested
lists / tuples / dicts of them — the role of the ``Traverse`` / ``Construct`` traits
(traverse.rs:60-80).  The graph cache itself lives in the library (hj_fcache_*, the reference's
``FCache``, record.rs:116-119), keyed by hash(function identity, per-input ``VarRef`` hash).
"""
from __future__ import annotations

import ctypes
import hashlib
import itertools
import os
import pickle

from . import tr
from ._lib import check, lib

_fn_ids = itertools.count(1)
_MASK = (1 << 64) - 1


def _traverse(obj, out: list):
    """Flatten ``obj`` into ``out`` and return its layout (traverse.rs:9-58)."""
    if isinstance(obj, tr.VarRef):
        out.append(obj)
        return "v"
    if isinstance(obj, (list, tuple)):
        return (type(obj).__name__, [_traverse(o, out) for o in obj])
    if isinstance(obj, dict):
        return ("dict", [(k, _traverse(v, out)) for k, v in obj.items()])
    if obj is None:
        return "none"
    if hasattr(obj, "traverse") and hasattr(obj, "construct"):
        return ("obj", type(obj), obj.traverse(out))
    raise TypeError(f"cannot traverse {type(obj)} (expected VarRef or nested list / tuple / dict)")


def _construct(layout, it):
    if layout == "v":
        return next(it)
    if layout == "none":
        return None
    kind = layout[0]
    if kind in ("list", "tuple"):
        items = [_construct(l, it) for l in layout[1]]
        return items if kind == "list" else tuple(items)
    if kind == "dict":
        return {k: _construct(l, it) for k, l in layout[1]}
    if kind == "obj":
        return layout[1].construct(it, layout[2])
    raise TypeError(layout)


def _layout_hash(layout) -> int:
    return hash(repr(layout)) & _MASK


def _mix(h: int, x: int) -> int:
    h ^= x & _MASK
    return (h * 0x100000001B3) & _MASK


def _stable_key(name: str, in_layout, flat) -> str | None:
    """A key that means the same in another process: function name, input layout, and per input
    (scalar type, size, dynamic?).  None if an input has a composite type (TypeIds of composite
    types depend on the interning order of the process)."""
    h = hashlib.sha256(name.encode())
    h.update(repr(in_layout).encode())
    for v in flat:
        ty = v.ty()
        if ty > 12:  # not a scalar VarType (vartype.rs:89-104)
            return None
        h.update(f"|{ty}:{v.capacity()}:{int(v.is_dynamic())}:{int(v.is_unsized())}".encode())
    return h.hexdigest()[:32]


def record(f, cache_dir: str | None = None, name: str | None = None):
    """``record(f)``.  With ``cache_dir`` the compiled graphs also persist on disk (graph wire format,
    csrc/tgraph_io.cpp, next to the cubin cache of csrc/jit.cpp): a later process that records a
    function under the same ``name`` with the same input signature launches the stored graph without
    tracing, scheduling or compiling.  Buffers the function captured are stored with the contents
    they had when the graph was written.  ``name`` defaults to the function's qualified name plus a
    hash of its bytecode."""
    fn_id = next(_fn_ids)  # TypeId::of::<F>() (record.rs:157)
    layouts = {}
    if name is None:
        code = getattr(f, "__code__", None)
        name = f"{getattr(f, '__module__', '')}.{getattr(f, '__qualname__', 'fn')}:" + (
            hashlib.sha256(code.co_code + repr(code.co_consts).encode()).hexdigest()[:16] if code else "")

    def call(device, *inputs):
        flat = []
        in_layout = _traverse(list(inputs), flat)
        resource_inputs = [v for v in flat if not v.is_unsized()]
        # evaluate the inputs so their dependencies are not collected into the function's graph
        for v in flat:
            v.schedule()
        tr.compile().launch(device)
        key = _mix(0xCBF29CE484222325, fn_id)
        key = _mix(key, _layout_hash(in_layout))
        for v in flat:
            key = _mix(key, v.hash())
        g = ctypes.c_void_p()
        check(lib.hj_fcache_get(key, ctypes.byref(g)))
        graph = tr.Graph(g.value) if g.value else None
        path = None
        if graph is None and cache_dir is not None:
            stable = _stable_key(name, in_layout, flat)
            path = os.path.join(cache_dir, stable + ".hjgraph") if stable else None
            if path and os.path.exists(path) and os.path.exists(path + ".layout"):
                try:
                    with open(path, "rb") as fh:
                        graph = tr.Graph.deserialize(fh.read(), device)
                    with open(path + ".layout", "rb") as fh:
                        layouts[key] = pickle.load(fh)
                    check(lib.hj_fcache_put(key, graph._h))
                except Exception:  # damaged or written by another ABI version: trace again
                    graph = None
        if graph is None:
            output = f(*inputs)
            outs = []
            layouts[key] = _traverse(output, outs)
            graph = tr.compile_fn(resource_inputs, outs)
            check(lib.hj_fcache_put(key, graph._h))
            del outs, output
            if path:
                try:
                    blob, lay = graph.serialize(), pickle.dumps(layouts[key])
                    os.makedirs(cache_dir, exist_ok=True)
                    for target, data in ((path, blob), (path + ".layout", lay)):
                        tmp = f"{target}.{os.getpid()}.tmp"
                        with open(tmp, "wb") as fh:
                            fh.write(data)
                        os.replace(tmp, target)
                except Exception:  # e.g. an output layout that holds a user type: stay in memory only
                    pass
        report, outputs = graph.launch_with(device, resource_inputs)
        return _construct(layouts[key], iter(outputs)), report

    call.__name__ = getattr(f, "__name__", "recorded")
    return call


def recorded(f):
    """``#[recorded] fn f(x: &VarRef) -> VarRef`` => ``f(&device, &x)``."""
    return record(f)
