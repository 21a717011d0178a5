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
import itertools

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


def record(f):
    fn_id = next(_fn_ids)  # TypeId::of::<F>() (record.rs:157)
    layouts = {}

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
        if not g.value:
            output = f(*inputs)
            outs = []
            layouts[key] = _traverse(output, outs)
            graph = tr.compile_fn(resource_inputs, outs)
            check(lib.hj_fcache_put(key, graph._h))
            del outs, output
        else:
            graph = tr.Graph(g.value)
        report, outputs = graph.launch_with(device, resource_inputs)
        return _construct(layouts[key], iter(outputs)), report

    call.__name__ = getattr(f, "__name__", "recorded")
    return call


def recorded(f):
    """``#[recorded] fn f(x: &VarRef) -> VarRef`` => ``f(&device, &x)``."""
    return record(f)
