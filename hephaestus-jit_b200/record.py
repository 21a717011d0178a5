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
import json
import os
import types

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
    (scalar type, size, dynamic?).  Unsized inputs are literals whose VALUE is baked into the
    kernel IR as a constant (trace.rs:257-261 hashes it into the in-process key as well), so their
    ``VarRef`` hash — op, type and literal bits, all process-independent for scalars — is part of
    the key: ``f(x, literal(3))`` and ``f(x, literal(5))`` are different graphs.  None if an input
    has a composite type (TypeIds of composite types depend on the interning order of the
    process) or a dynamic size (its size variable is a process-local id)."""
    h = hashlib.sha256(name.encode())
    h.update(repr(in_layout).encode())
    for v in flat:
        ty = v.ty()
        if ty > 12 or v.is_dynamic():  # not a scalar VarType (vartype.rs:89-104)
            return None
        h.update(f"|{ty}:{v.capacity()}:{int(v.is_unsized())}".encode())
        if v.is_unsized():
            h.update(f":{v.hash():016x}".encode())
    return h.hexdigest()[:32]


_PLAIN = (int, float, str, bool, bytes, type(None))


def _code_fingerprint(f, h, seen) -> bool:
    """Feeds everything that decides what ``f`` traces into ``h``: its bytecode, constants and names,
    the code of nested functions, plain values in its closure and the plain globals / helper
    functions it names.  Returns False when something it depends on cannot be fingerprinted (a
    closure over an arbitrary object): such a function only persists under an explicit ``name``."""
    code = getattr(f, "__code__", None)
    if code is None or id(f) in seen:
        return code is not None
    seen.add(id(f))

    def feed_code(c):
        h.update(c.co_code)
        h.update(repr(c.co_names).encode())
        for k in c.co_consts:
            if isinstance(k, types.CodeType):
                feed_code(k)
            else:
                h.update(repr(k).encode())

    feed_code(code)
    ok = True
    for cell in getattr(f, "__closure__", None) or ():
        try:
            val = cell.cell_contents
        except ValueError:
            continue
        if isinstance(val, _PLAIN) or (isinstance(val, tuple) and all(isinstance(x, _PLAIN) for x in val)):
            h.update(repr(val).encode())
        elif isinstance(val, types.FunctionType):
            ok = _code_fingerprint(val, h, seen) and ok
        else:
            ok = False
    g = getattr(f, "__globals__", {})
    for nm in code.co_names:
        val = g.get(nm)
        if isinstance(val, types.FunctionType):
            ok = _code_fingerprint(val, h, seen) and ok
        elif isinstance(val, _PLAIN) and nm in g:
            h.update(f"{nm}={val!r}".encode())
    return ok


def _layout_to_json(layout):
    """JSON form of an output layout, or None when it holds a user type ('obj': those construct
    through a class only this process can name) or a dict key JSON cannot carry."""
    if layout in ("v", "none"):
        return layout
    kind = layout[0]
    if kind in ("list", "tuple"):
        items = [_layout_to_json(l) for l in layout[1]]
        return None if any(i is None for i in items) else [kind, items]
    if kind == "dict":
        items = []
        for k, l in layout[1]:
            j = _layout_to_json(l)
            if j is None or not isinstance(k, (str, int)) or isinstance(k, bool):
                return None
            items.append([["s", k] if isinstance(k, str) else ["i", k], j])
        return ["dict", items]
    return None


def _layout_from_json(j):
    if j in ("v", "none"):
        return j
    if not isinstance(j, list) or len(j) != 2 or j[0] not in ("list", "tuple", "dict") or not isinstance(j[1], list):
        raise ValueError("malformed layout")
    if j[0] == "dict":
        out = []
        for item in j[1]:
            (kind, key), sub = item
            if kind not in ("s", "i") or not isinstance(key, (str, int)):
                raise ValueError("malformed layout key")
            out.append((key, _layout_from_json(sub)))
        return ("dict", out)
    return (j[0], [_layout_from_json(x) for x in j[1]])


def _count_vars(layout) -> int:
    if layout == "v":
        return 1
    if layout == "none":
        return 0
    return sum(_count_vars(l if layout[0] != "dict" else l[1]) for l in layout[1])


def record(f, cache_dir: str | None = None, name: str | None = None):
    """``record(f)``.  With ``cache_dir`` the compiled graphs also persist on disk (graph wire format,
    csrc/tgraph_io.cpp, next to the cubin cache of csrc/jit.cpp): a later process that records a
    function under the same ``name`` with the same input signature launches the stored graph without
    tracing, scheduling or compiling.  Buffers the function captured are stored with the contents
    they had when the graph was written.

    ``name`` defaults to the function's qualified name plus a fingerprint of everything that decides
    its trace and can be seen from here: its bytecode, constants and names, nested functions, plain
    closure values, and the plain globals / helper functions it refers to by name.  A function that
    closes over anything else (an object, a VarRef) is NOT persisted unless ``name`` is given; with
    an explicit ``name`` the caller owns the versioning — change it whenever the function, a callee
    or a captured value changes.  The output layout is stored as JSON next to the graph (nothing on
    disk is ever unpickled); outputs that hold user types stay in memory only."""
    fn_id = next(_fn_ids)  # TypeId::of::<F>() (record.rs:157)
    layouts = {}
    persist = cache_dir is not None
    if name is None:
        fp = hashlib.sha256()
        if not _code_fingerprint(f, fp, set()):
            persist = False
        name = f"{getattr(f, '__module__', '')}.{getattr(f, '__qualname__', 'fn')}:{fp.hexdigest()[:16]}"

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
        if graph is None and persist:
            stable = _stable_key(name, in_layout, flat)
            path = os.path.join(cache_dir, stable + ".hjgraph") if stable else None
            if path and os.path.exists(path) and os.path.exists(path + ".layout"):
                try:
                    with open(path, "rb") as fh:
                        graph = tr.Graph.deserialize(fh.read(), device)
                    with open(path + ".layout", "r", encoding="utf-8") as fh:
                        lay = _layout_from_json(json.load(fh))
                    if _count_vars(lay) != graph.n_outputs():
                        raise ValueError("layout does not match the stored graph")
                    layouts[key] = lay
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
                    lay = _layout_to_json(layouts[key])
                    if lay is None:
                        raise TypeError("output layout holds a user type")
                    blob, lay = graph.serialize(), json.dumps(lay).encode()
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
