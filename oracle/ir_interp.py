"""numpy interpreter for the fused-kernel IR (TEST INFRASTRUCTURE — see oracle/hj_oracle.c).

Gives every IR op the meaning the reference's GLSL code generator assigns to it
(hephaestus-jit/src/backend/vulkan/codegen/glsl/mod.rs:336-1035; GLSL 4.60 for the operators),
evaluated for all elements at once.  The reference itself executes these kernels inside a Vulkan
driver and cannot run here; what pins this interpreter are the reference's own
elementwise / scatter / gather / loop known-answer tests (tests/golden/reference_kats.json ->
"programs"), checked in tests/test_ir_interp.py.

Input is the flat IR (same arrays as ``hj_ir`` in include/hj.h): ``types`` =
[(kind, elem, num, cols, rows, first_field)], ``struct_fields``, ``vars`` =
[(ty, op, arg, dep_start, dep_end, data)], ``deps``.  Buffers are numpy arrays in MEMORY layout
(bool as uint8, Vec/Array of n scalars as an (N, n) array) and are updated in place.

Float transcendentals are evaluated in float64 and rounded once (the correctly rounded f32 up
to double rounding); FMA is the IEEE fused multiply-add (deviation D1: the reference emits
nothing for it).  Side effects are applied in element order, which matches the GPU whenever the
result does not depend on the order (the only cases the tests compare exactly).
"""
from __future__ import annotations

import numpy as np

(VOID, BOOL, I8, U8, I16, U16, I32, U32, I64, U64, F16, F32, F64, VEC, ARRAY, MAT, STRUCT) = range(17)
(OP_NOP, OP_SCATTER, OP_SCATTER_REDUCE, OP_SCATTER_ATOMIC, OP_ATOMIC_INC, OP_GATHER, OP_INDEX,
 OP_LITERAL, OP_EXTRACT, OP_DYN_EXTRACT, OP_CONSTRUCT, OP_SELECT, OP_LOOP_START, OP_LOOP_END,
 OP_IF_START, OP_IF_END, OP_TEX_LOOKUP, OP_TRACE_RAY, OP_BOP, OP_UOP, OP_FMA, OP_BUFFER_REF,
 OP_TEXTURE_REF, OP_ACCEL_REF) = range(24)
(BOP_ADD, BOP_SUB, BOP_MUL, BOP_DIV, BOP_MODULUS, BOP_MIN, BOP_MAX, BOP_INNER, BOP_AND, BOP_OR,
 BOP_XOR, BOP_SHL, BOP_SHR, BOP_EQ, BOP_NEQ, BOP_LT, BOP_LE, BOP_GT, BOP_GE) = range(19)
(UOP_CAST, UOP_BITCAST, UOP_NEG, UOP_SQRT, UOP_ABS, UOP_SIN, UOP_COS, UOP_EXP2, UOP_LOG2) = range(9)
R_MAX, R_MIN, R_SUM, R_PROD, R_OR, R_AND, R_XOR = range(7)

NP = {BOOL: np.bool_, I8: np.int8, U8: np.uint8, I16: np.int16, U16: np.uint16, I32: np.int32,
      U32: np.uint32, I64: np.int64, U64: np.uint64, F16: np.float16, F32: np.float32,
      F64: np.float64}
MEM = dict(NP)
MEM[BOOL] = np.uint8
_UINT_OF_SIZE = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


def _is_float(k):
    return k in (F16, F32, F64)


def _is_int(k):
    return I8 <= k <= U64


class Interp:
    def __init__(self, types, struct_fields, vars_, deps, n_buffers):
        self.types, self.fields, self.vars, self.deps, self.n_buffers = types, struct_fields, vars_, deps, n_buffers

    # ---- helpers ---------------------------------------------------------------------------
    def kind(self, t):
        return self.types[t][0]

    def dep(self, i, k):
        return self.deps[self.vars[i][3] + k]

    def ndeps(self, i):
        return self.vars[i][4] - self.vars[i][3]

    def zero(self, t, n):
        k = self.kind(t)
        if k <= F64:
            return np.zeros(n, dtype=NP[k])
        if k in (VEC, ARRAY):
            return [self.zero(self.types[t][1], n) for _ in range(self.types[t][2])]
        if k == MAT:
            return [self.zero(self.types[t][1], n) for _ in range(self.types[t][3] * self.types[t][4])]
        first = self.types[t][5]
        return [self.zero(self.fields[first + j], n) for j in range(self.types[t][2])]

    @staticmethod
    def _where(mask, a, b):
        if isinstance(a, list):
            return [Interp._where(mask, x, y) for x, y in zip(a, b)]
        return np.where(mask, a, b)

    @staticmethod
    def _bcast(x, n):
        if isinstance(x, list):
            return [Interp._bcast(c, n) for c in x]
        x = np.asarray(x)
        return np.broadcast_to(x, (n,)).copy() if x.ndim == 0 else x

    # ---- scalar op tables -------------------------------------------------------------------
    def bop(self, op, k, a, b):
        dt = NP[k]
        with np.errstate(all="ignore"):
            if op == BOP_ADD:
                return (a + b).astype(dt)
            if op == BOP_SUB:
                return (a - b).astype(dt)
            if op in (BOP_MUL, BOP_INNER):
                return (a * b).astype(dt)
            if op == BOP_DIV:
                if _is_float(k):
                    return (a / b).astype(dt)
                return self._int_div(a, b, dt)
            if op == BOP_MODULUS:
                if _is_float(k):
                    return np.fmod(a, b).astype(dt)
                q = self._int_div(a, b, dt)
                return (a - q * b).astype(dt)
            if op == BOP_MIN:
                return (np.fmin(a, b) if _is_float(k) else np.minimum(a, b)).astype(dt)
            if op == BOP_MAX:
                return (np.fmax(a, b) if _is_float(k) else np.maximum(a, b)).astype(dt)
            if op == BOP_AND:
                return np.logical_and(a, b) if k == BOOL else (a & b).astype(dt)
            if op == BOP_OR:
                return np.logical_or(a, b) if k == BOOL else (a | b).astype(dt)
            if op == BOP_XOR:
                return np.logical_xor(a, b) if k == BOOL else (a ^ b).astype(dt)
            if op == BOP_SHL:
                return np.left_shift(a, b.astype(dt)).astype(dt)
            if op == BOP_SHR:
                return np.right_shift(a, b.astype(dt)).astype(dt)
            if op == BOP_EQ:
                return a == b
            if op == BOP_NEQ:
                return a != b
            if op == BOP_LT:
                return a < b
            if op == BOP_LE:
                return a <= b
            if op == BOP_GT:
                return a > b
            if op == BOP_GE:
                return a >= b
        raise ValueError(f"bop {op}")

    @staticmethod
    def _int_div(a, b, dt):
        """C / GLSL integer division: truncation toward zero (numpy's // floors)."""
        a = np.asarray(a)
        b = np.asarray(b)
        bb = np.where(b == 0, 1, b)
        if np.issubdtype(dt, np.unsignedinteger):
            q = a // bb
        else:
            q = np.abs(a.astype(np.int64)) // np.abs(bb.astype(np.int64))
            q = np.where((a < 0) != (bb < 0), -q, q)
        return np.where(b == 0, 0, q).astype(dt)

    def uop(self, op, k, x):
        dt = NP[k]
        with np.errstate(all="ignore"):
            if op == UOP_NEG:
                return np.logical_not(x) if k == BOOL else (-x).astype(dt) if not np.issubdtype(dt, np.unsignedinteger) else (0 - x).astype(dt)
            if op == UOP_ABS:
                return np.abs(x).astype(dt) if not np.issubdtype(dt, np.unsignedinteger) else x
            wide = x.astype(np.float64)
            if op == UOP_SQRT:
                return np.sqrt(x).astype(dt)  # correctly rounded in the native type
            if op == UOP_SIN:
                return np.sin(wide).astype(dt)
            if op == UOP_COS:
                return np.cos(wide).astype(dt)
            if op == UOP_EXP2:
                return np.exp2(wide).astype(dt)
            if op == UOP_LOG2:
                return np.log2(wide).astype(dt)
        raise ValueError(f"uop {op}")

    def cast(self, dk, sk, x):
        if dk == sk:
            return x
        with np.errstate(all="ignore"):
            if dk == BOOL:
                return x != 0
            if _is_float(sk) and _is_int(dk):
                x = np.trunc(x.astype(np.float64))
                info = np.iinfo(NP[dk])
                x = np.clip(np.nan_to_num(x, nan=0.0), info.min, info.max)  # CUDA saturates
                return x.astype(NP[dk])
            return x.astype(NP[dk])

    def cast_any(self, dt, st, x):
        dk, sk = self.kind(dt), self.kind(st)
        if dk <= F64 and sk <= F64:
            return self.cast(dk, sk, x)
        if dk in (VEC, ARRAY) and sk in (VEC, ARRAY):
            de, se = self.kind(self.types[dt][1]), self.kind(self.types[st][1])
            return [self.cast(de, se, c) for c in x]
        if dk == STRUCT and sk == STRUCT:
            df, sf = self.types[dt][5], self.types[st][5]
            return [self.cast_any(self.fields[df + j], self.fields[sf + j], c) for j, c in enumerate(x)]
        raise NotImplementedError("cast between these types is todo!() in the reference")

    # ---- memory ----------------------------------------------------------------------------
    def load(self, buf, t, idx, mask):
        k = self.kind(t)
        n = idx.shape[0]
        safe = np.where(mask, idx, 0).astype(np.int64)
        if k <= F64:
            v = buf.reshape(-1)[safe]
            v = (v != 0) if k == BOOL else v.astype(NP[k])
            return np.where(mask, v, np.zeros(1, dtype=v.dtype))
        if k in (VEC, ARRAY, MAT):  # matrices are stored column by column, packed
            num = self.types[t][2] if k != MAT else self.types[t][3] * self.types[t][4]
            b2 = buf.reshape(-1, num)
            return [np.where(mask, b2[safe, c], 0).astype(b2.dtype) for c in range(num)]
        raise NotImplementedError("gather of struct types")

    def store(self, buf, t, idx, val, mask):
        k = self.kind(t)
        sel = np.flatnonzero(mask)
        if k <= F64:
            flat = buf.reshape(-1)
            v = self._bcast(val, mask.shape[0])
            flat[idx[sel].astype(np.int64)] = v[sel].astype(flat.dtype)
            return
        if k in (VEC, ARRAY, MAT):
            num = self.types[t][2] if k != MAT else self.types[t][3] * self.types[t][4]
            b2 = buf.reshape(-1, num)
            for c in range(num):
                v = self._bcast(val[c], mask.shape[0])
                b2[idx[sel].astype(np.int64), c] = v[sel].astype(b2.dtype)
            return
        raise NotImplementedError("scatter of struct types")

    # ---- execution -------------------------------------------------------------------------
    def run(self, size, buffers, size_buf=None, index_base=0):
        n = int(size)
        if size_buf is not None:
            n = min(n, int(size_buf[0]))
        self.n = n
        self.buffers = buffers
        self.index = np.arange(n, dtype=np.uint32)
        self.gindex = (self.index + np.uint32(index_base)).astype(np.uint32)
        self.val = [None] * len(self.vars)
        if n == 0:
            return
        self.exec_range(0, len(self.vars), np.ones(n, dtype=bool))

    def matching_end(self, start):
        depth = 0
        for j in range(start, len(self.vars)):
            op = self.vars[j][1]
            if op in (OP_LOOP_START, OP_IF_START):
                depth += 1
            elif op in (OP_LOOP_END, OP_IF_END):
                depth -= 1
                if depth == 0:
                    return j
        raise ValueError("unterminated loop")

    def addr(self, idx_var):
        """The bare Index var addresses local memory (glsl/mod.rs:580-582 `index`)."""
        if self.vars[idx_var][1] == OP_INDEX:
            return self.index
        return self._bcast(self.val[idx_var], self.n).astype(np.int64)

    def exec_range(self, lo, hi, active):
        i = lo
        while i < hi:
            ty, op, arg, ds, de, data = self.vars[i]
            if op in (OP_LOOP_START, OP_IF_START):
                end = self.matching_end(i)
                state = self._bcast(self.val[self.dep(i, 0)], self.n)
                state = [np.array(c, copy=True) if not isinstance(c, list) else c for c in state]
                first = True
                while True:
                    cond = np.logical_and(active, state[0].astype(bool))
                    if not cond.any():
                        break
                    self.val[i] = state
                    self.exec_range(i + 1, end, cond)
                    nxt = self._bcast(self.val[self.dep(end, 1)], self.n)
                    state = self._where(cond, nxt, state)
                    if op == OP_IF_START:
                        break
                    first = False
                self.val[i] = state
                self.val[end] = state
                i = end + 1
                continue
            self.val[i] = self.exec_one(i, active)
            i += 1

    def exec_one(self, i, active):
        ty, op, arg, ds, de, data = self.vars[i]
        k = self.kind(ty)
        n = self.n
        d = lambda j: self.val[self.dep(i, j)]
        if op == OP_NOP:
            return d(0)
        if op == OP_BUFFER_REF:
            return ("buffer", int(data))
        if op == OP_INDEX:
            return self.gindex
        if op == OP_LITERAL:
            raw = int(data).to_bytes(8, "little")
            if k == BOOL:
                return np.bool_(data != 0)
            return np.frombuffer(raw, dtype=NP[k], count=1)[0]
        if op == OP_GATHER:
            slot = d(0)[1]
            mask = active.copy()
            if self.ndeps(i) > 2:
                mask &= self._bcast(d(2), n).astype(bool)
            return self.load(self.buffers[slot], ty, self.addr(self.dep(i, 1)), mask)
        if op == OP_SCATTER:
            slot = d(0)[1]
            mask = active.copy()
            if self.ndeps(i) > 3:
                mask &= self._bcast(d(3), n).astype(bool)
            src_t = self.vars[self.dep(i, 1)][0]
            self.store(self.buffers[slot], src_t, self.addr(self.dep(i, 2)), d(1), mask)
            return None
        if op in (OP_SCATTER_REDUCE, OP_SCATTER_ATOMIC):
            if arg == R_PROD:
                raise NotImplementedError("ScatterReduce(Prod) is todo!() in the reference")
            slot = d(0)[1]
            mask = active.copy()
            if self.ndeps(i) > 3:
                mask &= self._bcast(d(3), n).astype(bool)
            flat = self.buffers[slot].reshape(-1)
            idx = self.addr(self.dep(i, 2))
            src = self._bcast(d(1), n).astype(flat.dtype)
            sel = np.flatnonzero(mask)
            if op == OP_SCATTER_REDUCE:
                fn = {R_MAX: np.maximum, R_MIN: np.minimum, R_SUM: np.add, R_OR: np.bitwise_or,
                      R_AND: np.bitwise_and, R_XOR: np.bitwise_xor}[arg]
                with np.errstate(all="ignore"):
                    fn.at(flat, idx[sel].astype(np.int64), src[sel])
                return None
            old = np.zeros(n, dtype=flat.dtype)
            for e in sel:  # returns the previous value: inherently sequential
                a = int(idx[e])
                old[e] = flat[a]
                x, y = flat[a], src[e]
                with np.errstate(all="ignore"):
                    flat[a] = {R_MAX: max(x, y), R_MIN: min(x, y), R_SUM: x + y, R_OR: x | y if arg == R_OR else 0,
                               R_AND: x & y if arg == R_AND else 0, R_XOR: x ^ y if arg == R_XOR else 0}[arg]
            return old
        if op == OP_ATOMIC_INC:
            slot = d(0)[1]
            flat = self.buffers[slot].reshape(-1)
            idx = self.addr(self.dep(i, 1))
            mask = np.logical_and(active, self._bcast(d(2), n).astype(bool))
            out = np.zeros(n, dtype=NP[k])
            for e in np.flatnonzero(mask):
                a = int(idx[e]) if idx.shape[0] > 1 else int(idx[0])
                out[e] = flat[a]
                flat[a] += 1
            return out
        if op == OP_EXTRACT:
            return d(0)[arg]
        if op == OP_DYN_EXTRACT:
            comps = [self._bcast(c, n) for c in d(0)]
            sel = self._bcast(d(1), n).astype(np.int64)
            return np.stack(comps, axis=0)[sel, np.arange(n)]
        if op == OP_CONSTRUCT:
            parts = [d(j) for j in range(self.ndeps(i))]
            if k == MAT:  # columns of `rows` components, column-major
                return [c for col in parts for c in col]
            return parts
        if op == OP_SELECT:
            return self._where(self._bcast(d(0), n).astype(bool), self._bcast(d(1), n), self._bcast(d(2), n))
        if op == OP_BOP:
            at = self.vars[self.dep(i, 0)][0]
            ak = self.kind(at)
            a, b = d(0), d(1)
            if ak <= F64:
                return self.bop(arg, ak, np.asarray(a), np.asarray(b))
            ek = self.kind(self.types[at][1])
            if arg == BOP_INNER and ak == VEC and k <= F64:
                acc = np.zeros(n, dtype=NP[k])
                for x, y in zip(a, b):
                    acc = (acc + (x * y).astype(NP[k])).astype(NP[k])
                return acc
            if arg == BOP_MUL and ak == MAT and self.types[at][3] == self.types[at][4]:
                # GLSL `*` on matrices (glsl/mod.rs emits `a * b`): column-major linear-algebra product
                nn = self.types[at][4]
                out = []
                for c in range(nn):
                    for q in range(nn):
                        acc = np.zeros(n, dtype=NP[ek])
                        for kk in range(nn):
                            acc = (acc + (self._bcast(a[kk * nn + q], n) * self._bcast(b[c * nn + kk], n)).astype(NP[ek])).astype(NP[ek])
                        out.append(acc)
                return out
            if arg in (BOP_EQ, BOP_NEQ):
                eq = np.ones(n, dtype=bool)
                for x, y in zip(a, b):
                    eq &= (self._bcast(x, n) == self._bcast(y, n))
                return eq if arg == BOP_EQ else ~eq
            return [self.bop(arg, ek, np.asarray(x), np.asarray(y)) for x, y in zip(a, b)]
        if op == OP_UOP:
            st = self.vars[self.dep(i, 0)][0]
            sk = self.kind(st)
            x = d(0)
            if arg == UOP_CAST:
                return self.cast_any(ty, st, x)
            if arg == UOP_BITCAST:
                if _is_int(k) and _is_int(sk):
                    return self.cast(k, sk, np.asarray(x))
                x = np.ascontiguousarray(self._bcast(x, n) if np.ndim(x) else np.array([x]))
                out = x.view(NP[k])
                return out if np.ndim(d(0)) else out[0]
            if sk <= F64:
                return self.uop(arg, sk, np.asarray(x))
            ek = self.kind(self.types[st][1])
            return [self.uop(arg, ek, np.asarray(c)) for c in x]
        if op == OP_FMA:
            a, b, c = d(0), d(1), d(2)

            def one(kk, x, y, z):
                x, y, z = np.asarray(x), np.asarray(y), np.asarray(z)
                if kk in (F32, F16):  # x*y is exact in f64; rounded once to f64, once to the type
                    return (x.astype(np.float64) * y.astype(np.float64) + z.astype(np.float64)).astype(NP[kk])
                if kk == F64:
                    import math
                    if hasattr(math, "fma"):
                        return np.frompyfunc(math.fma, 3, 1)(x, y, z).astype(np.float64)
                    return x * y + z
                with np.errstate(all="ignore"):
                    return (x * y + z).astype(NP[kk])
            if k <= F64:
                return one(k, a, b, c)
            ek = self.kind(self.types[ty][1])
            return [one(ek, x, y, z) for x, y, z in zip(a, b, c)]
        raise NotImplementedError(f"op {op} is out of scope")


def run_ir(builder_or_tuple, size, buffers, size_buf=None, index_base=0):
    """Interpret an IR.  ``builder_or_tuple``: an object with .types/.struct_fields/.vars/.deps/
    .n_buffers (e.g. the product's IRBuilder — only its plain lists are read) or that 5-tuple."""
    b = builder_or_tuple
    if not isinstance(b, tuple):
        b = (b.types, b.struct_fields, b.vars, b.deps, b.n_buffers)
    it = Interp(*b)
    it.run(size, buffers, size_buf=size_buf, index_base=index_base)
    return it
