// ir.cpp — validation, layout rules, hashing and debug printing of the flat fused-kernel IR.
#include "ir.h"

#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <sstream>

namespace hj {

uint64_t hash_bytes(const void* data, size_t n, uint64_t h) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; i++) {
        h ^= p[i];
        h *= 0x100000001b3ull;  // FNV-1a 64
    }
    return h;
}

void analyse_slot_access(const hj_ir* ir, std::vector<SlotAccess>* out) {
    IRView v(ir);
    out->assign(ir->n_buffers, SlotAccess());
    int depth = 0;
    for (uint32_t i = 0; i < ir->n_vars; i++) {
        const hj_ir_var& var = v.var(i);
        switch (var.op) {
        case HJ_OP_LOOP_START: case HJ_OP_IF_START: depth++; break;
        case HJ_OP_LOOP_END: case HJ_OP_IF_END: depth--; break;
        case HJ_OP_GATHER: {
            const hj_ir_var& buf = v.var(v.dep(i, 0));
            if (buf.op != HJ_OP_BUFFER_REF || buf.data >= ir->n_buffers) break;
            SlotAccess& s = (*out)[buf.data];
            s.read = true;
            s.last_read = i;
            if (v.n_deps(i) > 2) s.cond_gather = true;
            if (v.var(v.dep(i, 1)).op != HJ_OP_INDEX) s.index_only = false;
            else s.any_index = true;
            if (v.var(v.dep(i, 1)).op != HJ_OP_INDEX || depth != 0) s.identity = false;
            break;
        }
        case HJ_OP_SCATTER: case HJ_OP_SCATTER_REDUCE: case HJ_OP_SCATTER_ATOMIC: case HJ_OP_ATOMIC_INC: {
            const hj_ir_var& buf = v.var(v.dep(i, 0));
            if (buf.op != HJ_OP_BUFFER_REF || buf.data >= ir->n_buffers) break;
            SlotAccess& s = (*out)[buf.data];
            s.written = true;
            if (s.first_write > (int64_t)i) s.first_write = i;
            // AtomicInc: deps = [dst, idx, active] — a counter, never a shard
            if (var.op == HJ_OP_ATOMIC_INC || v.var(v.dep(i, 2)).op != HJ_OP_INDEX) s.index_only = false;
            else s.any_index = true;
            if (var.op != HJ_OP_SCATTER) { s.read = true; s.identity = false; break; }  // atomics read-modify-write
            if (v.var(v.dep(i, 2)).op != HJ_OP_INDEX || depth != 0) s.identity = false;
            break;
        }
        default: break;
        }
    }
}

bool analyse_segment_access(const hj_ir* ir, uint32_t seg_slot, std::vector<SegmentAccess>* out, std::string* why) {
    IRView v(ir);
    out->assign(ir->n_buffers, SegmentAccess());
    auto is_index = [&](uint32_t id) { return v.var(id).op == HJ_OP_INDEX; };
    auto is_segment_gather = [&](uint32_t id) {
        const hj_ir_var& g = v.var(id);
        if (g.op != HJ_OP_GATHER) return false;
        const hj_ir_var& buf = v.var(v.dep(id, 0));
        if (buf.op != HJ_OP_BUFFER_REF || buf.data != seg_slot || !is_index(v.dep(id, 1))) return false;
        if (v.n_deps(id) == 2) return true;
        const hj_ir_var& c = v.var(v.dep(id, 2));  // an inactive lane would read index 0: only `true` is safe
        return c.op == HJ_OP_LITERAL && c.data != 0;
    };
    for (uint32_t i = 0; i < ir->n_vars; i++) {
        const hj_ir_var& var = v.var(i);
        int idx_pos = -1;  // which dependency is the access index
        switch (var.op) {
        case HJ_OP_GATHER: case HJ_OP_ATOMIC_INC: idx_pos = 1; break;
        case HJ_OP_SCATTER: case HJ_OP_SCATTER_REDUCE: case HJ_OP_SCATTER_ATOMIC: idx_pos = 2; break;
        default: break;
        }
        if (idx_pos < 0) continue;
        const hj_ir_var& buf = v.var(v.dep(i, 0));
        if (buf.op != HJ_OP_BUFFER_REF || buf.data >= ir->n_buffers) continue;
        SegmentAccess& s = (*out)[buf.data];
        if (var.op == HJ_OP_GATHER) s.read = true;
        else s.written = true;
        if (var.op != HJ_OP_GATHER && var.op != HJ_OP_SCATTER) s.read = true;  // atomics read-modify-write
        const uint32_t idx = v.dep(i, (uint32_t)idx_pos);
        if (!is_index(idx)) s.index_only = false;
        if (!is_segment_gather(idx)) s.through_segment = false;
    }
    if (seg_slot >= ir->n_buffers) {
        if (why) *why = "segment slot out of range";
        return false;
    }
    const SegmentAccess& seg = (*out)[seg_slot];
    if (seg.written || !seg.read || !seg.index_only) {
        if (why) *why = "the index segment must only be read, at the bare Index";
        return false;
    }
    return true;
}

static bool dep_count_ok(uint32_t op, uint32_t n, const char** want) {
    switch (op) {
    case HJ_OP_NOP: *want = "1"; return n == 1;
    case HJ_OP_SCATTER: case HJ_OP_SCATTER_REDUCE: case HJ_OP_SCATTER_ATOMIC: *want = "3 or 4"; return n == 3 || n == 4;
    case HJ_OP_ATOMIC_INC: *want = "3"; return n == 3;
    case HJ_OP_GATHER: *want = "2 or 3"; return n == 2 || n == 3;
    case HJ_OP_INDEX: case HJ_OP_LITERAL: case HJ_OP_BUFFER_REF: *want = "0"; return n == 0;
    case HJ_OP_EXTRACT: *want = "1"; return n == 1;
    case HJ_OP_DYN_EXTRACT: *want = "2"; return n == 2;
    case HJ_OP_CONSTRUCT: *want = ">= 1"; return n >= 1;
    case HJ_OP_SELECT: case HJ_OP_FMA: *want = "3"; return n == 3;
    case HJ_OP_LOOP_START: case HJ_OP_IF_START: *want = "1"; return n == 1;
    case HJ_OP_LOOP_END: case HJ_OP_IF_END: *want = ">= 2"; return n >= 2;
    case HJ_OP_BOP: *want = "2"; return n == 2;
    case HJ_OP_UOP: *want = "1"; return n == 1;
    default: *want = "?"; return true;
    }
}

std::string validate_ir(const hj_ir* ir) {
    char buf[256];
    if (!ir) return "null IR";
    if (ir->n_vars && !ir->vars) return "IR: vars is null";
    if (ir->n_deps && !ir->deps) return "IR: deps is null";
    if (!ir->types || !ir->n_types) return "IR: empty type table";
    for (uint32_t t = 0; t < ir->n_types; t++) {
        const hj_type_desc& d = ir->types[t];
        if (d.kind > HJ_STRUCT) { snprintf(buf, sizeof(buf), "type %u: unknown kind %u", t, d.kind); return buf; }
        if ((d.kind == HJ_VEC || d.kind == HJ_ARRAY || d.kind == HJ_MAT) && d.elem >= t) {
            snprintf(buf, sizeof(buf), "type %u: element type %u must precede it", t, d.elem); return buf;
        }
        if ((d.kind == HJ_VEC || d.kind == HJ_ARRAY) && d.num == 0) { snprintf(buf, sizeof(buf), "type %u: zero length", t); return buf; }
        if (d.kind == HJ_MAT && (d.cols == 0 || d.rows == 0)) { snprintf(buf, sizeof(buf), "type %u: empty matrix", t); return buf; }
        if (d.kind == HJ_STRUCT) {
            if (d.num == 0) { snprintf(buf, sizeof(buf), "type %u: empty struct", t); return buf; }
            if ((uint64_t)d.first_field + d.num > ir->n_struct_fields) { snprintf(buf, sizeof(buf), "type %u: struct fields out of range", t); return buf; }
            for (uint32_t k = 0; k < d.num; k++)
                if (ir->struct_fields[d.first_field + k] >= t) { snprintf(buf, sizeof(buf), "type %u: field type must precede it", t); return buf; }
        }
    }
    int depth = 0;
    for (uint32_t i = 0; i < ir->n_vars; i++) {
        const hj_ir_var& v = ir->vars[i];
        if (v.ty >= ir->n_types) { snprintf(buf, sizeof(buf), "var%u: type index %u out of range", i, v.ty); return buf; }
        if (v.dep_start > v.dep_end || v.dep_end > ir->n_deps) { snprintf(buf, sizeof(buf), "var%u: dependency range out of bounds", i); return buf; }
        for (uint32_t k = v.dep_start; k < v.dep_end; k++)
            if (ir->deps[k] >= i) { snprintf(buf, sizeof(buf), "var%u: depends on var%u which is not defined before it", i, ir->deps[k]); return buf; }
        const char* want = "";
        if (!dep_count_ok(v.op, v.dep_end - v.dep_start, &want)) {
            snprintf(buf, sizeof(buf), "var%u: op %u takes %s dependencies, got %u", i, v.op, want, v.dep_end - v.dep_start); return buf;
        }
        if (v.op == HJ_OP_BUFFER_REF && v.data >= ir->n_buffers) { snprintf(buf, sizeof(buf), "var%u: buffer slot %" PRIu64 " >= n_buffers %u", i, v.data, ir->n_buffers); return buf; }
        if (v.op == HJ_OP_TEX_LOOKUP || v.op == HJ_OP_TRACE_RAY || v.op == HJ_OP_TEXTURE_REF || v.op == HJ_OP_ACCEL_REF)
            return "texture / ray-tracing ops are out of scope for the B200 backend";
        if (v.op > HJ_OP_ACCEL_REF) { snprintf(buf, sizeof(buf), "var%u: unknown op %u", i, v.op); return buf; }
        if (v.op == HJ_OP_LOOP_START || v.op == HJ_OP_IF_START) {
            depth++;
            const hj_type_desc& d = ir->types[v.ty];
            if (d.kind != HJ_STRUCT || ir->types[ir->struct_fields[d.first_field]].kind != HJ_BOOL)
                return "loop/if state must be a struct whose first field is the bool condition (trace.rs:408-464)";
        }
        if (v.op == HJ_OP_LOOP_END || v.op == HJ_OP_IF_END) {
            if (--depth < 0) return "LoopEnd/IfEnd without a matching start";
            uint32_t start = ir->deps[v.dep_start];
            if (ir->vars[start].op != HJ_OP_LOOP_START && ir->vars[start].op != HJ_OP_IF_START)
                return "LoopEnd/IfEnd: first dependency must be the start node";
        }
    }
    if (depth != 0) return "unterminated loop/if";
    return "";
}

size_t type_align(const IRView& v, uint32_t t) {
    const hj_type_desc& d = v.type(t);
    switch (d.kind) {
    case HJ_VOID: return 0;
    case HJ_BOOL: case HJ_I8: case HJ_U8: return 1;
    case HJ_I16: case HJ_U16: case HJ_F16: return 2;
    case HJ_I32: case HJ_U32: case HJ_F32: return 4;
    case HJ_I64: case HJ_U64: case HJ_F64: return 8;
    case HJ_VEC: case HJ_ARRAY: return type_align(v, d.elem);
    case HJ_MAT: return type_size(v, d.elem) * d.rows;  // vartype.rs:186
    case HJ_STRUCT: {
        size_t a = 0;
        for (uint32_t k = 0; k < d.num; k++) a = std::max(a, type_align(v, v.field(d, k)));
        return a;
    }
    default: return 0;
    }
}

static size_t align_up(size_t x, size_t a) { return a ? (x + a - 1) / a * a : x; }

size_t struct_offset(const IRView& v, uint32_t t, uint32_t elem) {  // vartype.rs:156-168
    const hj_type_desc& d = v.type(t);
    size_t off = 0;
    for (uint32_t k = 0; k < elem; k++) {
        off += type_size(v, v.field(d, k));
        off = align_up(off, type_align(v, v.field(d, k + 1)));
    }
    return off;
}

size_t type_size(const IRView& v, uint32_t t) {  // vartype.rs:125-155
    const hj_type_desc& d = v.type(t);
    switch (d.kind) {
    case HJ_VOID: return 0;
    case HJ_BOOL: case HJ_I8: case HJ_U8: return 1;
    case HJ_I16: case HJ_U16: case HJ_F16: return 2;
    case HJ_I32: case HJ_U32: case HJ_F32: return 4;
    case HJ_I64: case HJ_U64: case HJ_F64: return 8;
    case HJ_VEC: case HJ_ARRAY: return type_size(v, d.elem) * d.num;
    case HJ_MAT: return type_size(v, d.elem) * d.cols * d.rows;
    case HJ_STRUCT: {
        size_t off = struct_offset(v, t, d.num - 1);
        return align_up(off + type_size(v, v.field(d, d.num - 1)), type_align(v, t));
    }
    default: return 0;
    }
}

// ---- Debug formatting identical to the reference's derived/handwritten Debug impls ----------
static std::string type_debug(const IRView& v, uint32_t t) {
    static const char* names[] = {"Void", "Bool", "I8", "U8", "I16", "U16", "I32", "U32", "I64", "U64", "F16", "F32", "F64"};
    const hj_type_desc& d = v.type(t);
    std::ostringstream s;
    switch (d.kind) {
    case HJ_VEC: s << "Vec { ty: " << type_debug(v, d.elem) << ", num: " << d.num << " }"; break;
    case HJ_ARRAY: s << "Array { ty: " << type_debug(v, d.elem) << ", num: " << d.num << " }"; break;
    case HJ_MAT: s << "Mat { ty: " << type_debug(v, d.elem) << ", rows: " << d.rows << ", cols: " << d.cols << " }"; break;
    case HJ_STRUCT:
        s << "Struct { tys: [";
        for (uint32_t k = 0; k < d.num; k++) s << (k ? ", " : "") << type_debug(v, v.field(d, k));
        s << "] }";
        break;
    default: s << names[d.kind]; break;
    }
    return s.str();
}

static std::string op_debug(const hj_ir_var& var) {
    static const char* rops[] = {"Max", "Min", "Sum", "Prod", "Or", "And", "Xor"};
    static const char* bops[] = {"Add", "Sub", "Mul", "Div", "Modulus", "Min", "Max", "Inner", "And", "Or", "Xor",
                                 "Shl", "Shr", "Eq", "Neq", "Lt", "Le", "Gt", "Ge"};
    static const char* uops[] = {"Cast", "BitCast", "Neg", "Sqrt", "Abs", "Sin", "Cos", "Exp2", "Log2"};
    std::ostringstream s;
    switch (var.op) {
    case HJ_OP_NOP: return "Nop";
    case HJ_OP_SCATTER: return "Scatter";
    case HJ_OP_SCATTER_REDUCE: s << "ScatterReduce(" << rops[var.arg % 7] << ")"; return s.str();
    case HJ_OP_SCATTER_ATOMIC: s << "ScatterAtomic(" << rops[var.arg % 7] << ")"; return s.str();
    case HJ_OP_ATOMIC_INC: return "AtomicInc";
    case HJ_OP_GATHER: return "Gather";
    case HJ_OP_INDEX: return "Index";
    case HJ_OP_LITERAL: return "Literal";
    case HJ_OP_EXTRACT: s << "Extract(" << var.arg << ")"; return s.str();
    case HJ_OP_DYN_EXTRACT: return "DynExtract";
    case HJ_OP_CONSTRUCT: return "Construct";
    case HJ_OP_SELECT: return "Select";
    case HJ_OP_LOOP_START: return "LoopStart";
    case HJ_OP_LOOP_END: return "LoopEnd";
    case HJ_OP_IF_START: return "IfStart";
    case HJ_OP_IF_END: return "IfEnd";
    case HJ_OP_BOP: s << "Bop(" << bops[var.arg % 19] << ")"; return s.str();
    case HJ_OP_UOP: s << "Uop(" << uops[var.arg % 9] << ")"; return s.str();
    case HJ_OP_FMA: return "FMA";
    case HJ_OP_BUFFER_REF: return "BufferRef";
    default: return "?";
    }
}

// `impl Debug for IR` (ir.rs:47-90) in `{:#?}` form as stored in the insta snapshots.
std::string ir_debug_string(const hj_ir* ir) {
    IRView v(ir);
    std::ostringstream s;
    s << "IR {\n    vars: [\n";
    for (uint32_t i = 0; i < ir->n_vars; i++) {
        const hj_ir_var& var = ir->vars[i];
        s << "    \tvar" << i << ": " << type_debug(v, var.ty) << " = " << op_debug(var) << "(";
        uint32_t n = var.dep_end - var.dep_start;
        for (uint32_t k = 0; k < n; k++) s << "var" << ir->deps[var.dep_start + k] << (k + 1 < n ? ", " : "");
        if (var.op == HJ_OP_BUFFER_REF || var.op == HJ_OP_LITERAL) s << var.data;
        s << ")\n";
    }
    s << "    ],\n    n_buffers: " << ir->n_buffers << ",\n    n_textures: 0,\n    n_accels: 0,\n}";
    return s.str();
}

}  // namespace hj

extern "C" uint64_t hj_ir_hash(const hj_ir* ir) {
    using namespace hj;
    if (!ir) return 0;
    uint64_t h = hash_bytes("hj_ir_v1", 8);
    // vars are hashed field by field: the struct has padding
    for (uint32_t i = 0; i < ir->n_vars; i++) {
        const hj_ir_var& v = ir->vars[i];
        uint32_t f[5] = {v.ty, v.op, v.arg, v.dep_end - v.dep_start, 0};
        h = hash_bytes(f, sizeof(f), h);
        h = hash_bytes(&v.data, 8, h);
        if (v.dep_end > v.dep_start) h = hash_bytes(ir->deps + v.dep_start, 4 * (v.dep_end - v.dep_start), h);
    }
    for (uint32_t t = 0; t < ir->n_types; t++) {
        const hj_type_desc& d = ir->types[t];
        uint32_t f[5] = {d.kind, d.elem, d.num, d.cols, d.rows};
        h = hash_bytes(f, sizeof(f), h);
        if (d.kind == HJ_STRUCT) h = hash_bytes(ir->struct_fields + d.first_field, 4 * d.num, h);
    }
    h = hash_bytes(&ir->n_buffers, 4, h);
    return h;
}
