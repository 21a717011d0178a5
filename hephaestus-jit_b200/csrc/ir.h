// ir.h — C++ view over the flat hj_ir (mirror of hephaestus-jit/src/ir.rs:12-46) plus the
// VarType layout rules (hephaestus-jit/src/vartype.rs:125-189).
#pragma once
#include <cstdint>
#include <climits>
#include <string>
#include <vector>

#include "../../include/hj.h"

namespace hj {

struct IRView {
    const hj_ir* ir;
    explicit IRView(const hj_ir* p) : ir(p) {}
    uint32_t n_vars() const { return ir->n_vars; }
    const hj_ir_var& var(uint32_t i) const { return ir->vars[i]; }
    uint32_t n_deps(uint32_t i) const { return ir->vars[i].dep_end - ir->vars[i].dep_start; }
    uint32_t dep(uint32_t i, uint32_t k) const { return ir->deps[ir->vars[i].dep_start + k]; }
    const hj_type_desc& type(uint32_t t) const { return ir->types[t]; }
    uint32_t field(const hj_type_desc& t, uint32_t k) const { return ir->struct_fields[t.first_field + k]; }
    uint32_t var_type(uint32_t i) const { return ir->vars[i].ty; }
    hj_type_kind kind(uint32_t t) const { return (hj_type_kind)ir->types[t].kind; }
};

// Structural validation (indices in range, dep counts per op, no out-of-scope ops).
// Returns an empty string when the IR is well formed, else a description of the problem.
std::string validate_ir(const hj_ir* ir);

// How one kernel touches each of its buffer slots — what decides whether two slots may be bound to
// the SAME buffer (Graph::launch_with's lifetime aliasing hands a buffer released in pass p to a
// resource first used in pass p, graph.rs:237-296; hj_kernel_launch rejects every other duplicate).
struct SlotAccess {
    bool read = false, written = false;
    // every access is a plain Gather / Scatter at the bare Index variable, outside loops and ifs
    bool identity = true;
    // every access uses the bare Index variable as its index (loops / ifs / conditions allowed): the
    // slot can hold one contiguous shard of a global array (graph_exec.cpp, sharded execution)
    bool index_only = true;
    bool any_index = false;              // some access uses the bare Index variable
    bool cond_gather = false;            // some Gather carries a condition (inactive lanes read as 0)
    int64_t last_read = -1;              // position (var index) of the last Gather
    int64_t first_write = INT64_MAX;     // position of the first Scatter
};
// `ir` must have passed validate_ir.
void analyse_slot_access(const hj_ir* ir, std::vector<SlotAccess>* out);
// A slot that is only read and a slot that is only written may share one buffer when each thread
// reads element `Index` of the first before it writes element `Index` of the second and nothing else:
// no thread ever sees another thread's store.
inline bool slots_may_share(const SlotAccess& r, const SlotAccess& w) {
    return r.identity && w.identity && !r.written && !w.read && r.last_read < w.first_write;
}

// A DynSize kernel that runs over a per-rank compacted SEGMENT (the index output of a sharded Compress,
// bound to `seg_slot`): how it reaches each of its slots.  The segment holds GLOBAL indices that fall into
// the rank's own block, so a slot addressed only through `Gather(segment, Index)` can be that block.
struct SegmentAccess {
    bool read = false, written = false;
    bool index_only = true;       // every access at the bare Index: the slot is aligned with the segment
    bool through_segment = true;  // every access index is Gather(BufferRef(seg_slot), Index [, literal true])
};
// false (+ why) when the kernel cannot run per segment: the segment is written or not read at the bare
// Index.  (KernelOp::Index as a VALUE is the position in the global compacted sequence: the launcher hands
// the kernel the rank's offset through the count buffer, codegen.cpp: HJ_SIZE.)  `ir` must have passed validate_ir.
bool analyse_segment_access(const hj_ir* ir, uint32_t seg_slot, std::vector<SegmentAccess>* out, std::string* why);

// vartype.rs:125-189
size_t type_size(const IRView& v, uint32_t t);
size_t type_align(const IRView& v, uint32_t t);
size_t struct_offset(const IRView& v, uint32_t t, uint32_t elem);

inline bool is_scalar_kind(uint32_t k) { return k >= HJ_BOOL && k <= HJ_F64; }
inline bool is_float_kind(uint32_t k) { return k == HJ_F16 || k == HJ_F32 || k == HJ_F64; }
inline bool is_int_kind(uint32_t k) { return k >= HJ_I8 && k <= HJ_U64; }
inline bool is_signed_kind(uint32_t k) { return k == HJ_I8 || k == HJ_I16 || k == HJ_I32 || k == HJ_I64; }

// Debug text of an IR in the format of `impl Debug for IR` (ir.rs:47-90), used by the
// snapshot-parity tests.
std::string ir_debug_string(const hj_ir* ir);

struct CodegenResult {
    std::string source;
    bool has_vec_entry = false;  // "hj_kernel_vec" present (all stores are Index-addressed)
    uint32_t vec = 1;            // elements per vector access
    uint32_t unroll = 1;         // vectors per thread
    uint32_t threads = 256;
    bool uses_f16 = false;
    // per buffer slot: bit 0 = only read through the bare Index (staged loads), bit 1 = only
    // written through the bare Index at top level (staged stores); 0 = accessed some other way
    std::vector<uint8_t> slot_flags;
    std::vector<uint32_t> slot_elem_bytes;
};
// Lower `ir` to CUDA C++; on failure returns false and sets `err`.
bool codegen_cuda(const hj_ir* ir, CodegenResult* out, std::string* err);

uint64_t hash_bytes(const void* data, size_t n, uint64_t seed = 0xcbf29ce484222325ull);

}  // namespace hj
