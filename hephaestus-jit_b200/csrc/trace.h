// trace.h — host-side trace / schedule / graph layer (C++ restatement of the reference's
// hephaestus-jit/src/{trace,graph,compiler,extent,op,resource,vartype}.rs for the hot path).
//
// The reference records every array operation as a `Var` in a global ref-counted DAG
// (trace.rs:81-227), keeps a thread-local schedule of variables to evaluate
// (ThreadState, trace.rs:31-69), cuts the schedule into passes (graph.rs:436-614), lowers
// each kernel pass to a flat SSA IR (compiler.rs:20-230) and hands the pass list to the backend
// (graph.rs:192-400).  This file restates those layers so that programs traced against the same
// op vocabulary produce the same Graph / IR and drive the CUDA backend end to end.
// Textures, acceleration structures, MatMul and FusedMlp are out of scope.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/hj.h"

namespace hj {
namespace tr {

// ---- interned type tree (vartype.rs:20-122) --------------------------------------------------
using TypeId = uint32_t;
struct TypeNode {
    uint32_t kind = HJ_VOID;  // hj_type_kind
    TypeId elem = 0;          // Vec / Array / Mat
    uint32_t num = 0;         // Vec / Array length
    uint32_t cols = 0, rows = 0;
    std::vector<TypeId> fields;  // Struct
};
TypeId type_scalar(uint32_t kind);
TypeId type_vector(TypeId elem, uint32_t num);
TypeId type_array(TypeId elem, uint32_t num);
TypeId type_matrix(TypeId elem, uint32_t cols, uint32_t rows);
TypeId type_struct(const TypeId* fields, uint32_t n);
TypeNode type_node(TypeId t);
size_t type_size(TypeId t);       // vartype.rs:125-155
size_t type_alignment(TypeId t);  // vartype.rs:169-189
int type_num_elements(TypeId t);  // vartype.rs:190-198, -1 for scalars
std::string type_debug(TypeId t);

// ---- ids ------------------------------------------------------------------------------------
// slot index in the low 32 bits, generation in the high 32 bits (slotmap::DefaultKey)
using VarId = uint64_t;
constexpr VarId NO_VAR = 0;

enum class OpKind : uint8_t { Nop, Ref, Buffer, DeviceOp, KernelOp };  // op.rs:145-160
enum DeviceOpKind : uint32_t { DOP_REDUCE = 0, DOP_PREFIX_SUM = 1, DOP_COMPRESS = 2 };  // op.rs:104-130 (in scope)

struct Op {
    OpKind kind = OpKind::Nop;
    bool ref_mutable = false;
    uint32_t code = 0;  // KernelOp: hj_kernel_op; DeviceOp: DeviceOpKind
    uint32_t arg = 0;   // KernelOp payload (Bop/Uop/ReduceOp/element); DeviceOp: reduce op / inclusive flag
    bool operator==(const Op& o) const {
        return kind == o.kind && ref_mutable == o.ref_mutable && code == o.code && arg == o.arg;
    }
};

struct Extent {  // extent.rs:6-19 (Size / DynSize; Texture / Accel out of scope)
    bool dynamic = false;
    size_t n = 0;           // Size: size; DynSize: capacity
    VarId size_var = NO_VAR;
    bool operator==(const Extent& o) const { return dynamic == o.dynamic && n == o.n && size_var == o.size_var; }
    bool operator!=(const Extent& o) const { return !(*this == o); }
    bool is_unsized() const { return n == 0; }
};

}  // namespace tr
}  // namespace hj
