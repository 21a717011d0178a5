// trace.cpp — trace DAG, thread-local schedule, op constructors (see trace.h).
//
// Restates hephaestus-jit/src/trace.rs.  Each function cites the lines it follows.  The
// graph compiler and launcher live in tgraph.cpp.
#include <type_traits>
#include <cstring>

#include "trace_internal.h"

#include <algorithm>
#include <cstring>

namespace hj {
namespace tr {

// =============================================================================================
// types (vartype.rs)
// =============================================================================================
namespace {
std::mutex g_types_mu;
std::vector<TypeNode> g_types;
void ensure_scalars() {
    if (g_types.empty()) {
        for (uint32_t k = HJ_VOID; k <= HJ_F64; k++) {
            TypeNode n;
            n.kind = k;
            g_types.push_back(n);
        }
    }
}
bool same(const TypeNode& a, const TypeNode& b) {
    return a.kind == b.kind && a.elem == b.elem && a.num == b.num && a.cols == b.cols && a.rows == b.rows &&
           a.fields == b.fields;
}
TypeId intern(const TypeNode& n) {
    std::lock_guard<std::mutex> g(g_types_mu);
    ensure_scalars();
    for (size_t i = 0; i < g_types.size(); i++)
        if (same(g_types[i], n)) return (TypeId)i;
    g_types.push_back(n);
    return (TypeId)(g_types.size() - 1);
}
size_t align_up(size_t x, size_t a) { return a ? (x + a - 1) / a * a : x; }
}  // namespace

TypeId type_scalar(uint32_t kind) { return kind <= HJ_F64 ? (TypeId)kind : 0; }
TypeId type_vector(TypeId elem, uint32_t num) {
    TypeNode n; n.kind = HJ_VEC; n.elem = elem; n.num = num; return intern(n);
}
TypeId type_array(TypeId elem, uint32_t num) {
    TypeNode n; n.kind = HJ_ARRAY; n.elem = elem; n.num = num; return intern(n);
}
TypeId type_matrix(TypeId elem, uint32_t cols, uint32_t rows) {
    TypeNode n; n.kind = HJ_MAT; n.elem = elem; n.cols = cols; n.rows = rows; return intern(n);
}
TypeId type_struct(const TypeId* fields, uint32_t nf) {
    TypeNode n; n.kind = HJ_STRUCT; n.num = nf; n.fields.assign(fields, fields + nf); return intern(n);
}
TypeNode type_node(TypeId t) {
    std::lock_guard<std::mutex> g(g_types_mu);
    ensure_scalars();
    return t < g_types.size() ? g_types[t] : TypeNode();
}
size_t type_alignment(TypeId t) {  // vartype.rs:169-189
    TypeNode n = type_node(t);
    switch (n.kind) {
    case HJ_VOID: return 0;
    case HJ_BOOL: case HJ_I8: case HJ_U8: return 1;
    case HJ_I16: case HJ_U16: case HJ_F16: return 2;
    case HJ_I32: case HJ_U32: case HJ_F32: return 4;
    case HJ_I64: case HJ_U64: case HJ_F64: return 8;
    case HJ_VEC: case HJ_ARRAY: return type_alignment(n.elem);
    case HJ_MAT: return type_size(n.elem) * n.rows;
    case HJ_STRUCT: {
        size_t a = 0;
        for (TypeId f : n.fields) a = std::max(a, type_alignment(f));
        return a;
    }
    default: return 0;
    }
}
size_t type_size(TypeId t) {  // vartype.rs:125-155
    TypeNode n = type_node(t);
    switch (n.kind) {
    case HJ_VOID: return 0;
    case HJ_BOOL: case HJ_I8: case HJ_U8: return 1;
    case HJ_I16: case HJ_U16: case HJ_F16: return 2;
    case HJ_I32: case HJ_U32: case HJ_F32: return 4;
    case HJ_I64: case HJ_U64: case HJ_F64: return 8;
    case HJ_VEC: case HJ_ARRAY: return type_size(n.elem) * n.num;
    case HJ_MAT: return type_size(n.elem) * n.cols * n.rows;
    case HJ_STRUCT: {
        if (n.fields.empty()) return 0;
        size_t off = 0;
        for (size_t i = 0; i + 1 < n.fields.size(); i++) {
            off += type_size(n.fields[i]);
            off = align_up(off, type_alignment(n.fields[i + 1]));
        }
        return align_up(off + type_size(n.fields.back()), type_alignment(t));
    }
    default: return 0;
    }
}
int type_num_elements(TypeId t) {
    TypeNode n = type_node(t);
    switch (n.kind) {
    case HJ_VEC: case HJ_ARRAY: return (int)n.num;
    case HJ_MAT: return (int)(n.rows * n.cols);
    case HJ_STRUCT: return (int)n.fields.size();
    default: return -1;
    }
}
std::string type_debug(TypeId t) {  // #[derive(Debug)] of VarType, non-alternate form
    static const char* names[] = {"Void", "Bool", "I8", "U8", "I16", "U16", "I32", "U32", "I64", "U64", "F16", "F32", "F64"};
    TypeNode n = type_node(t);
    std::string s;
    switch (n.kind) {
    case HJ_VEC: return "Vec { ty: " + type_debug(n.elem) + ", num: " + std::to_string(n.num) + " }";
    case HJ_ARRAY: return "Array { ty: " + type_debug(n.elem) + ", num: " + std::to_string(n.num) + " }";
    case HJ_MAT:
        return "Mat { ty: " + type_debug(n.elem) + ", rows: " + std::to_string(n.rows) + ", cols: " + std::to_string(n.cols) + " }";
    case HJ_STRUCT:
        s = "Struct { tys: [";
        for (size_t i = 0; i < n.fields.size(); i++) s += (i ? ", " : "") + type_debug(n.fields[i]);
        return s + "] }";
    default: return names[n.kind <= HJ_F64 ? n.kind : 0];
    }
}

// =============================================================================================
// trace (trace.rs:81-227)
// =============================================================================================
std::mutex g_trace_mu;  // the reference's `static TRACE: Lazy<Mutex<Trace>>` (trace.rs:71)
Trace g_trace;
thread_local ThreadState t_ts;

Var* Trace::get(VarId id) {
    uint32_t idx = (uint32_t)id, gen = (uint32_t)(id >> 32);
    if (id == NO_VAR || idx >= slots.size() || slots[idx].gen != gen || !slots[idx].live) return nullptr;
    return &slots[idx].var;
}
Var& Trace::var(VarId id) {
    Var* v = get(id);
    if (!v) throw TraceError("use of a variable that is no longer part of the trace");
    return *v;
}
VarId Trace::new_var_id(Var v) {  // trace.rs:104-111
    for (VarId d : v.deps) inc_rc(d);
    if (v.extent.dynamic) inc_rc(v.extent.size_var);
    v.rc = 1;
    uint32_t idx;
    if (!free_list.empty()) {
        idx = free_list.back();
        free_list.pop_back();
    } else {
        idx = (uint32_t)slots.size();
        slots.emplace_back();
        slots[idx].gen = 0;
    }
    Slot& s = slots[idx];
    s.gen += 1;  // generations start at 1, so a VarId is never 0
    s.live = true;
    s.var = std::move(v);
    n_live++;
    return ((VarId)s.gen << 32) | idx;
}
void Trace::inc_rc(VarId id) { var(id).rc++; }
void Trace::dec_rc(VarId first) {  // trace.rs:147-160; a work list instead of the reference's recursion:
    std::vector<VarId> work{first};  // dropping the end of a 300 000-operation chain must not overflow the stack
    while (!work.empty()) {
        const VarId id = work.back();
        work.pop_back();
        Var& v = var(id);
        if (--v.rc != 0) continue;
        work.insert(work.end(), v.deps.begin(), v.deps.end());
        if (v.extent.dynamic) work.push_back(v.extent.size_var);
        if (v.data.kind == Resource::Buffer && v.data.buf) hj_buffer_release(v.data.buf);
        if (v.data.seed) hj_buffer_release(v.data.seed);
        const uint32_t idx = (uint32_t)id;
        slots[idx].live = false;
        slots[idx].var = Var();
        free_list.push_back(idx);
        n_live--;
    }
}
void Trace::advance(VarId id) {  // trace.rs:129-140
    Var& v = var(id);
    v.op = resulting_op(v.op);
    std::vector<VarId> deps;
    deps.swap(v.deps);
    for (VarId d : deps) dec_rc(d);
}

Op resulting_op(const Op& op) {  // op.rs:167-177, DeviceOp::resulting_op op.rs:131-142
    Op r;
    switch (op.kind) {
    case OpKind::Buffer: r.kind = OpKind::Buffer; return r;
    case OpKind::KernelOp: r.kind = OpKind::Buffer; return r;
    case OpKind::DeviceOp:
        r.kind = op.code == DOP_COMPRESS ? OpKind::Nop : OpKind::Buffer;
        return r;
    default: throw TraceError("resulting_op of a Nop / Ref variable (todo!() in the reference, op.rs:175)");
    }
}

void set_resource(Var& v, const Resource& r) {
    if (r.kind == Resource::Buffer && r.buf) hj_buffer_retain(r.buf);
    if (r.seed) hj_buffer_retain(r.seed);
    if (v.data.kind == Resource::Buffer && v.data.buf) hj_buffer_release(v.data.buf);
    if (v.data.seed) hj_buffer_release(v.data.seed);
    v.data = r;
}

// ---- VarRef-like helpers (all take the trace lock themselves) ---------------------------------
namespace {
struct Lock {
    std::lock_guard<std::mutex> g;
    Lock() : g(g_trace_mu) {}
};
}  // namespace

VarId ref_clone(VarId id) {
    Lock l;
    g_trace.inc_rc(id);
    return id;
}
void ref_drop(VarId id) {
    Lock l;
    g_trace.dec_rc(id);
}
Extent extent_of(VarId id) { Lock l; return g_trace.var(id).extent; }
TypeId type_of(VarId id) { Lock l; return g_trace.var(id).ty; }
Op op_of(VarId id) { Lock l; return g_trace.var(id).op; }
static bool is_evaluated(VarId id) { return op_of(id).kind == OpKind::Buffer; }

Extent resulting_extent(const Extent& a, const Extent& b) {  // extent.rs:81-110
    if (!a.dynamic && !b.dynamic) {
        Extent e; e.n = std::max(a.n, b.n); return e;
    }
    if (a.dynamic != b.dynamic) {
        const Extent& d = a.dynamic ? a : b;
        const Extent& s = a.dynamic ? b : a;
        Extent e; e.dynamic = true; e.n = std::max(s.n, d.n); e.size_var = d.size_var; return e;
    }
    if (a.n != b.n || a.size_var != b.size_var)
        throw TraceError("operands have different dynamic extents (assert_eq! in extent.rs:96-97)");
    return a;
}
static Extent resulting_extent(std::initializer_list<VarId> refs) {  // trace.rs:667-671
    Extent e;
    for (VarId r : refs) e = resulting_extent(e, extent_of(r));
    return e;
}

// ThreadState::new_group (trace.rs:46-57)
void ThreadState::new_group() {
    if (!groups.empty()) {
        Lock l;
        for (size_t i = groups.back().first; i < groups.back().second; i++)
            if (Var* v = g_trace.get(scheduled[i])) v->dirty = false;
    }
    size_t end = scheduled.size();
    if (start != end) {
        groups.emplace_back(start, end);
        start = end;
    }
}
void ThreadState::clear() {
    for (VarId id : scheduled) ref_drop(id);
    for (VarId id : recorded_se) ref_drop(id);
    scheduled.clear();
    scheduled_set.clear();
    groups.clear();
    start = 0;
    recorded_se.clear();
    recorded_se_start.clear();
}

void schedule_eval() { t_ts.new_group(); }  // trace.rs:540-545

void schedule(VarId id) {  // trace.rs:919-935
    Extent e = extent_of(id);
    if (is_evaluated(id) || e.is_unsized()) return;
    if (!t_ts.recorded_se_start.empty() && type_of(id) == type_scalar(HJ_VOID)) {
        t_ts.recorded_se.push_back(ref_clone(id));
    } else if (t_ts.scheduled_set.insert(id).second) {
        t_ts.scheduled.push_back(ref_clone(id));
    }
}

// new_var (trace.rs:364-399): the kernel-boundary rules
VarId new_var(Var v, const std::vector<VarId>& deps) {
    const bool is_device_op = v.op.kind == OpKind::DeviceOp;
    if (is_device_op)
        for (VarId d : deps) schedule(d);
    if (is_device_op) schedule_eval();
    bool any_dirty = false;
    {
        Lock l;
        for (VarId d : deps) any_dirty |= g_trace.var(d).dirty;
    }
    if (any_dirty) schedule_eval();
    v.deps = deps;
    VarId res;
    {
        Lock l;
        res = g_trace.new_var_id(std::move(v));
    }
    if (is_device_op) {
        schedule(res);
        schedule_eval();
    }
    return res;
}

static Var make(OpKind kind, uint32_t code, uint32_t arg, TypeId ty, const Extent& e) {
    Var v;
    v.op.kind = kind;
    v.op.code = code;
    v.op.arg = arg;
    v.ty = ty;
    v.extent = e;
    return v;
}
static Var kvar(uint32_t kop, uint32_t arg, TypeId ty, const Extent& e) { return make(OpKind::KernelOp, kop, arg, ty, e); }

// ---- constructors (trace.rs:547-663) -----------------------------------------------------------
VarId index() { return new_var(kvar(HJ_OP_INDEX, 0, type_scalar(HJ_U32), Extent()), {}); }
VarId sized_index(size_t n) {
    Extent e; e.n = n;
    return new_var(kvar(HJ_OP_INDEX, 0, type_scalar(HJ_U32), e), {});
}
VarId dynamic_index(size_t capacity, VarId size) {  // trace.rs:578-597
    schedule(size);
    schedule_eval();
    Extent e; e.dynamic = true; e.n = capacity; e.size_var = size;
    return new_var(kvar(HJ_OP_INDEX, 0, type_scalar(HJ_U32), e), {});
}
VarId literal(TypeId ty, uint64_t bits, size_t size) {  // trace.rs:602-641
    Extent e; e.n = size;
    Var v = kvar(HJ_OP_LITERAL, 0, ty, e);
    v.data.kind = Resource::Literal;
    v.data.lit = bits;
    return new_var(std::move(v), {});
}
VarId array(hj_device* dev, TypeId ty, const void* data, size_t n) {  // trace.rs:647-663
    hj_buffer* buf = nullptr;
    hj_status s = hj_buffer_create_from_slice(dev, data, n * type_size(ty), &buf);
    if (s != HJ_OK) throw TraceError(std::string("create_buffer_from_slice failed: ") + hj_last_error());
    Extent e; e.n = n;
    Var v = make(OpKind::Buffer, 0, 0, ty, e);
    v.data.kind = Resource::Buffer;
    v.data.buf = buf;  // ownership of the creation reference moves into the trace
    return new_var(std::move(v), {});
}
VarId array_async(hj_device* dev, TypeId ty, const void* pinned, size_t n) {  // runtime.cpp: chunk-wise upload on a side stream
    hj_buffer* buf = nullptr;
    if (hj_buffer_create_from_host_async(dev, pinned, n * type_size(ty), type_size(ty), &buf) != HJ_OK)
        throw TraceError(std::string("create_buffer_from_host_async failed: ") + hj_last_error());
    Extent e; e.n = n;
    Var v = make(OpKind::Buffer, 0, 0, ty, e);
    v.data.kind = Resource::Buffer;
    v.data.buf = buf;
    return new_var(std::move(v), {});
}
VarId from_buffer(hj_buffer* buf, TypeId ty, size_t n) {
    hj_buffer_retain(buf);
    Extent e; e.n = n;
    Var v = make(OpKind::Buffer, 0, 0, ty, e);
    v.data.kind = Resource::Buffer;
    v.data.buf = buf;
    return new_var(std::move(v), {});
}

// ---- sharded arrays (no reference counterpart: SURVEY 8e) ------------------------------------------
namespace {
void local_block(hj_comm* comm, size_t n_global, uint64_t* start, uint64_t* count) {
    int32_t rank = 0, world = 1;
    if (hj_comm_info(comm, &rank, &world, nullptr, nullptr) != HJ_OK) throw TraceError(std::string("hj_comm_info failed: ") + hj_last_error());
    uint64_t s0 = 0, s1 = 0;
    hj_shard_bounds(n_global, world, rank, &s0, &s1);
    *start = s0;
    *count = s1 - s0;
}
hj_device* comm_dev(hj_comm* comm) {
    hj_device* dev = nullptr;
    if (hj_comm_device(comm, &dev) != HJ_OK || !dev) throw TraceError(std::string("hj_comm_device failed: ") + hj_last_error());
    return dev;
}
}  // namespace
VarId array_sharded(hj_comm* comm, TypeId ty, const void* local_data, size_t n_global) {
    if (!comm) throw TraceError("array_sharded: null communicator");
    uint64_t s0, cnt;
    local_block(comm, n_global, &s0, &cnt);
    if (cnt == 0) throw TraceError("array_sharded: fewer elements than ranks");
    hj_device* dev = comm_dev(comm);  // the buffer must live where the sharded kernels run
    hj_buffer* buf = nullptr;
    if (hj_buffer_create_from_slice(dev, local_data, cnt * type_size(ty), &buf) != HJ_OK)
        throw TraceError(std::string("create_buffer_from_slice failed: ") + hj_last_error());
    Extent e; e.n = n_global;
    Var v = make(OpKind::Buffer, 0, 0, ty, e);
    v.data.kind = Resource::Buffer;
    v.data.buf = buf;
    v.data.comm = comm;
    return new_var(std::move(v), {});
}
VarId from_buffer_sharded(hj_comm* comm, hj_buffer* local_buf, TypeId ty, size_t n_global) {
    if (!comm || !local_buf) throw TraceError("from_buffer_sharded: null argument");
    uint64_t s0, cnt;
    local_block(comm, n_global, &s0, &cnt);
    size_t bytes = 0;
    hj_buffer_size(local_buf, &bytes);
    if (cnt == 0 || bytes < cnt * type_size(ty)) throw TraceError("from_buffer_sharded: the buffer does not hold this rank's block");
    hj_buffer_retain(local_buf);
    Extent e; e.n = n_global;
    Var v = make(OpKind::Buffer, 0, 0, ty, e);
    v.data.kind = Resource::Buffer;
    v.data.buf = local_buf;
    v.data.comm = comm;
    return new_var(std::move(v), {});
}
ShardInfo shard_info(VarId id) {
    hj_comm* comm = nullptr;
    hj_buffer* seed = nullptr;
    ShardInfo si;
    size_t n = 0;
    {
        Lock l;
        const Var& v = g_trace.var(id);
        comm = v.data.kind == Resource::Buffer ? v.data.comm : nullptr;
        si.deferred = v.data.deferred;
        si.segment = v.data.segment;
        si.segment_local = v.data.segment_local;
        n = v.extent.n;
        if (comm && si.segment && v.extent.dynamic) {  // a static `compress` index keeps its capacity and zero tail
            seed = v.data.seed;
            if (!seed) throw TraceError("a sharded segment without its count buffer");
            hj_buffer_retain(seed);
        }
    }
    if (!comm) return ShardInfo();
    si.sharded = true;
    local_block(comm, n, &si.start, &si.count);
    if (seed) {
        uint32_t valid = 0;
        const hj_status s = hj_buffer_to_host(seed, 0, 4, &valid);
        hj_buffer_release(seed);
        if (s != HJ_OK) throw TraceError(std::string("to_host failed: ") + hj_last_error());
        si.count = std::min<uint64_t>(si.count, valid);
    }
    return si;
}
void materialise(VarId id) {
    hj_buffer *buf = nullptr, *seed = nullptr;
    hj_comm* comm = nullptr;
    TypeId ty;
    size_t n;
    {
        Lock l;
        const Var& v = g_trace.var(id);
        if (v.data.kind != Resource::Buffer || !v.data.comm || !v.data.deferred) return;
        buf = v.data.buf; seed = v.data.seed; comm = v.data.comm; ty = v.ty; n = v.extent.n;
        hj_buffer_retain(buf);
        hj_buffer_retain(seed);
    }
    uint64_t s0, cnt;
    local_block(comm, n, &s0, &cnt);
    hj_status s = hj_apply_seed(comm_dev(comm), (hj_type_kind)type_node(ty).kind, cnt, buf, seed);
    hj_buffer_release(buf);
    hj_buffer_release(seed);
    if (s != HJ_OK) throw TraceError(std::string("hj_apply_seed failed: ") + hj_last_error());
    Lock l;
    if (Var* v = g_trace.get(id)) {
        if (v->data.buf == buf) {
            v->data.deferred = false;
            if (v->data.seed) hj_buffer_release(v->data.seed);
            v->data.seed = nullptr;
        }
    }
}

// ---- elementwise ops (trace.rs:968-1082, 1335-1351, 1483-1504) ------------------------------------
VarId bop(uint32_t op, VarId a, VarId b) {
    TypeId ty = type_of(a);
    if (type_of(b) != ty) throw TraceError("binary op on operands of different types (assert_eq!, trace.rs:977)");
    if (op >= HJ_BOP_EQ) ty = type_scalar(HJ_BOOL);
    return new_var(kvar(HJ_OP_BOP, op, ty, resulting_extent({a, b})), {a, b});
}
VarId uop(uint32_t op, VarId a) { return new_var(kvar(HJ_OP_UOP, op, type_of(a), resulting_extent({a})), {a}); }
VarId cast(VarId a, TypeId ty) { return new_var(kvar(HJ_OP_UOP, HJ_UOP_CAST, ty, resulting_extent({a})), {a}); }
VarId bitcast(VarId a, TypeId ty) { return new_var(kvar(HJ_OP_UOP, HJ_UOP_BITCAST, ty, resulting_extent({a})), {a}); }
VarId fma(VarId a, VarId b, VarId c) {
    return new_var(kvar(HJ_OP_FMA, 0, type_of(a), resulting_extent({a, b, c})), {a, b, c});
}
VarId select(VarId true_val, VarId cond, VarId false_val) {
    if (type_of(cond) != type_scalar(HJ_BOOL)) throw TraceError("select: condition is not Bool (trace.rs:1488)");
    if (type_of(true_val) != type_of(false_val)) throw TraceError("select: value types differ (trace.rs:1489)");
    return new_var(kvar(HJ_OP_SELECT, 0, type_of(true_val), resulting_extent({cond, true_val, false_val})),
                   {cond, true_val, false_val});
}
VarId extract(VarId a, uint32_t elem) {  // trace.rs:1439-1459
    TypeNode n = type_node(type_of(a));
    TypeId ty;
    if (n.kind == HJ_VEC || n.kind == HJ_ARRAY) ty = n.elem;
    else if (n.kind == HJ_STRUCT && elem < n.fields.size()) ty = n.fields[elem];
    else throw TraceError("extract: not a Vec / Array / Struct, or element out of range");
    return new_var(kvar(HJ_OP_EXTRACT, elem, ty, extent_of(a)), {a});
}
VarId extract_dyn(VarId a, VarId elem) {
    TypeNode n = type_node(type_of(a));
    if (n.kind != HJ_ARRAY) throw TraceError("extract_dyn: not an Array (todo!() in the reference)");
    return new_var(kvar(HJ_OP_DYN_EXTRACT, 0, n.elem, extent_of(a)), {a, elem});
}
static VarId construct(TypeId ty, const std::vector<VarId>& refs) {
    Extent e;
    for (VarId r : refs) e = resulting_extent(e, extent_of(r));
    return new_var(kvar(HJ_OP_CONSTRUCT, 0, ty, e), refs);
}
VarId composite(const std::vector<VarId>& refs) {  // trace.rs:675-693
    std::vector<TypeId> tys;
    for (VarId r : refs) tys.push_back(type_of(r));
    return construct(type_struct(tys.data(), (uint32_t)tys.size()), refs);
}
VarId vec(const std::vector<VarId>& refs) { return construct(type_vector(type_of(refs.at(0)), (uint32_t)refs.size()), refs); }
VarId arr(const std::vector<VarId>& refs) { return construct(type_array(type_of(refs.at(0)), (uint32_t)refs.size()), refs); }
// tr::mat (trace.rs:734-756): columns are vectors of `rows` elements, storage is column-major
VarId mat(const std::vector<VarId>& columns) {
    const TypeNode col = type_node(type_of(columns.at(0)));
    if (col.kind != HJ_VEC) throw TraceError("mat: the columns must be vectors");
    return construct(type_matrix(col.elem, (uint32_t)columns.size(), col.num), columns);
}

// ---- references, gather, scatter (trace.rs:1084-1334, 1354-1375) -----------------------------------
static VarId get_ref(VarId a, bool mutable_) {
    Extent e;  // Size(0)
    Var v = make(OpKind::Ref, 0, 0, type_of(a), e);
    v.op.ref_mutable = mutable_;
    return new_var(std::move(v), {a});
}
static void mark_dirty(VarId id) { Lock l; g_trace.var(id).dirty = true; }

// VarRef::reindex (trace.rs:1084-1121): re-trace a pure expression of Index at a new index
static bool reindex(VarId self, VarId new_idx, VarId* out, int depth = 0) {
    // an expression too deep to re-trace recursively is evaluated and gathered from instead
    if (depth > 4096) return false;
    Var snapshot;
    {
        Lock l;
        const Var& v = g_trace.var(self);
        if (v.data.kind != Resource::None) return false;   // is_data()
        snapshot.op = v.op;
        snapshot.ty = v.ty;
        snapshot.deps = v.deps;
    }
    if (snapshot.op.kind == OpKind::KernelOp && snapshot.op.code == HJ_OP_BUFFER_REF) {  // is_ref() (trace.rs:951-953)
        *out = ref_clone(self);
        return true;
    }
    // A reference made by get_ref (Op::Ref) is not is_ref() in the reference, which therefore re-indexes
    // the variable BEHIND it: that fails for evaluated variables (-> the caller evaluates and gathers, the
    // only outcome that works there), but for a scheduled, not yet launched pure expression it
    // "succeeds" and leaves a reference to a variable nobody scheduled — a kernel that cannot be
    // compiled (found by random-program fuzzing: `b = a.gather(i); b.gather(j)` with a = index + 1).
    // Stopping at the reference gives the working outcome in both cases.
    if (snapshot.op.kind == OpKind::Ref) return false;
    // The reference re-creates a device op (reduce / scan / compress) met on the way with the extent
    // of the new index (trace.rs:1110-1118) and then panics in the compiler (`todo!()`,
    // compiler.rs:131) — `x.reduce_sum().gather(0)` over a pure index expression cannot be traced
    // there.  A device op is not an elementwise expression: it is evaluated and gathered from, like
    // any other buffer.
    if (snapshot.op.kind == OpKind::DeviceOp) return false;
    std::vector<VarId> deps;
    for (VarId d : snapshot.deps) {
        VarId nd;
        if (!reindex(d, new_idx, &nd, depth + 1)) {
            for (VarId x : deps) ref_drop(x);
            return false;
        }
        deps.push_back(nd);
    }
    if (snapshot.op.kind == OpKind::KernelOp && snapshot.op.code == HJ_OP_INDEX) {
        *out = ref_clone(new_idx);
    } else {
        Var v;
        v.op = snapshot.op;
        v.ty = snapshot.ty;
        v.extent = extent_of(new_idx);
        *out = new_var(std::move(v), deps);
    }
    for (VarId x : deps) ref_drop(x);
    return true;
}

VarId gather_if(VarId self, VarId idx, VarId active) {  // trace.rs:1125-1164
    if (extent_of(self).is_unsized()) {  // resize a literal
        Lock l;
        const Var& src = g_trace.var(self);
        Var v;
        v.op = src.op;
        v.ty = src.ty;
        v.extent = g_trace.var(idx).extent;
        v.data = src.data;  // Literal (buffers are never unsized)
        return g_trace.new_var_id(std::move(v));
    }
    VarId re;
    if (reindex(self, idx, &re)) return re;
    schedule(self);
    schedule_eval();
    VarId src_ref = get_ref(self, false);
    VarId res = new_var(kvar(HJ_OP_GATHER, 0, type_of(self), extent_of(idx)), {src_ref, idx, active});
    ref_drop(src_ref);
    return res;
}

// scatter / scatter_if / scatter_reduce(_if) / scatter_atomic(_if) share one shape
// (trace.rs:1167-1309); `active` may be NO_VAR.  Returns NO_VAR for the self-scheduling void ops.
VarId scatter_like(uint32_t kop, uint32_t rop, VarId self, VarId dst, VarId idx, VarId active) {
    schedule(dst);
    schedule_eval();
    Extent e = active != NO_VAR ? resulting_extent({self, idx, active}) : resulting_extent({self, idx});
    VarId dst_ref = get_ref(dst, true);
    const bool returns = kop == HJ_OP_SCATTER_ATOMIC;
    TypeId ty = returns ? type_of(self) : type_scalar(HJ_VOID);
    std::vector<VarId> deps = {dst_ref, self, idx};
    if (active != NO_VAR) deps.push_back(active);
    VarId res = new_var(kvar(kop, rop, ty, e), deps);
    ref_drop(dst_ref);
    mark_dirty(dst);
    if (returns) return res;  // NOTE: do not schedule result of scatter_atomic (trace.rs:1277)
    schedule(res);            // auto schedule
    ref_drop(res);
    return NO_VAR;
}
VarId atomic_inc(VarId self, VarId idx, VarId active) {  // trace.rs:1310-1334
    schedule(self);
    schedule_eval();
    Extent ie = extent_of(idx);
    if (ie.dynamic || ie.n > 1) throw TraceError("atomic_inc: the index must have extent <= 1 (trace.rs:1320)");
    VarId dst_ref = get_ref(self, true);
    VarId res = new_var(kvar(HJ_OP_ATOMIC_INC, 0, type_of(self), resulting_extent({active})), {dst_ref, idx, active});
    ref_drop(dst_ref);
    mark_dirty(self);
    return res;
}

// ---- recorded loops / ifs (trace.rs:408-521) ---------------------------------------------------
static std::vector<VarId> extract_all(VarId s) {
    int n = type_num_elements(type_of(s));
    std::vector<VarId> out;
    for (int i = 0; i < n; i++) out.push_back(extract(s, (uint32_t)i));
    return out;
}
VarId scope_start(bool is_loop, const std::vector<VarId>& state_vars, std::vector<VarId>* state_out) {
    VarId state = composite(state_vars);
    t_ts.recorded_se_start.push_back(t_ts.recorded_se.size());
    VarId start = new_var(kvar(is_loop ? HJ_OP_LOOP_START : HJ_OP_IF_START, 0, type_of(state), extent_of(state)), {state});
    ref_drop(state);
    *state_out = extract_all(start);
    return start;
}
void scope_end(VarId start, const std::vector<VarId>& state_vars, std::vector<VarId>* state_out) {
    VarId state = composite(state_vars);
    if (t_ts.recorded_se_start.empty()) throw TraceError("loop_end / if_end without a matching start");
    size_t first = t_ts.recorded_se_start.back();
    t_ts.recorded_se_start.pop_back();
    std::vector<VarId> side_effects(t_ts.recorded_se.begin() + first, t_ts.recorded_se.end());
    t_ts.recorded_se.resize(first);
    std::vector<VarId> deps = {start, state};
    deps.insert(deps.end(), side_effects.begin(), side_effects.end());
    // NOTE: the reference emits KernelOp::LoopEnd for if_end as well (trace.rs:510); codegen treats
    // both end ops identically.
    VarId end = new_var(kvar(HJ_OP_LOOP_END, 0, type_of(state), extent_of(state)), deps);
    ref_drop(state);
    for (VarId s : side_effects) ref_drop(s);
    *state_out = extract_all(end);
    ref_drop(end);
}

// ---- device ops (trace.rs:1583-1674) ----------------------------------------------------------------
static Var dvar(uint32_t dop, uint32_t arg, TypeId ty, const Extent& e) { return make(OpKind::DeviceOp, dop, arg, ty, e); }

void compress(VarId mask, VarId* count, VarId* index) {  // trace.rs:1595-1620
    if (type_of(mask) != type_scalar(HJ_BOOL)) throw TraceError("compress: the mask is not Bool (trace.rs:1596)");
    Extent e = extent_of(mask);
    *count = literal(type_scalar(HJ_U32), 0, 1);
    *index = literal(type_scalar(HJ_U32), 0, e.n);
    schedule(*count);
    schedule(*index);
    schedule(mask);
    schedule_eval();
    VarId res = new_var(dvar(DOP_COMPRESS, 0, type_scalar(HJ_VOID), e), {*index, *count, mask});
    ref_drop(res);  // kept alive by the schedule (auto-scheduled by new_var)
}
VarId compress_dyn(VarId mask) {  // trace.rs:1583-1591
    size_t capacity = extent_of(mask).n;
    VarId count, idx;
    compress(mask, &count, &idx);
    VarId dyn = dynamic_index(capacity, count);
    VarId t = literal(type_scalar(HJ_BOOL), 1, 0);
    VarId out = gather_if(idx, dyn, t);
    ref_drop(t);
    ref_drop(dyn);
    ref_drop(count);
    ref_drop(idx);
    return out;
}
VarId prefix_sum(VarId a, bool inclusive) {  // trace.rs:1623-1637
    Extent e = extent_of(a);
    if (e.dynamic) throw TraceError("prefix_sum of a dynamically sized variable (todo!() in the reference, trace.rs:1380)");
    return new_var(dvar(DOP_PREFIX_SUM, inclusive ? 1 : 0, type_of(a), e), {a});
}
VarId reduce(VarId a, uint32_t op) {  // trace.rs:1641-1653
    Extent e; e.n = 1;
    return new_var(dvar(DOP_REDUCE, op, type_of(a), e), {a});
}

// ---- hashing for the function cache (trace.rs:250-265) ----------------------------------------
uint64_t var_hash(VarId id) {
    Lock l;
    const Var& v = g_trace.var(id);
    uint64_t f[8] = {(uint64_t)v.op.kind, v.op.ref_mutable, v.op.code, v.op.arg, v.ty, v.extent.dynamic, v.extent.n,
                     v.extent.size_var};
    uint64_t h = hash_bytes(f, sizeof(f));
    if (v.op.kind == OpKind::KernelOp && v.op.code == HJ_OP_LITERAL && v.data.kind == Resource::Literal)
        h = hash_bytes(&v.data.lit, 8, h);
    return h;
}

// ---- to_vec (trace.rs:1404-1438) ------------------------------------------------------------------
size_t current_size(VarId id) {
    Extent e = extent_of(id);
    if (!e.dynamic) return e.n;
    hj_buffer* b = nullptr;
    {
        Lock l;
        const Var& sv = g_trace.var(e.size_var);
        if (sv.data.kind != Resource::Buffer) throw TraceError("the size variable of a DynSize extent is not evaluated");
        b = sv.data.buf;
    }
    int32_t n = 0;
    if (hj_buffer_to_host(b, 0, 4, &n) != HJ_OK) throw TraceError(std::string("to_host failed: ") + hj_last_error());
    return (size_t)n;
}
// For a sharded variable `start_elem` / `n_elem` address this rank's BLOCK (shard_info); a deferred
// scan result is completed on the host: the block is downloaded as it is and the rank's offset added.
void to_host(VarId id, size_t start_elem, size_t n_elem, void* dst) {
    hj_buffer *b = nullptr, *seed = nullptr;
    size_t es;
    uint32_t kind;
    {
        Lock l;
        const Var& v = g_trace.var(id);
        if (v.data.kind != Resource::Buffer) throw TraceError("to_vec of a variable that has not been evaluated");
        b = v.data.buf;
        es = type_size(v.ty);
        kind = type_node(v.ty).kind;
        if (v.data.deferred) seed = v.data.seed;
    }
    if (n_elem == 0) return;
    if (hj_buffer_to_host(b, start_elem * es, n_elem * es, dst) != HJ_OK)
        throw TraceError(std::string("to_host failed: ") + hj_last_error());
    if (!seed) return;
    unsigned long long raw = 0;
    if (hj_buffer_to_host(seed, 0, es, &raw) != HJ_OK) throw TraceError(std::string("to_host failed: ") + hj_last_error());
    auto add = [&](auto* p) {
        using T = std::remove_pointer_t<decltype(p)>;
        T s;
        memcpy(&s, &raw, sizeof(T));
        for (size_t i = 0; i < n_elem; i++) p[i] = (T)(p[i] + s);
    };
    switch (kind) {
    case HJ_I8: case HJ_U8: add((uint8_t*)dst); break;
    case HJ_I16: case HJ_U16: add((uint16_t*)dst); break;
    case HJ_I32: case HJ_U32: add((uint32_t*)dst); break;
    case HJ_I64: case HJ_U64: add((uint64_t*)dst); break;
    case HJ_F32: add((float*)dst); break;
    case HJ_F64: add((double*)dst); break;
    default: throw TraceError("to_vec: a deferred seed on a non-scalar variable");
    }
}

}  // namespace tr
}  // namespace hj
