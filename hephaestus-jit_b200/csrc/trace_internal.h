// trace_internal.h — data structures shared by trace.cpp, tgraph.cpp and trace_abi.cpp.
#pragma once
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

#include "ir.h"
#include "trace.h"

namespace hj {
namespace tr {

struct TraceError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// resource.rs:29-36 (Texture / Accel out of scope)
struct Resource {
    enum Kind : uint8_t { None, Literal, Buffer } kind = None;
    uint64_t lit = 0;
    hj_buffer* buf = nullptr;  // the Var owns one reference
    // Sharded arrays (multi-GPU, one process per GPU): `buf` holds this rank's contiguous block of
    // the variable's global extent (hj_shard_bounds).  `deferred`: the block is a LOCAL scan whose
    // global value is buf[i] + seed[0] (hj_sharded_prefix_sum_deferred); the Var owns a reference to
    // `seed`.  The communicator is borrowed: it must outlive the variables sharded over it.
    // `segment`: the block is a per-rank compacted SEGMENT (the index output of a sharded Compress, or what
    // a DynSize kernel wrote over it): only the first seed[0] (u32) entries are defined.
    hj_comm* comm = nullptr;
    bool deferred = false;
    bool segment = false;
    bool segment_local = false;  // ... whose entries are positions in the rank's part of the parent sequence (HJ_SHARD_SEGMENT_LOCAL)
    hj_buffer* seed = nullptr;
};

// trace.rs:301-315
struct Var {
    Op op;
    TypeId ty = 0;  // Void
    Extent extent;
    bool dirty = false;
    Resource data;
    std::vector<VarId> deps;
    size_t rc = 0;
};

// trace.rs:81-227: slot map of ref-counted variables
struct Trace {
    struct Slot {
        uint32_t gen = 0;
        bool live = false;
        Var var;
    };
    std::vector<Slot> slots;
    std::vector<uint32_t> free_list;
    size_t n_live = 0;

    Var* get(VarId id);   // nullptr if the variable no longer exists
    Var& var(VarId id);   // throws
    VarId new_var_id(Var v);
    void inc_rc(VarId id);
    void dec_rc(VarId id);
    void advance(VarId id);
};

// trace.rs:31-69
struct ThreadState {
    std::vector<VarId> scheduled;              // insertion-ordered (IndexMap), each entry owns a reference
    std::unordered_set<VarId> scheduled_set;
    std::vector<std::pair<size_t, size_t>> groups;
    size_t start = 0;
    std::vector<size_t> recorded_se_start;
    std::vector<VarId> recorded_se;            // each entry owns a reference
    void new_group();
    void clear();  // drops every reference held
};

extern std::mutex g_trace_mu;
extern Trace g_trace;
extern thread_local ThreadState t_ts;

Op resulting_op(const Op& op);
void set_resource(Var& v, const Resource& r);  // retains the new buffer, releases the old one (trace lock held)

// reference counting of user-held references (VarRef clone / drop)
VarId ref_clone(VarId id);
void ref_drop(VarId id);
Extent extent_of(VarId id);
TypeId type_of(VarId id);
Op op_of(VarId id);
Extent resulting_extent(const Extent& a, const Extent& b);

void schedule(VarId id);
void schedule_eval();
VarId new_var(Var v, const std::vector<VarId>& deps);

VarId index();
VarId sized_index(size_t n);
VarId dynamic_index(size_t capacity, VarId size);
VarId literal(TypeId ty, uint64_t bits, size_t size);
VarId array(hj_device* dev, TypeId ty, const void* data, size_t n);
VarId array_async(hj_device* dev, TypeId ty, const void* pinned, size_t n);
VarId from_buffer(hj_buffer* buf, TypeId ty, size_t n);
// this rank's block of an n_global-element array: from host memory / from an existing device buffer
VarId array_sharded(hj_comm* comm, TypeId ty, const void* local_data, size_t n_global);
VarId from_buffer_sharded(hj_comm* comm, hj_buffer* local_buf, TypeId ty, size_t n_global);
// `count`: elements of the rank's block — for a DynSize segment the rank's own count (read from the device)
struct ShardInfo { bool sharded = false, deferred = false, segment = false, segment_local = false; uint64_t start = 0, count = 0; };
ShardInfo shard_info(VarId id);
void materialise(VarId id);  // a deferred scan result becomes an ordinary shard (buf[i] += seed)
VarId bop(uint32_t op, VarId a, VarId b);
VarId uop(uint32_t op, VarId a);
VarId cast(VarId a, TypeId ty);
VarId bitcast(VarId a, TypeId ty);
VarId fma(VarId a, VarId b, VarId c);
VarId select(VarId true_val, VarId cond, VarId false_val);
VarId extract(VarId a, uint32_t elem);
VarId extract_dyn(VarId a, VarId elem);
VarId composite(const std::vector<VarId>& refs);
VarId vec(const std::vector<VarId>& refs);
VarId arr(const std::vector<VarId>& refs);
VarId mat(const std::vector<VarId>& columns);
VarId gather_if(VarId self, VarId idx, VarId active);
VarId scatter_like(uint32_t kop, uint32_t rop, VarId self, VarId dst, VarId idx, VarId active);
VarId atomic_inc(VarId self, VarId idx, VarId active);
VarId scope_start(bool is_loop, const std::vector<VarId>& state_vars, std::vector<VarId>* state_out);
void scope_end(VarId start, const std::vector<VarId>& state_vars, std::vector<VarId>* state_out);
void compress(VarId mask, VarId* count, VarId* index);
VarId compress_dyn(VarId mask);
VarId prefix_sum(VarId a, bool inclusive);
VarId reduce(VarId a, uint32_t op);
uint64_t var_hash(VarId id);
size_t current_size(VarId id);
void to_host(VarId id, size_t start_elem, size_t n_elem, void* dst);

// ---- graph (tgraph.cpp; graph.rs) ---------------------------------------------------------------
struct KernelIR {  // an owned flat IR (ir.rs:40-46) plus the hj_ir view the backend takes
    std::vector<hj_ir_var> vars;
    std::vector<uint32_t> deps;
    std::vector<hj_type_desc> types;
    std::vector<uint32_t> struct_fields;
    uint32_t n_buffers = 0;
    hj_ir view() const;
};
struct Pass {  // graph.rs:409-425
    std::vector<uint32_t> resources;
    int32_t size_buffer = -1;
    bool is_kernel = false;
    KernelIR ir;
    size_t size = 0;
    Op device_op;
};
struct BufferDesc {  // resource.rs / backend BufferDesc
    size_t size = 0;
    TypeId ty = 0;
    bool operator==(const BufferDesc& o) const { return size == o.size && ty == o.ty; }
};
struct GraphResource {  // graph.rs:44-51
    enum Kind : uint8_t { Input, Captured, Internal } kind = Internal;
    VarId id = NO_VAR;  // Captured: owns a reference; Internal: weak
};
struct Graph {
    uint64_t uid = 0;                        // key of its captured CUDA graphs (hj_execute_graph_cached)
    mutable hj_device* launched_on = nullptr;  // device holding captured instances of this graph
    std::vector<Pass> passes;
    std::vector<BufferDesc> resource_descs;
    std::vector<GraphResource> resources;
    std::vector<uint32_t> inputs, outputs;
    // sharded launches: the seed buffer of every sharded integer PrefixSum destination, kept from launch to
    // launch (stable addresses let the captured CUDA graph replay); handed over to the variable — and
    // replaced here — when a launch leaves a live variable or an output in the deferred state
    mutable std::vector<hj_buffer*> seed_cache;
    ~Graph();
};
struct LaunchReport {  // graph.rs:138-143
    float aliasing_rate = 0.f;
    double aliasing_duration_us = 0.0;
    uint32_t n_passes = 0;
    double backend_cpu_us = 0.0;
};

// graph::compile (graph.rs:436-614); consumes `ts`
Graph* compile_graph(ThreadState& ts, const std::vector<VarId>& inputs, const std::vector<VarId>& outputs);
// Graph::launch_with (graph.rs:192-400); `outputs` receives new references
// `comm` != NULL (or any input / captured variable sharded): the graph runs over arrays partitioned
// across the ranks of the communicator (hj_execute_graph_sharded)
void launch_graph(const Graph& g, hj_device* dev, const std::vector<VarId>& inputs, std::vector<VarId>* outputs,
                  LaunchReport* report, hj_report* backend_report, hj_comm* comm = nullptr);
std::string graph_debug_string(const Graph& g);
// wire format of a compiled graph (tgraph_io.cpp); captured buffers travel with their contents
std::vector<uint8_t> serialize_graph(const Graph& g);
Graph* deserialize_graph(hj_device* dev, const void* bytes, size_t n_bytes);

}  // namespace tr
}  // namespace hj
