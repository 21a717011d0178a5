// trace_abi.cpp — C ABI of the trace / schedule / graph layer (include/hj.h, section "trace").
//
// One entry point per user-visible function of hephaestus-jit/src/trace.rs, graph.rs and
// record.rs that is on the hot path; exceptions from the C++ layer (the reference's panics /
// graph::Error values) become status codes + hj_last_error().
#include <memory>
#include <unordered_map>

#include "hj_internal.h"
#include "trace_internal.h"

using namespace hj;
using namespace hj::tr;

struct hj_graph {
    std::unique_ptr<Graph> g;
    std::atomic<int> rc{1};
    std::vector<hj_ir> ir_views;  // hj_graph_pass_ir: views into the passes' owned IR, built on first use
    std::mutex views_mu;
};

namespace {
std::mutex g_fcache_mu;
std::unordered_map<uint64_t, hj_graph*> g_fcache;  // record.rs:116-119 (FCache.graphs)

template <typename F>
hj_status guarded(F&& f) {
    try {
        f();
        return HJ_OK;
    } catch (const TraceError& e) {
        return fail(HJ_ERR_INVALID, "%s", e.what());
    } catch (const std::exception& e) {
        return fail(HJ_ERR_INVALID, "trace: %s", e.what());
    }
}
std::vector<VarId> vec_of(const uint64_t* p, uint32_t n) { return std::vector<VarId>(p, p + n); }
}  // namespace

extern "C" {

// ---- types ---------------------------------------------------------------------------------------
uint32_t hj_tr_type_scalar(uint32_t kind) { return type_scalar(kind); }
uint32_t hj_tr_type_vector(uint32_t elem, uint32_t num) { return type_vector(elem, num); }
uint32_t hj_tr_type_array(uint32_t elem, uint32_t num) { return type_array(elem, num); }
uint32_t hj_tr_type_matrix(uint32_t elem, uint32_t cols, uint32_t rows) { return type_matrix(elem, cols, rows); }
uint32_t hj_tr_type_struct(const uint32_t* fields, uint32_t n) { return type_struct(fields, n); }
size_t hj_tr_type_size(uint32_t ty) { return type_size(ty); }
size_t hj_tr_type_alignment(uint32_t ty) { return type_alignment(ty); }
size_t hj_tr_type_offset(uint32_t ty, uint32_t elem) {  // VarType::offset (vartype.rs:156-168)
    TypeNode n = type_node(ty);
    size_t off = 0;
    for (uint32_t i = 0; i < elem && i + 1 < n.fields.size(); i++) {
        off += type_size(n.fields[i]);
        size_t a = type_alignment(n.fields[i + 1]);
        if (a) off = (off + a - 1) / a * a;
    }
    return off;
}
uint32_t hj_tr_type_kind(uint32_t ty) { return type_node(ty).kind; }

// ---- variables -------------------------------------------------------------------------------------
hj_status hj_tr_var_retain(uint64_t v) { return guarded([&] { ref_clone(v); }); }
hj_status hj_tr_var_release(uint64_t v) { return guarded([&] { ref_drop(v); }); }
hj_status hj_tr_var_info(uint64_t v, uint32_t* ty, int32_t* dynamic, uint64_t* extent, int32_t* evaluated,
                         uint64_t* rc, int32_t* dirty) {
    return guarded([&] {
        std::lock_guard<std::mutex> l(g_trace_mu);
        const Var& var = g_trace.var(v);
        if (ty) *ty = var.ty;
        if (dynamic) *dynamic = var.extent.dynamic;
        if (extent) *extent = var.extent.n;
        if (evaluated) *evaluated = var.op.kind == OpKind::Buffer;
        if (rc) *rc = var.rc;
        if (dirty) *dirty = var.dirty;
    });
}
uint64_t hj_tr_var_hash(uint64_t v) {
    uint64_t h = 0;
    guarded([&] { h = var_hash(v); });
    return h;
}
int32_t hj_tr_is_empty(void) {
    std::lock_guard<std::mutex> l(g_trace_mu);
    return g_trace.n_live == 0;
}
uint64_t hj_tr_n_live(void) {
    std::lock_guard<std::mutex> l(g_trace_mu);
    return g_trace.n_live;
}
hj_status hj_tr_var_buffer(uint64_t v, hj_buffer** out) {  // borrowed, NULL if not evaluated
    return guarded([&] {
        std::lock_guard<std::mutex> l(g_trace_mu);
        const Var& var = g_trace.var(v);
        *out = var.data.kind == Resource::Buffer ? var.data.buf : nullptr;
    });
}

#define OUT(expr) return guarded([&] { *out = (expr); })
hj_status hj_tr_index(uint64_t* out) { OUT(index()); }
hj_status hj_tr_sized_index(uint64_t n, uint64_t* out) { OUT(sized_index(n)); }
hj_status hj_tr_dynamic_index(uint64_t capacity, uint64_t size_var, uint64_t* out) { OUT(dynamic_index(capacity, size_var)); }
hj_status hj_tr_literal(uint32_t ty, uint64_t bits, uint64_t* out) { OUT(literal(ty, bits, 0)); }
hj_status hj_tr_sized_literal(uint32_t ty, uint64_t bits, uint64_t n, uint64_t* out) { OUT(literal(ty, bits, n)); }
hj_status hj_tr_array(hj_device* dev, uint32_t ty, const void* data, uint64_t n, uint64_t* out) { OUT(array(dev, ty, data, n)); }
hj_status hj_tr_from_buffer(hj_buffer* buf, uint32_t ty, uint64_t n, uint64_t* out) { OUT(from_buffer(buf, ty, n)); }
hj_status hj_tr_array_async(hj_device* dev, uint32_t ty, const void* data, uint64_t n, uint64_t* out) { OUT(array_async(dev, ty, data, n)); }
hj_status hj_tr_array_sharded(hj_comm* comm, uint32_t ty, const void* local_data, uint64_t n_global, uint64_t* out) {
    OUT(array_sharded(comm, ty, local_data, n_global));
}
hj_status hj_tr_from_buffer_sharded(hj_comm* comm, hj_buffer* local_buf, uint32_t ty, uint64_t n_global, uint64_t* out) {
    OUT(from_buffer_sharded(comm, local_buf, ty, n_global));
}
hj_status hj_tr_var_shard(uint64_t v, int32_t* sharded, uint64_t* start, uint64_t* count, int32_t* deferred) {
    return guarded([&] {
        ShardInfo si = shard_info(v);
        if (sharded) *sharded = si.sharded;
        if (start) *start = si.start;
        if (count) *count = si.count;
        if (deferred)
            *deferred = si.segment ? (int32_t)(si.segment_local ? HJ_SHARD_SEGMENT_LOCAL : HJ_SHARD_SEGMENT) : si.deferred ? (int32_t)HJ_SHARD_DEFERRED : 0;
    });
}
hj_status hj_tr_materialise(uint64_t v) {
    return guarded([&] { materialise(v); });
}
hj_status hj_tr_bop(uint32_t op, uint64_t a, uint64_t b, uint64_t* out) { OUT(bop(op, a, b)); }
hj_status hj_tr_uop(uint32_t op, uint64_t a, uint64_t* out) { OUT(uop(op, a)); }
hj_status hj_tr_cast(uint64_t a, uint32_t ty, uint64_t* out) { OUT(cast(a, ty)); }
hj_status hj_tr_bitcast(uint64_t a, uint32_t ty, uint64_t* out) { OUT(bitcast(a, ty)); }
hj_status hj_tr_fma(uint64_t a, uint64_t b, uint64_t c, uint64_t* out) { OUT(fma(a, b, c)); }
hj_status hj_tr_select(uint64_t true_val, uint64_t cond, uint64_t false_val, uint64_t* out) { OUT(select(true_val, cond, false_val)); }
hj_status hj_tr_extract(uint64_t a, uint32_t elem, uint64_t* out) { OUT(extract(a, elem)); }
hj_status hj_tr_extract_dyn(uint64_t a, uint64_t elem, uint64_t* out) { OUT(extract_dyn(a, elem)); }
hj_status hj_tr_composite(const uint64_t* refs, uint32_t n, uint64_t* out) { OUT(composite(vec_of(refs, n))); }
hj_status hj_tr_vec(const uint64_t* refs, uint32_t n, uint64_t* out) { OUT(vec(vec_of(refs, n))); }
hj_status hj_tr_arr(const uint64_t* refs, uint32_t n, uint64_t* out) { OUT(arr(vec_of(refs, n))); }
hj_status hj_tr_mat(const uint64_t* columns, uint32_t n, uint64_t* out) { OUT(mat(vec_of(columns, n))); }
hj_status hj_tr_gather(uint64_t src, uint64_t idx, uint64_t active, uint64_t* out) { OUT(gather_if(src, idx, active)); }
hj_status hj_tr_scatter(uint64_t src, uint64_t dst, uint64_t idx, uint64_t active) {
    return guarded([&] { scatter_like(HJ_OP_SCATTER, 0, src, dst, idx, active); });
}
hj_status hj_tr_scatter_reduce(uint64_t src, uint64_t dst, uint64_t idx, uint64_t active, uint32_t op) {
    return guarded([&] { scatter_like(HJ_OP_SCATTER_REDUCE, op, src, dst, idx, active); });
}
hj_status hj_tr_scatter_atomic(uint64_t src, uint64_t dst, uint64_t idx, uint64_t active, uint32_t op, uint64_t* out) {
    OUT(scatter_like(HJ_OP_SCATTER_ATOMIC, op, src, dst, idx, active));
}
hj_status hj_tr_atomic_inc(uint64_t dst, uint64_t idx, uint64_t active, uint64_t* out) { OUT(atomic_inc(dst, idx, active)); }
hj_status hj_tr_prefix_sum(uint64_t a, int32_t inclusive, uint64_t* out) { OUT(prefix_sum(a, inclusive != 0)); }
hj_status hj_tr_reduce(uint64_t a, uint32_t op, uint64_t* out) { OUT(reduce(a, op)); }
hj_status hj_tr_compress(uint64_t mask, uint64_t* out_count, uint64_t* out_index) {
    return guarded([&] { compress(mask, out_count, out_index); });
}
hj_status hj_tr_compress_dyn(uint64_t mask, uint64_t* out) { OUT(compress_dyn(mask)); }
#undef OUT

// loop_start / if_start: `state_out` receives n new references (the extracted state)
hj_status hj_tr_scope_start(int32_t is_loop, const uint64_t* state, uint32_t n, uint64_t* out_scope, uint64_t* state_out) {
    return guarded([&] {
        std::vector<VarId> so;
        *out_scope = scope_start(is_loop != 0, vec_of(state, n), &so);
        for (uint32_t i = 0; i < n; i++) state_out[i] = so.at(i);
    });
}
hj_status hj_tr_scope_end(uint64_t scope, const uint64_t* state, uint32_t n, uint64_t* state_out) {
    return guarded([&] {
        std::vector<VarId> so;
        scope_end(scope, vec_of(state, n), &so);
        for (uint32_t i = 0; i < n; i++) state_out[i] = so.at(i);
    });
}

hj_status hj_tr_schedule(uint64_t v) { return guarded([&] { schedule(v); }); }
hj_status hj_tr_schedule_eval(void) { return guarded([&] { schedule_eval(); }); }
hj_status hj_tr_reset_schedule(void) { return guarded([&] { t_ts.clear(); }); }

// number of elements currently valid (reads the device-resident count of a DynSize extent)
hj_status hj_tr_var_size(uint64_t v, uint64_t* out) { return guarded([&] { *out = current_size(v); }); }
hj_status hj_tr_to_host(uint64_t v, uint64_t start_elem, uint64_t n_elem, void* dst) {
    return guarded([&] { to_host(v, start_elem, n_elem, dst); });
}

// ---- graphs -------------------------------------------------------------------------------------------
// tr::compile() (trace.rs:528-536)
hj_status hj_tr_compile(hj_graph** out) {
    return guarded([&] {
        schedule_eval();
        Graph* g = compile_graph(t_ts, {}, {});
        *out = new hj_graph();
        (*out)->g.reset(g);
    });
}
// graph::compile(&ts, &inputs, &outputs) as used by FCache::call (record.rs:168-193)
hj_status hj_tr_compile_fn(const uint64_t* inputs, uint32_t n_in, const uint64_t* outputs, uint32_t n_out, hj_graph** out) {
    return guarded([&] {
        for (uint32_t i = 0; i < n_out; i++) schedule(outputs[i]);
        schedule_eval();
        Graph* g = compile_graph(t_ts, vec_of(inputs, n_in), vec_of(outputs, n_out));
        *out = new hj_graph();
        (*out)->g.reset(g);
    });
}
hj_status hj_graph_retain(hj_graph* g) {
    HJ_REQUIRE(g, "hj_graph_retain: null graph");
    g->rc.fetch_add(1);
    return HJ_OK;
}
hj_status hj_graph_release(hj_graph* g) {
    if (g && g->rc.fetch_sub(1) == 1) {
        try { delete g; } catch (...) {}
    }
    return HJ_OK;
}
uint32_t hj_graph_n_passes(hj_graph* g) { return g ? (uint32_t)g->g->passes.size() : 0; }
uint32_t hj_graph_n_outputs(hj_graph* g) { return g ? (uint32_t)g->g->outputs.size() : 0; }
hj_status hj_graph_debug_string(hj_graph* g, char** out) {
    HJ_REQUIRE(g && out, "hj_graph_debug_string: null argument");
    return guarded([&] {
        std::string s = graph_debug_string(*g->g);
        *out = (char*)malloc(s.size() + 1);
        memcpy(*out, s.c_str(), s.size() + 1);
    });
}
// The flat IR of kernel pass `pass` (what hj_execute_graph hands to the NVRTC stage), valid while the
// graph is alive; *out = NULL for device-op passes.  Lets a host inspect / pre-compile the kernels of a
// graph (hj_ir_codegen, hj_ir_compile_cubin) without launching it.
hj_status hj_graph_pass_ir(hj_graph* g, uint32_t pass, const hj_ir** out) {
    HJ_REQUIRE(g && out, "hj_graph_pass_ir: null argument");
    HJ_REQUIRE(pass < g->g->passes.size(), "hj_graph_pass_ir: pass %u out of range (%zu passes)", pass, g->g->passes.size());
    std::lock_guard<std::mutex> lock(g->views_mu);
    if (g->ir_views.empty()) {
        g->ir_views.resize(g->g->passes.size());
        for (size_t i = 0; i < g->g->passes.size(); i++)
            if (g->g->passes[i].is_kernel) g->ir_views[i] = g->g->passes[i].ir.view();
    }
    *out = g->g->passes[pass].is_kernel ? &g->ir_views[pass] : nullptr;
    return HJ_OK;
}

// Wire format of a compiled graph (tgraph_io.cpp).  *bytes_out is malloc'ed: free with hj_free_string.
hj_status hj_graph_serialize(hj_graph* g, void** bytes_out, size_t* n_out) {
    HJ_REQUIRE(g && bytes_out && n_out, "hj_graph_serialize: null argument");
    return guarded([&] {
        const std::vector<uint8_t> v = serialize_graph(*g->g);
        void* p = malloc(v.size());
        if (!p) throw TraceError("hj_graph_serialize: out of host memory");
        memcpy(p, v.data(), v.size());
        *bytes_out = p;
        *n_out = v.size();
    });
}
hj_status hj_graph_deserialize(hj_device* dev, const void* bytes, size_t n, hj_graph** out) {
    HJ_REQUIRE(bytes && out, "hj_graph_deserialize: null argument");
    return guarded([&] {
        Graph* g = deserialize_graph(dev, bytes, n);
        *out = new hj_graph();
        (*out)->g.reset(g);
    });
}
// Graph::launch_with (graph.rs:192-400).  outputs_out receives hj_graph_n_outputs new references.
hj_status hj_graph_launch_sharded(hj_graph* g, hj_device* dev, hj_comm* comm, const uint64_t* inputs, uint32_t n_in,
                                  uint64_t* outputs_out, hj_graph_report* report) {
    HJ_REQUIRE(g && dev, "hj_graph_launch: null argument");
    return guarded([&] {
        std::vector<VarId> outs;
        LaunchReport lr;
        std::vector<hj_pass_report> pr;
        hj_report br = {};
        const bool timed = report && report->passes && report->passes_capacity >= g->g->passes.size();
        if (timed) {
            br.passes = report->passes;
            br.passes_capacity = report->passes_capacity;
        }
        launch_graph(*g->g, dev, vec_of(inputs, n_in), &outs, &lr, timed ? &br : nullptr, comm);  // per-pass timings need the pass-by-pass path
        if (outputs_out)
            for (size_t i = 0; i < outs.size(); i++) outputs_out[i] = outs[i];
        else
            for (VarId o : outs) ref_drop(o);
        if (report) {
            report->aliasing_rate = lr.aliasing_rate;
            report->aliasing_duration_us = lr.aliasing_duration_us;
            report->n_passes = lr.n_passes;
            report->backend_cpu_us = lr.backend_cpu_us;
        }
    });
}
hj_status hj_graph_launch(hj_graph* g, hj_device* dev, const uint64_t* inputs, uint32_t n_in, uint64_t* outputs_out,
                          hj_graph_report* report) {
    return hj_graph_launch_sharded(g, dev, nullptr, inputs, n_in, outputs_out, report);
}

// ---- function cache (record.rs:116-210): key -> compiled graph ------------------------------------------
hj_status hj_fcache_get(uint64_t key, hj_graph** out) {  // *out = NULL on a miss; retained on a hit
    HJ_REQUIRE(out, "hj_fcache_get: null argument");
    std::lock_guard<std::mutex> l(g_fcache_mu);
    auto it = g_fcache.find(key);
    *out = it == g_fcache.end() ? nullptr : it->second;
    if (*out) (*out)->rc.fetch_add(1);
    return HJ_OK;
}
hj_status hj_fcache_put(uint64_t key, hj_graph* g) {
    HJ_REQUIRE(g, "hj_fcache_put: null graph");
    std::lock_guard<std::mutex> l(g_fcache_mu);
    auto it = g_fcache.find(key);
    if (it != g_fcache.end()) hj_graph_release(it->second);
    g->rc.fetch_add(1);
    g_fcache[key] = g;
    return HJ_OK;
}
hj_status hj_fcache_clear(void) {
    std::lock_guard<std::mutex> l(g_fcache_mu);
    for (auto& kv : g_fcache) hj_graph_release(kv.second);
    g_fcache.clear();
    return HJ_OK;
}
uint64_t hj_fcache_size(void) {
    std::lock_guard<std::mutex> l(g_fcache_mu);
    return g_fcache.size();
}

}  // extern "C"
