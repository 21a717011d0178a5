// tgraph.cpp — schedule -> Graph (passes + resources), trace -> IR lowering, Graph launch.
//
// Restates hephaestus-jit/src/graph.rs (compile :436-614, launch_with :192-400) and
// hephaestus-jit/src/compiler.rs (:20-230).  The pass list it produces is handed to
// hj_execute_graph (graph_exec.cpp) — the BackendDevice::execute_graph boundary.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <tuple>
#include <map>
#include <sstream>
#include <unordered_map>

#include <atomic>

#include "trace_internal.h"

namespace hj {
namespace tr {

hj_ir KernelIR::view() const {
    hj_ir ir;
    ir.vars = vars.data();
    ir.n_vars = (uint32_t)vars.size();
    ir.deps = deps.data();
    ir.n_deps = (uint32_t)deps.size();
    ir.types = types.data();
    ir.n_types = (uint32_t)types.size();
    ir.struct_fields = struct_fields.data();
    ir.n_struct_fields = (uint32_t)struct_fields.size();
    ir.n_buffers = n_buffers;
    return ir;
}

Graph::~Graph() {
    if (launched_on && uid) hj_graph_cache_drop(launched_on, uid);
    for (GraphResource& r : resources)
        if (r.kind == GraphResource::Captured && r.id != NO_VAR) ref_drop(r.id);
    for (hj_buffer* b : seed_cache)
        if (b) hj_buffer_release(b);
}

namespace {

// ---- compiler.rs -------------------------------------------------------------------------------
struct Compiler {
    Trace& trace;
    KernelIR ir;
    std::unordered_map<VarId, uint32_t> visited;
    // trivial-variable CSE (compiler.rs:206-230): key = (ty, op, arg, data) of dependency-free vars
    std::map<std::tuple<uint32_t, uint32_t, uint32_t, uint64_t>, uint32_t> trivial;
    std::vector<VarId> buffers;  // IndexSet<trace::VarId>: first-touch order = IR buffer slots
    std::unordered_map<TypeId, uint32_t> local_types;

    explicit Compiler(Trace& t) : trace(t) {}

    // global interned type -> index in this IR's own type table (children before parents)
    uint32_t local_type(TypeId t) {
        auto it = local_types.find(t);
        if (it != local_types.end()) return it->second;
        TypeNode n = type_node(t);
        hj_type_desc d = {};
        d.kind = n.kind;
        if (n.kind == HJ_VEC || n.kind == HJ_ARRAY) {
            d.elem = local_type(n.elem);
            d.num = n.num;
        } else if (n.kind == HJ_MAT) {
            d.elem = local_type(n.elem);
            d.cols = n.cols;
            d.rows = n.rows;
        } else if (n.kind == HJ_STRUCT) {
            std::vector<uint32_t> f;
            for (TypeId ft : n.fields) f.push_back(local_type(ft));
            d.num = (uint32_t)f.size();
            d.first_field = (uint32_t)ir.struct_fields.size();
            ir.struct_fields.insert(ir.struct_fields.end(), f.begin(), f.end());
        }
        uint32_t idx = (uint32_t)ir.types.size();
        ir.types.push_back(d);
        local_types[t] = idx;
        return idx;
    }

    uint32_t push_buffer(VarId id) {
        for (size_t i = 0; i < buffers.size(); i++)
            if (buffers[i] == id) return (uint32_t)i;
        buffers.push_back(id);
        return (uint32_t)(buffers.size() - 1);
    }

    uint32_t push_var(uint32_t op, uint32_t arg, TypeId ty, uint64_t data, const std::vector<uint32_t>& deps) {
        const uint32_t lt = local_type(ty);
        const bool triv = deps.empty();
        auto key = std::make_tuple(lt, op, arg, data);
        if (triv) {
            auto it = trivial.find(key);
            if (it != trivial.end()) return it->second;
        }
        hj_ir_var v = {};
        v.ty = lt;
        v.op = op;
        v.arg = arg;
        v.dep_start = (uint32_t)ir.deps.size();
        ir.deps.insert(ir.deps.end(), deps.begin(), deps.end());
        v.dep_end = (uint32_t)ir.deps.size();
        v.data = data;
        uint32_t id = (uint32_t)ir.vars.size();
        ir.vars.push_back(v);
        if (triv) trivial[key] = id;
        return id;
    }

    // compiler.rs:135-186
    uint32_t collect_data(VarId id) {
        // The reference looks `id` up in `visited` first (compiler.rs:137-139) — the map of VALUES.  A
        // variable that one kernel uses both as a value and through a reference (`b < b` next to
        // `b.gather(i)`), value first, then yields Gather(value, i): GLSL that does not compile.  The
        // BufferRef itself is de-duplicated by the trivial-variable CSE of push_var.
        const Var& var = trace.var(id);
        Op r = resulting_op(var.op);
        if (r.kind != OpKind::Buffer) throw TraceError("reference to a variable that does not evaluate to a buffer");
        uint32_t slot = push_buffer(id);
        return push_var(HJ_OP_BUFFER_REF, 0, var.ty, slot, {});
    }

    // compiler.rs:66-134.  The reference recurses over the dependencies; a chain of a few hundred
    // thousand dependent operations (its own `compile` bench traces 10 000) would overflow the stack,
    // so the same post-order walk — dependencies left to right, then the variable itself, which is
    // what fixes the numbering of the IR — runs on an explicit stack here.
    bool collect_leaf(VarId id, uint32_t* out) {  // everything that needs no walk of its own
        const Var& var = trace.var(id);
        switch (var.op.kind) {
        case OpKind::Ref: *out = collect_data(var.deps.at(0)); return true;
        case OpKind::Buffer: {
            uint32_t data = collect_data(id);
            uint32_t idx = push_var(HJ_OP_INDEX, 0, type_scalar(HJ_U32), 0, {});
            *out = push_var(HJ_OP_GATHER, 0, var.ty, 0, {data, idx});
            return true;
        }
        case OpKind::KernelOp:
            if (var.op.code != HJ_OP_LITERAL) return false;
            *out = push_var(HJ_OP_LITERAL, 0, var.ty, var.data.lit, {});
            return true;
        default:
            throw TraceError("a kernel depends on an unevaluated device op (todo!() in the reference, compiler.rs:131)");
        }
    }
    uint32_t collect(VarId root) {
        auto hit = visited.find(root);
        if (hit != visited.end()) return hit->second;
        uint32_t out;
        if (collect_leaf(root, &out)) {
            visited[root] = out;
            return out;
        }
        struct Frame {
            VarId id;
            std::vector<VarId> tdeps;
            std::vector<uint32_t> deps;
        };
        std::vector<Frame> stack;
        stack.push_back({root, trace.var(root).deps, {}});
        while (true) {
            Frame& f = stack.back();
            if (f.deps.size() < f.tdeps.size()) {
                const VarId d = f.tdeps[f.deps.size()];
                auto it = visited.find(d);
                if (it != visited.end()) {
                    f.deps.push_back(it->second);
                } else if (collect_leaf(d, &out)) {
                    visited[d] = out;
                    stack.back().deps.push_back(out);
                } else {
                    stack.push_back({d, trace.var(d).deps, {}});  // invalidates f
                }
                continue;
            }
            const Var& v = trace.var(f.id);
            out = push_var(v.op.code, v.op.arg, v.ty, 0, f.deps);
            visited[f.id] = out;
            stack.pop_back();
            if (stack.empty()) return out;
            stack.back().deps.push_back(out);
        }
    }

    // compiler.rs:20-65
    void compile(const VarId* ids, size_t n) {
        for (size_t i = 0; i < n; i++) {
            uint32_t src = collect(ids[i]);
            const Var& var = trace.var(ids[i]);
            if (type_size(var.ty) == 0) continue;  // side effects have no output store
            uint32_t slot = push_buffer(ids[i]);
            uint32_t dst = push_var(HJ_OP_BUFFER_REF, 0, var.ty, slot, {});
            uint32_t idx = push_var(HJ_OP_INDEX, 0, type_scalar(HJ_U32), 0, {});
            push_var(HJ_OP_SCATTER, 0, type_scalar(HJ_VOID), 0, {dst, src, idx});
        }
        ir.n_buffers = (uint32_t)buffers.size();
    }
};

// Extent::partial_cmp (extent.rs:25-44): DynSize sorts before Size
int extent_cmp(const Extent& a, const Extent& b) {
    auto cmp = [](uint64_t x, uint64_t y) { return x < y ? -1 : (x > y ? 1 : 0); };
    if (!a.dynamic && !b.dynamic) return cmp(a.n, b.n);
    if (a.dynamic && !b.dynamic) return -1;
    if (!a.dynamic && b.dynamic) return 1;
    int c = cmp(a.n, b.n);
    if (c) return c;
    // slotmap KeyData orders by index, then version
    c = cmp((uint32_t)a.size_var, (uint32_t)b.size_var);
    return c ? c : cmp(a.size_var >> 32, b.size_var >> 32);
}

struct GraphBuilder {  // graph.rs:68-112
    Trace& trace;
    std::vector<VarId> keys;  // IndexMap<VarId, ResourceDesc>
    std::vector<BufferDesc> descs;
    explicit GraphBuilder(Trace& t) : trace(t) {}
    // returns -1 where the reference returns Err(ResourceMissmatch)
    int try_push_resource(VarId id) {
        for (size_t i = 0; i < keys.size(); i++)
            if (keys[i] == id) return (int)i;
        Var* var = trace.get(id);
        if (!var) return -1;
        Op r;
        if (var->op.kind == OpKind::Nop || var->op.kind == OpKind::Ref) return -1;
        r = resulting_op(var->op);
        if (r.kind != OpKind::Buffer) return -1;
        BufferDesc d;
        d.size = var->extent.n;  // extent.capacity()
        d.ty = var->ty;
        trace.inc_rc(id);  // keeps the variable alive until the graph is assembled
        keys.push_back(id);
        descs.push_back(d);
        return (int)keys.size() - 1;
    }
};

}  // namespace

// graph::compile (graph.rs:436-614)
Graph* compile_graph(ThreadState& ts, const std::vector<VarId>& inputs, const std::vector<VarId>& outputs) {
    std::unique_ptr<Graph> graph(new Graph());
    static std::atomic<uint64_t> next_uid{1};
    graph->uid = next_uid.fetch_add(1);
    // The reference panics on the errors below (the process is gone); here they are reported to the
    // caller, so nothing may be left behind: the builder's temporary references and the schedule
    // (which graph::compile consumes either way, trace.rs:528-536) are released on every path.
    try {
        std::lock_guard<std::mutex> lock(g_trace_mu);
        Trace& trace = g_trace;
        GraphBuilder gb(trace);
        try {
        auto need = [&](VarId id) {
            int r = gb.try_push_resource(id);
            if (r < 0) throw TraceError("Resource does not match variable type! (graph::Error::ResourceMissmatch)");
            return (uint32_t)r;
        };
        for (VarId id : inputs) need(id);
        for (VarId id : outputs) need(id);
        for (VarId id : inputs) graph->inputs.push_back(need(id));
        for (VarId id : outputs) graph->outputs.push_back(need(id));

        std::vector<VarId> vars = ts.scheduled;
        // subdivide every group by extent (graph.rs:463-495): stable sort, then split on change
        std::vector<std::pair<size_t, size_t>> groups;
        for (auto& grp : ts.groups) {
            std::stable_sort(vars.begin() + grp.first, vars.begin() + grp.second, [&](VarId a, VarId b) {
                return extent_cmp(trace.var(a).extent, trace.var(b).extent) < 0;
            });
            Extent size = trace.var(vars[grp.first]).extent;
            size_t start = grp.first;
            for (size_t i = grp.first; i < grp.second; i++) {
                if (trace.var(vars[i]).extent != size) {
                    if (start != i) groups.emplace_back(start, i);
                    size = trace.var(vars[i]).extent;
                    start = i;
                }
            }
            if (start != grp.second) groups.emplace_back(start, grp.second);
        }

        for (auto& grp : groups) {
            const VarId first_id = vars[grp.first];
            const Extent extent = trace.var(first_id).extent;
            Pass pass;
            if (extent.dynamic) pass.size_buffer = gb.try_push_resource(extent.size_var);
            const Op first_op = trace.var(first_id).op;
            if (first_op.kind == OpKind::DeviceOp) {
                if (grp.second - grp.first != 1) throw TraceError("a device op must be alone in its group (graph.rs:513)");
                std::vector<VarId> ids = {first_id};
                const std::vector<VarId> deps = trace.var(first_id).deps;
                ids.insert(ids.end(), deps.begin(), deps.end());
                for (VarId id : ids) {  // flat_map: variables without a resource are dropped
                    int r = gb.try_push_resource(id);
                    if (r >= 0) pass.resources.push_back((uint32_t)r);
                }
                pass.is_kernel = false;
                pass.device_op = first_op;
            } else {
                Compiler c(trace);
                c.compile(vars.data() + grp.first, grp.second - grp.first);
                for (VarId id : c.buffers) {
                    int r = gb.try_push_resource(id);
                    if (r >= 0) pass.resources.push_back((uint32_t)r);
                }
                pass.is_kernel = true;
                pass.ir = std::move(c.ir);
                pass.size = extent.n;
            }
            graph->passes.push_back(std::move(pass));
            for (size_t i = grp.first; i < grp.second; i++) trace.advance(vars[i]);
        }

        // classify resources (graph.rs:575-597)
        for (size_t i = 0; i < gb.keys.size(); i++) {
            GraphResource r;
            const VarId id = gb.keys[i];
            const bool is_input = std::find(graph->inputs.begin(), graph->inputs.end(), (uint32_t)i) != graph->inputs.end();
            if (is_input) {
                r.kind = GraphResource::Input;
            } else if (trace.var(id).data.kind == Resource::Buffer) {
                r.kind = GraphResource::Captured;
                trace.inc_rc(id);
                r.id = id;
            } else {
                r.kind = GraphResource::Internal;
                r.id = id;
            }
            trace.dec_rc(id);  // the builder's temporary reference
            graph->resources.push_back(r);
        }
        graph->resource_descs = gb.descs;
        } catch (...) {
            for (size_t i = graph->resources.size(); i < gb.keys.size(); i++)
                if (trace.get(gb.keys[i])) trace.dec_rc(gb.keys[i]);
            throw;
        }
    } catch (...) {
        ts.clear();
        throw;
    }
    ts.clear();  // the schedule's references are dropped with the ThreadState (trace.rs:528-536)
    return graph.release();
}

namespace {
size_t round_pow2(size_t x) {  // utils.rs:29-39
    if (x <= 1) return x;
    size_t p = 1;
    while (p < x) p <<= 1;
    return p;
}
hj_buffer* create_buffer(hj_device* dev, const BufferDesc& d) {  // Resource::create (resource.rs:38-48)
    hj_buffer* b = nullptr;
    size_t bytes = d.size * type_size(d.ty);
    if (hj_buffer_create(dev, bytes ? bytes : 1, &b) != HJ_OK)
        throw TraceError(std::string("create_buffer failed: ") + hj_last_error());
    return b;
}
uint32_t shard_state(const Var& v) {  // hj_shard_desc.deferred of the variable's buffer
    if (v.data.segment) return v.data.segment_local ? HJ_SHARD_SEGMENT_LOCAL : HJ_SHARD_SEGMENT;
    return v.data.deferred ? HJ_SHARD_DEFERRED : HJ_SHARD_PLAIN;
}
uint32_t scalar_kind(TypeId t) {
    TypeNode n = type_node(t);
    return n.kind;
}
}  // namespace

// Graph::launch_with (graph.rs:192-400)
void launch_graph(const Graph& g, hj_device* dev, const std::vector<VarId>& inputs, std::vector<VarId>* outputs,
                  LaunchReport* report, hj_report* backend_report, hj_comm* comm) {
    const size_t nres = g.resources.size();
    std::vector<hj_buffer*> res(nres, nullptr);  // each non-null entry owns one reference
    std::vector<bool> seed_from_cache(nres, false);
    std::vector<hj_shard_desc> shards(nres);     // sharded launches: placement / deferred seed per resource
    for (hj_shard_desc& sd : shards) {
        sd.placement = HJ_RES_AUTO;
        sd.deferred = 0;
        sd.seed = nullptr;  // owns one reference when set
    }
    auto release_all = [&]() {
        for (hj_buffer* b : res)
            if (b) hj_buffer_release(b);
        for (hj_shard_desc& sd : shards)
            if (sd.seed) hj_buffer_release(sd.seed);
    };
    try {
        // what Graph::launch_with hands to BackendDevice::execute_graph (graph.rs:315-323): built first,
        // a sharded launch plans the placement of every resource from it before buffers are created
        std::vector<hj_pass> passes(g.passes.size());
        std::vector<hj_ir> irs(g.passes.size());
        for (size_t p = 0; p < g.passes.size(); p++) {
            const Pass& src = g.passes[p];
            hj_pass& dst = passes[p];
            memset(&dst, 0, sizeof(dst));
            dst.resources = src.resources.data();
            dst.n_resources = (uint32_t)src.resources.size();
            dst.size_buffer = src.size_buffer;
            if (src.is_kernel) {
                irs[p] = src.ir.view();
                dst.kind = HJ_PASS_KERNEL;
                dst.ir = &irs[p];
                dst.size = src.size;
            } else {
                switch (src.device_op.code) {
                case DOP_REDUCE: dst.kind = HJ_PASS_REDUCE; break;
                case DOP_PREFIX_SUM: dst.kind = HJ_PASS_PREFIX_SUM; break;
                default: dst.kind = HJ_PASS_COMPRESS; break;
                }
                dst.arg = src.device_op.arg;
            }
        }
        std::vector<hj_buffer_desc> descs(nres);
        for (size_t i = 0; i < nres; i++) {
            descs[i].size = g.resource_descs[i].size;
            descs[i].ty = scalar_kind(g.resource_descs[i].ty);
            descs[i].elem_bytes = (uint32_t)type_size(g.resource_descs[i].ty);
        }
        auto bind_var = [&](size_t rid, const Var& var) {  // trace lock held
            hj_buffer_retain(var.data.buf);
            if (res[rid]) hj_buffer_release(res[rid]);
            res[rid] = var.data.buf;
            if (shards[rid].seed) hj_buffer_release(shards[rid].seed);
            shards[rid].seed = nullptr;
            if (var.data.comm) {
                if (comm && comm != var.data.comm) throw TraceError("variables sharded over different communicators in one launch");
                comm = var.data.comm;
                shards[rid].placement = HJ_RES_SHARDED;
                shards[rid].deferred = shard_state(var);
                if (var.data.seed) {
                    hj_buffer_retain(var.data.seed);
                    shards[rid].seed = var.data.seed;
                }
            } else {
                shards[rid].placement = HJ_RES_REPLICATED;
            }
        };
        {
            std::lock_guard<std::mutex> lock(g_trace_mu);
            // inputs by position (graph.rs:205-219)
            for (size_t i = 0; i < inputs.size() && i < g.inputs.size(); i++) {
                const Var& var = g_trace.var(inputs[i]);
                const BufferDesc& d = g.resource_descs[g.inputs[i]];
                BufferDesc have;
                have.size = var.extent.n;
                have.ty = var.ty;
                if (!(have == d) || var.data.kind != Resource::Buffer)
                    throw TraceError("Resource does not match variable type! (graph::Error::ResourceMissmatch)");
                bind_var(g.inputs[i], var);
            }
            // captured resources (graph.rs:220-228)
            for (size_t i = 0; i < nres; i++) {
                if (res[i]) continue;
                const GraphResource& gr = g.resources[i];
                if (gr.kind == GraphResource::Captured) {
                    const Var& var = g_trace.var(gr.id);
                    if (var.data.kind == Resource::Buffer) bind_var(i, var);
                }
            }
        }
        // ---- sharded launch: where does every resource live, and how large is this rank's part
        int32_t rank = 0, world = 1;
        if (comm) {
            if (hj_comm_info(comm, &rank, &world, nullptr, nullptr) != HJ_OK)
                throw TraceError(std::string("hj_comm_info failed: ") + hj_last_error());
            if (hj_shard_plan(passes.data(), (uint32_t)passes.size(), descs.data(), (uint32_t)nres, shards.data()) != HJ_OK)
                throw TraceError(std::string("hj_shard_plan failed: ") + hj_last_error());
            // a sharded integer scan result keeps its cross-GPU offset beside it (8 instead of 12 bytes
            // per element; the consumers add it): give every such destination a seed slot
            auto give_seed = [&](uint32_t rid) {
                if (shards[rid].placement != HJ_RES_SHARDED || shards[rid].seed) return;
                if (g.seed_cache.size() < nres) g.seed_cache.resize(nres, nullptr);
                if (!g.seed_cache[rid] && hj_buffer_create(dev, 16, &g.seed_cache[rid]) != HJ_OK)
                    throw TraceError(std::string("create_buffer failed: ") + hj_last_error());
                hj_buffer_retain(g.seed_cache[rid]);
                shards[rid].seed = g.seed_cache[rid];
                seed_from_cache[rid] = true;
            };
            for (const hj_pass& p : passes) {
                if (p.kind == HJ_PASS_PREFIX_SUM && p.n_resources >= 2) {
                    const uint32_t k = descs[p.resources[0]].ty;
                    if (k >= HJ_I8 && k <= HJ_U64) give_seed(p.resources[0]);
                }
                // the index segment of a sharded Compress keeps the rank's own count beside it, and so does
                // everything a DynSize kernel writes over such a segment (hj.h: HJ_SHARD_SEGMENT)
                if (p.kind == HJ_PASS_COMPRESS && p.n_resources >= 3) give_seed(p.resources[0]);
                if (p.kind == HJ_PASS_KERNEL && p.size_buffer >= 0)
                    for (uint32_t b = 0; b < p.n_resources; b++)
                        if (descs[p.resources[b]].size == p.size) give_seed(p.resources[b]);
            }
        }
        auto local_desc = [&](size_t rid, BufferDesc d) {  // what this rank allocates for the resource
            if (comm && shards[rid].placement == HJ_RES_SHARDED) {
                uint64_t s0 = 0, s1 = 0;
                hj_shard_bounds(d.size, world, rank, &s0, &s1);
                d.size = (size_t)(s1 - s0);
            }
            return d;
        };
        {
            std::lock_guard<std::mutex> lock(g_trace_mu);
            // live internal resources (graph.rs:229-235)
            for (size_t i = 0; i < nres; i++) {
                if (res[i]) continue;
                const GraphResource& gr = g.resources[i];
                if (gr.kind == GraphResource::Internal && g_trace.get(gr.id))
                    res[i] = create_buffer(dev, local_desc(i, g.resource_descs[i]));
            }
        }
        // lifetime-based aliasing of dead internal resources (graph.rs:237-296).  The reference
        // re-inserts a buffer into its cache the moment the last pass using it comes up, so a later
        // resource of the SAME pass can pop it (its own test expects a hit rate of 2/3,
        // test.rs:1638-1670).  That is only sound when no thread of the pass can observe another
        // thread's store: here a buffer released in pass p is handed to a resource born in pass p
        // only if p is a kernel that reads the one and writes the other at the bare Index (read
        // before write, ir.h: slots_may_share) — x = x + 1 in place.  Every other release (gathers
        // through computed indices, device ops) becomes available from pass p + 1 on.
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::pair<size_t, size_t>> life(nres, {g.passes.size(), 0});
        for (size_t p = 0; p < g.passes.size(); p++)
            for (uint32_t rid : g.passes[p].resources) {
                life[rid].first = std::min(life[rid].first, p);
                life[rid].second = std::max(life[rid].second, p);
            }
        std::vector<bool> internal(nres);
        size_t n_internal = 0, misses = 0;
        for (size_t i = 0; i < nres; i++) {
            internal[i] = res[i] == nullptr;
            n_internal += internal[i];
        }
        std::vector<bool> is_output(nres, false);
        for (uint32_t rid : g.outputs) is_output[rid] = true;
        using Key = std::tuple<size_t, TypeId, uint32_t>;  // pow2-rounded desc + placement (a shard is smaller than a replica)
        std::map<Key, std::vector<hj_buffer*>> cache;
        struct Released { Key key; hj_buffer* buf; size_t slot; };
        std::vector<SlotAccess> access;
        for (size_t p = 0; p < g.passes.size(); p++) {
            const Pass& pass = g.passes[p];
            std::vector<Released> released;  // by this pass; each entry owns a reference
            access.clear();
            for (size_t k = 0; k < pass.resources.size(); k++) {
                const uint32_t rid = pass.resources[k];
                if (!internal[rid]) continue;
                BufferDesc rounded = g.resource_descs[rid];
                rounded.size = round_pow2(rounded.size);
                const Key key = std::make_tuple(rounded.size, rounded.ty, comm ? shards[rid].placement : 0u);
                if (life[rid].first == p && !res[rid]) {
                    // in-place reuse of a buffer this very pass has just released
                    if (pass.is_kernel && !released.empty()) {
                        if (access.empty()) {
                            hj_ir view = pass.ir.view();
                            analyse_slot_access(&view, &access);
                        }
                        for (size_t q = 0; q < released.size() && !res[rid]; q++)
                            if (released[q].key == key && k < access.size() && released[q].slot < access.size() &&
                                slots_may_share(access[released[q].slot], access[k])) {
                                res[rid] = released[q].buf;
                                released.erase(released.begin() + (long)q);
                            }
                    }
                    auto& bucket = cache[key];
                    if (res[rid]) {
                    } else if (!bucket.empty()) {
                        res[rid] = bucket.back();
                        bucket.pop_back();
                    } else {
                        misses++;
                        res[rid] = create_buffer(dev, local_desc(rid, rounded));
                    }
                }
                // (graph.rs:283-290 re-inserts every dead internal buffer; an OUTPUT of a recorded
                // function is such a resource once the variables of the first trace are gone, and a
                // later pass would overwrite it before launch_with hands it back — outputs stay out
                // of the cache here)
                if (life[rid].second == p && res[rid] && !is_output[rid]) {
                    hj_buffer_retain(res[rid]);
                    released.push_back({key, res[rid], k});
                }
            }
            for (Released& r : released) cache[r.key].push_back(r.buf);
        }
        for (auto& kv : cache)
            for (hj_buffer* b : kv.second) hj_buffer_release(b);
        auto t1 = std::chrono::steady_clock::now();
        if (report) {
            report->aliasing_rate = n_internal ? 1.f - (float)misses / (float)n_internal : 0.f;
            report->aliasing_duration_us = std::chrono::duration<double, std::micro>(t1 - t0).count();
            report->n_passes = (uint32_t)g.passes.size();
        }
        for (size_t i = 0; i < nres; i++)  // graph.rs:309-312: every resource must be bound
            if (!res[i])
                throw TraceError("A resource in the environment has been left empty! (graph::Error::UninitializedResourve)");

        // ---- the backend boundary: BackendDevice::execute_graph (graph.rs:315-323)
        if (!g.passes.empty()) {
            // Relaunches of a recorded graph with the same buffers replay ONE captured CUDA graph
            // instead of enqueueing pass by pass (sharded launches too: their exchange epochs live in
            // device memory); a launch that wants per-pass timings cannot.
            hj_status s;
            auto b0 = std::chrono::steady_clock::now();
            if (comm && backend_report) {
                s = hj_execute_graph_sharded(comm, passes.data(), (uint32_t)passes.size(), res.data(), descs.data(),
                                             (uint32_t)nres, shards.data(), backend_report);
            } else if (comm) {
                g.launched_on = dev;
                s = hj_execute_graph_sharded_cached(comm, g.uid, passes.data(), (uint32_t)passes.size(), res.data(),
                                                    descs.data(), (uint32_t)nres, shards.data(), nullptr);
            } else if (backend_report) {
                s = hj_execute_graph(dev, passes.data(), (uint32_t)passes.size(), res.data(), descs.data(),
                                     (uint32_t)nres, backend_report);
            } else {
                g.launched_on = dev;
                s = hj_execute_graph_cached(dev, g.uid, passes.data(), (uint32_t)passes.size(), res.data(),
                                            descs.data(), (uint32_t)nres, nullptr);
            }
            if (s != HJ_OK) throw TraceError(std::string("execute_graph failed: ") + hj_last_error());
            if (report)
                report->backend_cpu_us =
                    std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - b0).count();
        }

        // ---- write results back (graph.rs:332-393)
        auto resource_of = [&](size_t rid) {
            Resource r;
            r.kind = Resource::Buffer;
            r.buf = res[rid];
            if (comm && shards[rid].placement == HJ_RES_SHARDED) {
                r.comm = comm;
                r.deferred = shards[rid].deferred == HJ_SHARD_DEFERRED;
                r.segment = shards[rid].deferred == HJ_SHARD_SEGMENT || shards[rid].deferred == HJ_SHARD_SEGMENT_LOCAL;
                r.segment_local = shards[rid].deferred == HJ_SHARD_SEGMENT_LOCAL;
                r.seed = r.deferred || r.segment ? shards[rid].seed : nullptr;
                if (r.seed && seed_from_cache[rid] && rid < g.seed_cache.size() && g.seed_cache[rid] == r.seed) {
                    // the value outlives this launch inside a variable: the next launch needs a seed of its own
                    hj_buffer_release(g.seed_cache[rid]);
                    g.seed_cache[rid] = nullptr;
                }
            }
            return r;
        };
        std::lock_guard<std::mutex> lock(g_trace_mu);
        if (outputs) {
            outputs->clear();
            for (uint32_t rid : g.outputs) {
                Var v;
                v.op.kind = OpKind::Buffer;
                v.ty = g.resource_descs[rid].ty;
                v.extent.n = g.resource_descs[rid].size;
                set_resource(v, resource_of(rid));
                outputs->push_back(g_trace.new_var_id(std::move(v)));
            }
        }
        for (size_t i = 0; i < nres; i++) {
            const GraphResource& gr = g.resources[i];
            if (!res[i]) continue;
            if (gr.kind == GraphResource::Internal) {
                if (Var* var = g_trace.get(gr.id)) {
                    BufferDesc have;
                    have.size = var->extent.n;
                    have.ty = var->ty;
                    if (have == g.resource_descs[i]) set_resource(*var, resource_of(i));
                }
            } else if (comm && shards[i].placement == HJ_RES_SHARDED && gr.kind == GraphResource::Captured) {
                // a captured deferred scan result a device op of this launch had to materialise
                if (Var* var = g_trace.get(gr.id))
                    if (var->data.kind == Resource::Buffer && var->data.buf == res[i] && shard_state(*var) != shards[i].deferred)
                        set_resource(*var, resource_of(i));
            }
        }
        // inputs a pass of this launch materialised in place
        for (size_t i = 0; comm && i < inputs.size() && i < g.inputs.size(); i++) {
            const uint32_t rid = g.inputs[i];
            if (Var* var = g_trace.get(inputs[i]))
                if (var->data.kind == Resource::Buffer && var->data.buf == res[rid] && shard_state(*var) != shards[rid].deferred)
                    set_resource(*var, resource_of(rid));
        }
    } catch (...) {
        release_all();
        throw;
    }
    release_all();
}

// `{:#?}` of Graph as stored in the reference's insta snapshots
// (hephaestus-jit/src/snapshots/hephaestus_jit__test__{select,conditional_scatter,conditionals}.snap)
std::string graph_debug_string(const Graph& g) {
    static const char* rops[] = {"Max", "Min", "Sum", "Prod", "Or", "And", "Xor"};
    std::ostringstream s;
    auto indent = [](const std::string& text, const std::string& pad) {
        std::string out;
        size_t pos = 0;
        while (pos < text.size()) {
            size_t nl = text.find('\n', pos);
            std::string line = text.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
            out += (pos ? pad : "") + line;
            if (nl == std::string::npos) break;
            out += "\n";
            pos = nl + 1;
        }
        return out;
    };
    s << "Graph {\n    passes: [";
    if (g.passes.empty()) s << "],\n";
    else {
        s << "\n";
        for (const Pass& p : g.passes) {
            s << "        Pass {\n            resources: [";
            if (p.resources.empty()) s << "],\n";
            else {
                s << "\n";
                for (uint32_t r : p.resources) s << "                ResourceId(\n                    " << r << ",\n                ),\n";
                s << "            ],\n";
            }
            if (p.size_buffer < 0) s << "            size_buffer: None,\n";
            else s << "            size_buffer: Some(\n                ResourceId(\n                    " << p.size_buffer << ",\n                ),\n            ),\n";
            if (p.is_kernel) {
                hj_ir view = p.ir.view();
                s << "            op: Kernel {\n                ir: " << indent(ir_debug_string(&view), "                ")
                  << ",\n                size: " << p.size << ",\n            },\n";
            } else {
                s << "            op: DeviceOp(\n";
                if (p.device_op.code == DOP_REDUCE)
                    s << "                ReduceOp(\n                    " << rops[p.device_op.arg % 7] << ",\n                ),\n";
                else if (p.device_op.code == DOP_PREFIX_SUM)
                    s << "                PrefixSum {\n                    inclusive: " << (p.device_op.arg ? "true" : "false") << ",\n                },\n";
                else
                    s << "                Compress,\n";
                s << "            ),\n";
            }
            s << "        },\n";
        }
        s << "    ],\n";
    }
    s << "    resource_descs: [";
    if (g.resource_descs.empty()) s << "],\n";
    else {
        s << "\n";
        for (const BufferDesc& d : g.resource_descs)
            s << "        BufferDesc(\n            BufferDesc {\n                size: " << d.size << ",\n                ty: " << type_debug(d.ty)
              << ",\n            },\n        ),\n";
        s << "    ],\n";
    }
    s << "    resources: [";
    if (g.resources.empty()) s << "],\n";
    else {
        s << "\n";
        for (const GraphResource& r : g.resources)
            s << "        " << (r.kind == GraphResource::Input ? "Input" : r.kind == GraphResource::Captured ? "Captured" : "Internal") << ",\n";
        s << "    ],\n";
    }
    auto id_list = [&](const char* name, const std::vector<uint32_t>& ids) {
        s << "    " << name << ": [";
        if (ids.empty()) { s << "],\n"; return; }
        s << "\n";
        for (uint32_t r : ids) s << "        ResourceId(\n            " << r << ",\n        ),\n";
        s << "    ],\n";
    };
    id_list("inputs", g.inputs);
    id_list("outputs", g.outputs);
    s << "}";
    return s.str();
}

}  // namespace tr
}  // namespace hj
