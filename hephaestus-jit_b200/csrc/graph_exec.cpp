// graph_exec.cpp — hj_execute_graph: the pass interpreter behind BackendDevice::execute_graph.
//
// Restates VulkanDevice::execute_graph (hephaestus-jit/src/backend/vulkan/mod.rs:151-383) for
// the CUDA backend: walks Graph::passes() in order; a Kernel pass fetches the NVRTC-compiled
// kernel for its IR from the per-device cache and launches it over the pass size (or the
// device-resident DynSize count); a DeviceOp pass dispatches to the hand-written kernel with
// the resource order the reference uses (Reduce / PrefixSum: [dst, src], mod.rs:259-280;
// Compress: [index_out, out_count, src], mod.rs:281-299).  Stream order replaces the
// reference's render-graph barriers (vulkan_core/graph.rs:362-570): every pass is enqueued on
// the device stream, so a pass sees all writes of the passes before it.
//
// Timing: when the caller asks for a report, each pass is bracketed by CUDA events and the
// call blocks until they are resolved — the reference always blocks on a fence and reads GPU
// timestamps (vulkan_core/device.rs:297-334, profiler.rs:56-131).  Without a report the call
// is asynchronous.
#include <algorithm>
#include <chrono>

#include "hj_internal.h"
#include "ir.h"

using namespace hj;

namespace hj {
// Captured CUDA graphs of hj_execute_graph_cached.  An instance belongs to one pass list (key) AND
// one set of buffer addresses: kernel nodes hold raw pointers.  `generation` is the scratch
// generation at capture (a reallocated scratch invalidates every instance).
struct GraphInstance {
    std::vector<void*> ptrs;
    cudaGraphExec_t exec = nullptr;
    uint64_t generation = 0, last_use = 0;
    uint32_t n_epochs = 0;     // look-back epochs one replay consumes (host-side wrap accounting)
    uint64_t n_launches = 0;   // kernel launches one replay stands for
    bool uncapturable = false;
    std::vector<uint32_t> deferred_out;  // sharded launches: the `deferred` flags the pass list leaves behind
};
struct GraphCache {
    std::unordered_map<uint64_t, std::vector<GraphInstance>> by_key;
    uint64_t captured = 0, replayed = 0, plain = 0, clock = 0;
};
}  // namespace hj

namespace {
// A kernel pass whose whole IR is `dst[keys[i]] += literal` (u32 / i32) — what
// `sized_literal(v, n).scatter_reduce(&hist, &keys, ReduceOp::Sum)` (trace.rs:1210-1233) compiles
// to — is a histogram: it goes to the privatised shared-memory kernel (scatter.cu) instead of
// one global atomic per key from the JIT kernel.  Returns false if the IR has any other shape.
struct HistMatch {
    uint32_t dst_slot, key_slot, ty;
    uint64_t literal;
};
bool match_histogram(const hj_ir* ir, HistMatch* m) {
    IRView v(ir);
    if (ir->n_buffers != 2 || ir->n_vars != 6) return false;
    int sr = -1;
    for (uint32_t i = 0; i < ir->n_vars; i++) {
        const uint32_t op = v.var(i).op;
        if (op == HJ_OP_SCATTER_REDUCE) {
            if (sr >= 0) return false;
            sr = (int)i;
        } else if (op != HJ_OP_BUFFER_REF && op != HJ_OP_LITERAL && op != HJ_OP_INDEX && op != HJ_OP_GATHER) {
            return false;
        }
    }
    if (sr < 0 || v.n_deps(sr) != 3 || v.var(sr).arg != HJ_REDUCE_SUM) return false;
    const uint32_t dst = v.dep(sr, 0), src = v.dep(sr, 1), idx = v.dep(sr, 2);
    if (v.var(dst).op != HJ_OP_BUFFER_REF || v.var(src).op != HJ_OP_LITERAL || v.var(idx).op != HJ_OP_GATHER) return false;
    if (v.n_deps(idx) != 2 || v.var(v.dep(idx, 0)).op != HJ_OP_BUFFER_REF || v.var(v.dep(idx, 1)).op != HJ_OP_INDEX) return false;
    const uint32_t kind = v.kind(v.var_type(src));
    if ((kind != HJ_U32 && kind != HJ_I32) || v.kind(v.var_type(idx)) != HJ_U32 || v.kind(v.var_type(dst)) != kind) return false;
    m->dst_slot = (uint32_t)v.var(dst).data;
    m->key_slot = (uint32_t)v.var(v.dep(idx, 0)).data;
    if (m->dst_slot == m->key_slot) return false;
    m->ty = kind;
    m->literal = v.var(src).data;
    return true;
}

// The scheduler zero-fills the whole index buffer in the kernel pass in front of every Compress
// (`index = sized_literal(0, n)`, trace.rs:1600-1601; usually fused with the kernel that computes
// the mask): 4 bytes written per element, as much as a compaction at p = 0.5 moves in total, only
// to be overwritten below `count`.  When that pass holds `index[i] = 0` as a plain top-level
// Scatter of a zero literal at the bare Index and nothing else touches the buffer, the store is
// taken out of the kernel (or the pass is skipped when nothing else is left in it) and the
// Compress pass zeroes index[count..n) itself: same bytes in the buffer afterwards, 4 * count
// bytes of HBM writes less and often one launch less.
struct PrefillMatch {
    uint32_t scatter_var, literal_ty;
    bool only_side_effect;  // the kernel does nothing else: skip the pass
};
bool match_index_prefill(const hj_ir* ir, uint32_t slot, PrefillMatch* m) {
    IRView v(ir);
    int found = -1, depth = 0;
    uint32_t other_effects = 0;
    for (uint32_t i = 0; i < ir->n_vars; i++) {
        const hj_ir_var& var = v.var(i);
        if (var.op == HJ_OP_LOOP_START || var.op == HJ_OP_IF_START) depth++;
        if (var.op == HJ_OP_LOOP_END || var.op == HJ_OP_IF_END) depth--;
        const bool effect = var.op == HJ_OP_SCATTER || var.op == HJ_OP_SCATTER_REDUCE || var.op == HJ_OP_SCATTER_ATOMIC ||
                            var.op == HJ_OP_ATOMIC_INC;
        bool is_prefill = false;
        if (var.op == HJ_OP_SCATTER && depth == 0 && v.n_deps(i) == 3 && found < 0) {
            const hj_ir_var &dst = v.var(v.dep(i, 0)), &src = v.var(v.dep(i, 1)), &idx = v.var(v.dep(i, 2));
            is_prefill = dst.op == HJ_OP_BUFFER_REF && dst.data == slot && src.op == HJ_OP_LITERAL && src.data == 0 &&
                         v.kind(src.ty) == HJ_U32 && idx.op == HJ_OP_INDEX;
        }
        if (is_prefill) {
            found = (int)i;
            m->literal_ty = v.var(v.dep(i, 1)).ty;
        } else if (effect) {
            other_effects++;
        }
    }
    if (found < 0) return false;
    for (uint32_t i = 0; i < ir->n_vars; i++) {  // nothing else may read or write the buffer, or name the store
        if ((int)i == found) continue;
        for (uint32_t k = 0; k < v.n_deps(i); k++) {
            const uint32_t d = v.dep(i, k);
            if ((int)d == found) return false;
            if (v.var(d).op == HJ_OP_BUFFER_REF && v.var(d).data == slot) return false;
        }
    }
    m->scatter_var = (uint32_t)found;
    m->only_side_effect = other_effects == 0;
    return true;
}
}  // namespace

// What execute_graph decides for a kernel pass in front of a Compress pass, exposed so that the
// decision can be tested without a GPU: 1 if buffer slot `slot` of `ir` is only ever written by a
// top-level, unconditional `Scatter(BufferRef(slot), Literal 0u32, Index)`.
extern "C" int32_t hj_ir_index_zero_fill(const hj_ir* ir, uint32_t slot, uint32_t* scatter_var, int32_t* nothing_else) {
    if (!ir || !validate_ir(ir).empty()) return 0;
    PrefillMatch m;
    if (!match_index_prefill(ir, slot, &m)) return 0;
    if (scatter_var) *scatter_var = m.scatter_var;
    if (nothing_else) *nothing_else = m.only_side_effect ? 1 : 0;
    return 1;
}

namespace {

// CUDA events of a timed launch; destroyed on every path out of execute_passes
struct EventList {
    std::vector<cudaEvent_t> ev;
    ~EventList() {
        for (cudaEvent_t e : ev)
            if (e) cudaEventDestroy(e);
    }
    hj_status create(size_t n) {
        ev.assign(n, nullptr);
        for (auto& e : ev) HJ_CUDA(cudaEventCreate(&e));
        return HJ_OK;
    }
};

// Sharded execution (hj_execute_graph_sharded): which rank this is and where every resource lives
struct ShardCtx {
    hj_comm* comm;
    hj_shard_desc* shards;
    int rank, world;
};

void shard_bounds(uint64_t n, int world, int rank, uint64_t* start, uint64_t* end) {
    const uint64_t base = n / (uint64_t)world, rem = n % (uint64_t)world;
    *start = (uint64_t)rank * base + std::min<uint64_t>((uint64_t)rank, rem);
    *end = *start + base + ((uint64_t)rank < rem ? 1 : 0);
}

// An owned copy of an IR whose Gathers from the `deferred` slots add the slot's seed (a one-element
// buffer bound to a NEW slot behind the original ones): Gather(buf, i) becomes
// Gather(buf, i) + Gather(seed, 0).  This is how the consumers of a sharded scan apply the
// cross-GPU offset at load instead of the scan paying a second pass over its output.
struct SeededIR {
    std::vector<hj_ir_var> vars;
    std::vector<uint32_t> deps;
    std::vector<hj_type_desc> types;
    hj_ir view;
    std::vector<uint32_t> seed_of_slot;  // extra slot k (n_buffers + k of the original) carries the seed of this original slot
};
void build_seeded_ir(const hj_ir* ir, const std::vector<bool>& deferred, SeededIR* out) {
    IRView v(ir);
    out->types.assign(ir->types, ir->types + ir->n_types);
    uint32_t u32_ty = ir->n_types;
    for (uint32_t t = 0; t < ir->n_types; t++)
        if (ir->types[t].kind == HJ_U32) { u32_ty = t; break; }
    if (u32_ty == ir->n_types) {
        hj_type_desc d = {};
        d.kind = HJ_U32;
        out->types.push_back(d);
    }
    std::vector<int64_t> extra_slot(ir->n_buffers, -1);
    std::vector<uint32_t> remap(ir->n_vars);
    auto push = [&](uint32_t ty, uint32_t op, uint32_t arg, uint64_t data, std::initializer_list<uint32_t> d) {
        hj_ir_var nv = {};
        nv.ty = ty;
        nv.op = op;
        nv.arg = arg;
        nv.data = data;
        nv.dep_start = (uint32_t)out->deps.size();
        out->deps.insert(out->deps.end(), d.begin(), d.end());
        nv.dep_end = (uint32_t)out->deps.size();
        out->vars.push_back(nv);
        return (uint32_t)out->vars.size() - 1;
    };
    for (uint32_t i = 0; i < ir->n_vars; i++) {
        hj_ir_var nv = v.var(i);
        const uint32_t d0 = nv.dep_start, d1 = nv.dep_end;
        nv.dep_start = (uint32_t)out->deps.size();
        for (uint32_t k = d0; k < d1; k++) out->deps.push_back(remap[ir->deps[k]]);
        nv.dep_end = (uint32_t)out->deps.size();
        out->vars.push_back(nv);
        uint32_t id = (uint32_t)out->vars.size() - 1;
        if (nv.op == HJ_OP_GATHER && d1 - d0 == 2) {
            const hj_ir_var& buf = v.var(ir->deps[d0]);
            if (buf.op == HJ_OP_BUFFER_REF && buf.data < ir->n_buffers && deferred[buf.data]) {
                if (extra_slot[buf.data] < 0) {
                    extra_slot[buf.data] = (int64_t)ir->n_buffers + (int64_t)out->seed_of_slot.size();
                    out->seed_of_slot.push_back((uint32_t)buf.data);
                }
                const uint32_t ref = push(buf.ty, HJ_OP_BUFFER_REF, 0, (uint64_t)extra_slot[buf.data], {});
                const uint32_t zero = push(u32_ty, HJ_OP_LITERAL, 0, 0, {});
                const uint32_t seed = push(nv.ty, HJ_OP_GATHER, 0, 0, {ref, zero});
                id = push(nv.ty, HJ_OP_BOP, HJ_BOP_ADD, 0, {id, seed});
            }
        }
        remap[i] = id;
    }
    out->view = *ir;
    out->view.vars = out->vars.data();
    out->view.n_vars = (uint32_t)out->vars.size();
    out->view.deps = out->deps.data();
    out->view.n_deps = (uint32_t)out->deps.size();
    out->view.types = out->types.data();
    out->view.n_types = (uint32_t)out->types.size();
    out->view.n_buffers = ir->n_buffers + (uint32_t)out->seed_of_slot.size();
}

bool integer_kind(uint32_t k) { return k >= HJ_I8 && k <= HJ_U64; }
bool is_segment_state(uint32_t st) { return st == HJ_SHARD_SEGMENT || st == HJ_SHARD_SEGMENT_LOCAL; }

// buf[0 .. n_local) += seed: a deferred scan result becomes an ordinary shard
hj_status materialise(hj_device* dev, const ShardCtx* sc, uint32_t rid, hj_buffer* buf, const hj_buffer_desc& d) {
    hj_shard_desc& sd = sc->shards[rid];
    if (is_segment_state(sd.deferred))
        return fail(HJ_ERR_UNSUPPORTED, "resource %u is a per-rank compacted segment: only DynSize kernels sized by its count "
                    "run over it (SURVEY 8e; re-balance it into a block-sharded array first: hj_sharded_rebalance)", rid);
    if (sd.deferred != HJ_SHARD_DEFERRED) return HJ_OK;
    HJ_REQUIRE(sd.seed, "resource %u is marked deferred but carries no seed buffer", rid);
    uint64_t s0, s1;
    shard_bounds(d.size, sc->world, sc->rank, &s0, &s1);
    HJ_TRY(hj_apply_seed(dev, (hj_type_kind)d.ty, (size_t)(s1 - s0), buf, sd.seed));
    sd.deferred = HJ_SHARD_PLAIN;
    return HJ_OK;
}

hj_status execute_passes(hj_device* dev, const ShardCtx* sc, const hj_pass* passes, uint32_t n_passes,
                         hj_buffer* const* env, const hj_buffer_desc* descs, uint32_t n_resources, hj_report* report) {
    auto cpu_start = std::chrono::steady_clock::now();
    const bool timed = report && report->passes && report->passes_capacity >= n_passes;
    EventList events;
    std::vector<cudaEvent_t>& ev = events.ev;
    if (timed) {
        cudaSetDevice(dev->ordinal);
        HJ_TRY(events.create((size_t)n_passes + 1));
        HJ_CUDA(cudaEventRecord(ev[0], dev->stream));
    }
    if (dev->async_live.load(std::memory_order_acquire) != 0) {
        // Some buffer is still arriving chunk by chunk (hj_buffer_create_from_host_async).  A single kernel pass
        // over the bare Index joins the pipeline: it is launched per chunk behind the upload events and its
        // outputs inherit the schedule (jit.cpp: kernel_launch_streamed).  Everything else waits for the
        // whole upload: the device stream is ordered behind it.
        if (!sc && !timed && n_passes == 1 && passes[0].kind == HJ_PASS_KERNEL && passes[0].size_buffer < 0 && passes[0].ir &&
            passes[0].resources && validate_ir(passes[0].ir).empty() && passes[0].ir->n_buffers == passes[0].n_resources) {
            const hj_pass& p = passes[0];
            std::vector<hj_buffer*> bufs(p.n_resources);
            bool bound = true;
            for (uint32_t b = 0; b < p.n_resources; b++) {
                bound = bound && p.resources[b] < n_resources && env[p.resources[b]];
                if (bound) bufs[b] = env[p.resources[b]];
            }
            if (bound) {
                hj_kernel* k = nullptr;
                HJ_TRY(hj_kernel_get(dev, p.ir, &k));
                bool done = false;
                hj_status s = kernel_launch_streamed(dev, k, p.size, bufs.data(), (uint32_t)bufs.size(), &done);
                hj_kernel_release(k);
                HJ_TRY(s);
                if (done) {
                    if (report) {
                        report->n_passes = n_passes;
                        report->cpu_duration_us =
                            std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - cpu_start).count();
                    }
                    return HJ_OK;
                }
            }
        }
        if (!sc && !timed && n_passes == 1 && passes[0].kind == HJ_PASS_PREFIX_SUM && passes[0].n_resources >= 2 &&
            passes[0].resources && passes[0].resources[0] < n_resources && passes[0].resources[1] < n_resources &&
            env[passes[0].resources[0]] && env[passes[0].resources[1]]) {
            const hj_pass& p = passes[0];
            const uint32_t rd = p.resources[0], rs = p.resources[1];
            bool done = false;
            HJ_TRY(prefix_sum_arriving(dev, (hj_type_kind)descs[rd].ty, descs[rs].size, p.arg != 0 || getenv("HJ_REF_COMPAT"),
                                       env[rs], env[rd], &done));
            if (done) {
                if (report) {
                    report->n_passes = n_passes;
                    report->cpu_duration_us =
                        std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - cpu_start).count();
                }
                return HJ_OK;
            }
        }
        DeviceGuard g(dev);
        for (uint32_t i = 0; i < n_resources; i++) settle_locked(env[i]);
    }
    auto res = [&](const hj_pass& p, uint32_t k, hj_buffer** out, const hj_buffer_desc** desc) -> hj_status {
        HJ_REQUIRE(k < p.n_resources, "pass references resource slot %u but has %u", k, p.n_resources);
        uint32_t id = p.resources[k];
        HJ_REQUIRE(id < n_resources, "ResourceId %u out of range (%u resources)", id, n_resources);
        HJ_REQUIRE(env[id], "resource %u has been left empty (graph.rs: UninitializedResourve)", id);
        *out = env[id];
        if (desc) *desc = &descs[id];
        return HJ_OK;
    };
    auto sharded = [&](uint32_t rid) { return sc && sc->shards[rid].placement == HJ_RES_SHARDED; };
    // this rank's block of a sharded resource
    auto local_of = [&](uint32_t rid, uint64_t* start, uint64_t* count) -> hj_status {
        uint64_t s0, s1;
        shard_bounds(descs[rid].size, sc->world, sc->rank, &s0, &s1);
        HJ_REQUIRE(s1 > s0, "resource %u: %llu elements cannot be sharded over %d ranks", rid,
                   (unsigned long long)descs[rid].size, sc->world);
        HJ_REQUIRE(descs[rid].size <= 0xffffffffull, "resource %u: global size does not fit the u32 index type", rid);
        *start = s0;
        *count = s1 - s0;
        return HJ_OK;
    };

    // A DynSize kernel over sharded resources: the wavefront step behind a sharded Compress
    // (jit/test.rs:1020-1062: indices = mask.compress_dyn(); a.gather(indices) ... scatter(a, indices)).
    // Every rank runs it over ITS segment of the compacted indices, sized by its own count (the segment's
    // seed).  The indices are global and fall into the rank's block of every array of the mask's extent,
    // so such arrays are bound with their base moved back by the block's start and addressed in place.
    auto segment_kernel = [&](const hj_pass& p, uint32_t i) -> hj_status {
        std::vector<hj_buffer*> bufs(p.n_resources);
        for (uint32_t b = 0; b < p.n_resources; b++) HJ_TRY(res(p, b, &bufs[b], nullptr));
        std::vector<SlotAccess> slot;
        analyse_slot_access(p.ir, &slot);
        std::vector<SegmentAccess> acc;
        std::string why = "none of its resources is the index segment of a sharded Compress that carries a seed buffer";
        uint32_t seg_slot = p.n_resources;
        std::vector<uint64_t> shift(p.n_resources, 0);
        std::vector<uint32_t> born;  // slots written at Index for the first time: aligned with the segment from now on
        for (uint32_t c = 0; c < p.n_resources && seg_slot == p.n_resources; c++) {
            const uint32_t crid = p.resources[c];
            if (!sharded(crid) || !is_segment_state(sc->shards[crid].deferred) || !sc->shards[crid].seed ||
                sc->shards[crid].seed->bytes < 8 || descs[crid].ty != HJ_U32 || descs[crid].size != p.size)
                continue;
            // entries of a LOCAL segment are positions in the rank's part of its parent sequence: what is reached
            // through them must itself be a segment (bound as it is), not a block of a sharded array
            const bool local_positions = sc->shards[crid].deferred == HJ_SHARD_SEGMENT_LOCAL;
            if (!analyse_segment_access(p.ir, c, &acc, &why)) continue;
            bool ok = true;
            std::fill(shift.begin(), shift.end(), 0);
            born.clear();
            for (uint32_t b = 0; b < p.n_resources && ok; b++) {
                const uint32_t rid = p.resources[b];
                const SegmentAccess& a = acc[b];
                if (!a.read && !a.written) continue;
                char msg[160] = "";
                if (!sharded(rid)) {
                    if (a.written) snprintf(msg, sizeof(msg), "resource %u, a replica, is written", rid);
                    else if (slot[b].any_index) snprintf(msg, sizeof(msg), "replica %u is read at Index, the position in the rank's segment", rid);
                } else if (is_segment_state(sc->shards[rid].deferred)) {
                    if (a.index_only) {
                        if (a.written && b != c) born.push_back(b);  // rewritten over THIS segment: its count is this one's now
                    } else if (!(local_positions && a.through_segment && descs[rid].size == descs[crid].size)) {
                        snprintf(msg, sizeof(msg), "segment %u is addressed through a computed index", rid);
                    }
                } else if (a.through_segment) {
                    if (local_positions)
                        snprintf(msg, sizeof(msg), "resource %u, a block of a sharded array, is addressed through positions in a segment", rid);
                    else if (descs[rid].size != descs[crid].size)
                        snprintf(msg, sizeof(msg), "sharded resource %u has another extent than the compacted mask", rid);
                } else if (a.index_only && !a.read) {
                    if (descs[rid].size != p.size) snprintf(msg, sizeof(msg), "resource %u written at Index has another extent", rid);
                    else if (!sc->shards[rid].seed) snprintf(msg, sizeof(msg), "resource %u written at Index carries no seed buffer for its count", rid);
                    else born.push_back(b);
                } else {
                    snprintf(msg, sizeof(msg), "sharded resource %u is addressed neither through the index segment nor as a segment", rid);
                }
                if (msg[0]) {
                    why = msg;
                    ok = false;
                }
            }
            if (ok) seg_slot = c;
        }
        if (seg_slot == p.n_resources)
            return fail(HJ_ERR_UNSUPPORTED, "kernel pass %u: a DynSize kernel over sharded resources runs per compacted segment, but %s "
                        "(SURVEY 8e; otherwise re-balance the compacted sequence first: hj_sharded_rebalance)", i, why.c_str());
        const uint32_t seg_rid = p.resources[seg_slot];
        uint64_t s0 = 0, cnt = 0;
        HJ_TRY(local_of(seg_rid, &s0, &cnt));
        for (uint32_t b = 0; b < p.n_resources; b++) {
            const uint32_t rid = p.resources[b];
            if (!sharded(rid) || is_segment_state(sc->shards[rid].deferred) || !(acc[b].read || acc[b].written)) continue;
            if (acc[b].through_segment) {
                HJ_TRY(materialise(dev, sc, rid, bufs[b], descs[rid]));  // a deferred scan result is written / gathered in place
                HJ_REQUIRE(cnt * descs[rid].elem_bytes <= bufs[b]->bytes, "kernel pass %u: resource %u is smaller than the rank's block", i, rid);
                shift[b] = s0 * descs[rid].elem_bytes;
            }
        }
        hj_kernel* k = nullptr;
        HJ_TRY(hj_kernel_get(dev, p.ir, &k));
        // index_base = 0xffffffff: KernelOp::Index = seed[1] + local position, the position in the GLOBAL compacted
        // sequence (codegen.cpp: HJ_SIZE reads it behind the count)
        hj_status s = kernel_launch_shifted(dev, k, (size_t)cnt, sc->shards[seg_rid].seed, bufs.data(), (uint32_t)bufs.size(),
                                            0xffffffffu, shift.data());
        hj_kernel_release(k);
        HJ_TRY(s);
        for (uint32_t b : born) {
            hj_shard_desc& sd = sc->shards[p.resources[b]];
            HJ_REQUIRE(sd.seed->bytes >= 8, "kernel pass %u: the seed buffer of resource %u is smaller than 8 bytes", i, p.resources[b]);
            HJ_CUDA(cudaMemcpyAsync(sd.seed->ptr, sc->shards[seg_rid].seed->ptr, 8, cudaMemcpyDeviceToDevice, dev->stream));
            sd.deferred = HJ_SHARD_SEGMENT;
        }
        return HJ_OK;
    };

    int64_t zero_tail_for = -1;  // index of the Compress pass that has to zero index[count..n) itself
    for (uint32_t i = 0; i < n_passes; i++) {
        const hj_pass& p = passes[i];
        char name[64];
        hj_buffer* size_buf = nullptr;
        if (p.size_buffer >= 0) {
            HJ_REQUIRE((uint32_t)p.size_buffer < n_resources && env[p.size_buffer], "size buffer resource missing");
            size_buf = env[p.size_buffer];
        }
        for (uint32_t b = 0; b < p.n_resources; b++)
            HJ_REQUIRE(p.resources && p.resources[b] < n_resources, "pass %u: ResourceId out of range", i);
        switch (p.kind) {
        case HJ_PASS_KERNEL: {
            HJ_REQUIRE(p.ir, "kernel pass %u without IR", i);
            {   // an FFI caller may hand over anything: the matchers below index vars / deps by id
                const std::string verr = validate_ir(p.ir);
                if (!verr.empty()) return fail(HJ_ERR_INVALID, "kernel pass %u: IR rejected: %s", i, verr.c_str());
                HJ_REQUIRE(p.ir->n_buffers == p.n_resources, "kernel pass %u: IR names %u buffers, the pass binds %u", i,
                           p.ir->n_buffers, p.n_resources);
            }
            bool pass_sharded = false;
            for (uint32_t b = 0; b < p.n_resources; b++) pass_sharded = pass_sharded || sharded(p.resources[b]);
            if (pass_sharded && size_buf) {
                snprintf(name, sizeof(name), "JIT Kernel %u [segment of %llu]", i, (unsigned long long)p.size);
                HJ_TRY(segment_kernel(p, i));
                break;
            }
            HistMatch hm;
            if (!size_buf && p.n_resources == 2 && p.size >= (1u << 20) && match_histogram(p.ir, &hm) &&
                (!pass_sharded || (sharded(p.resources[hm.key_slot]) && !sharded(p.resources[hm.dst_slot])))) {
                snprintf(name, sizeof(name), "Histogram %u [%llu]", i, (unsigned long long)p.size);
                hj_buffer *dst = nullptr, *keys = nullptr;
                const hj_buffer_desc* ddst = nullptr;
                HJ_TRY(res(p, hm.dst_slot, &dst, &ddst));
                HJ_TRY(res(p, hm.key_slot, &keys, nullptr));
                if (pass_sharded) {
                    // privatised per GPU, then combined: the result is the sum over ranks of their copies of
                    // dst, so only rank 0 keeps what dst held before
                    uint64_t k0, kn;
                    HJ_TRY(local_of(p.resources[hm.key_slot], &k0, &kn));
                    HJ_TRY(materialise(dev, sc, p.resources[hm.key_slot], keys, descs[p.resources[hm.key_slot]]));
                    if (sc->rank != 0) HJ_CUDA(cudaMemsetAsync(dst->ptr, 0, ddst->size * 4, dev->stream));
                    HJ_TRY(hj_sharded_scatter_reduce(sc->comm, HJ_REDUCE_SUM, (hj_type_kind)hm.ty, (size_t)kn, keys, nullptr,
                                                     hm.literal, dst, ddst->size));
                } else {
                    HJ_TRY(hj_scatter_reduce(dev, HJ_REDUCE_SUM, (hj_type_kind)hm.ty, p.size, keys, nullptr, hm.literal, dst,
                                             ddst->size));
                }
                break;
            }
            // zero-fill of the index buffer of the Compress pass that follows (see match_index_prefill)
            const hj_ir* ir = p.ir;
            hj_ir ir_without_fill;
            std::vector<hj_ir_var> vars_without_fill;
            PrefillMatch pm;
            static const bool no_elision = getenv("HJ_NO_PREFILL_ELISION") != nullptr;
            // `count = sized_literal(0, 1)` (trace.rs:1599) gets a one-element kernel of its own; the
            // Compress pass always writes out_count[0], so a pass that does nothing else is dropped
            if (!no_elision && !size_buf && p.size == 1 && p.n_resources == 1) {
                bool dropped = false;
                for (uint32_t j = i + 1; j < n_passes && j <= i + 2 && !dropped; j++) {
                    const hj_pass& c = passes[j];
                    if (c.kind == HJ_PASS_COMPRESS && c.n_resources >= 3 && c.resources[1] == p.resources[0] &&
                        c.resources[0] != p.resources[0] && c.resources[2] != p.resources[0]) {
                        dropped = match_index_prefill(p.ir, 0, &pm) && pm.only_side_effect;
                        break;
                    }
                    bool touches = c.size_buffer >= 0 && (uint32_t)c.size_buffer == p.resources[0];
                    for (uint32_t b = 0; b < c.n_resources; b++) touches = touches || c.resources[b] == p.resources[0];
                    if (touches) break;
                }
                if (dropped) {
                    snprintf(name, sizeof(name), "JIT Kernel %u [1] (count zero-fill left to Compress)", i);
                    break;
                }
            }
            if (!no_elision && !size_buf && i + 1 < n_passes && passes[i + 1].kind == HJ_PASS_COMPRESS &&
                passes[i + 1].n_resources >= 3) {
                const hj_pass& c = passes[i + 1];
                const uint32_t index_res = c.resources[0], mask_res = c.resources[2];
                uint32_t slot = p.n_resources;
                for (uint32_t b = 0; b < p.n_resources; b++)
                    if (p.resources[b] == index_res) slot = slot == p.n_resources ? b : p.n_resources + 1;
                if (slot < p.n_resources && index_res < n_resources && mask_res < n_resources && env[index_res] &&
                    ((uintptr_t)env[index_res]->ptr & 15u) == 0 && descs[index_res].size == p.size &&
                    descs[mask_res].size == p.size && sharded(index_res) == sharded(mask_res) &&
                    match_index_prefill(p.ir, slot, &pm)) {
                    zero_tail_for = (int64_t)i + 1;
                    if (pm.only_side_effect) {
                        snprintf(name, sizeof(name), "JIT Kernel %u [%llu] (index zero-fill left to Compress)", i,
                                 (unsigned long long)p.size);
                        break;
                    }
                    vars_without_fill.assign(p.ir->vars, p.ir->vars + p.ir->n_vars);
                    hj_ir_var& dead = vars_without_fill[pm.scatter_var];  // becomes an unused `u32 r = 0`
                    dead.ty = pm.literal_ty;
                    dead.op = HJ_OP_LITERAL;
                    dead.arg = 0;
                    dead.dep_start = dead.dep_end = 0;
                    dead.data = 0;
                    ir_without_fill = *p.ir;
                    ir_without_fill.vars = vars_without_fill.data();
                    ir = &ir_without_fill;
                }
            }
            snprintf(name, sizeof(name), "JIT Kernel %u [%llu]", i, (unsigned long long)p.size);
            std::vector<hj_buffer*> bufs(p.n_resources);
            for (uint32_t b = 0; b < p.n_resources; b++) HJ_TRY(res(p, b, &bufs[b], nullptr));
            size_t launch_size = p.size;
            uint32_t index_base = 0;
            SeededIR seeded;
            struct Views {  // non-owning views of replicas, released on every path out of this pass
                std::vector<hj_buffer*> v;
                ~Views() { for (hj_buffer* b : v) hj_buffer_release(b); }
                void push_back(hj_buffer* b) { v.push_back(b); }
            } views;
            if (pass_sharded) {
                // every sharded resource is one contiguous block of a global array of p.size elements,
                // addressed by the bare Index; everything else is a replica and may only be read
                std::vector<SlotAccess> access;
                analyse_slot_access(ir, &access);
                uint64_t s0 = 0, cnt = 0;
                std::vector<uint32_t> replica_views;
                std::vector<bool> add_seed(p.n_resources, false);
                bool any_seed = false;
                for (uint32_t b = 0; b < p.n_resources; b++) {
                    const uint32_t rid = p.resources[b];
                    if (!sharded(rid)) {
                        if (access[b].written)
                            return fail(HJ_ERR_UNSUPPORTED, "kernel pass %u writes resource %u, a replica, from a kernel over "
                                        "sharded data: scatters into a sharded-over destination are replicas only (SURVEY 8e)", i, rid);
                        // the bare Index addresses the rank's block (codegen.cpp: addr_of), a replica holds the
                        // WHOLE array: a replica read at Index is bound as the view that starts at the block
                        if (access[b].any_index) {
                            if (!access[b].index_only || descs[rid].size != p.size)
                                return fail(HJ_ERR_UNSUPPORTED, "kernel pass %u reads replica %u both at Index and through computed "
                                            "indices (or with another extent) from a kernel over sharded data", i, rid);
                            replica_views.push_back(b);
                        }
                        continue;
                    }
                    if (descs[rid].size != p.size || !access[b].index_only)
                        return fail(HJ_ERR_UNSUPPORTED, "kernel pass %u addresses sharded resource %u through a computed index "
                                    "or with another extent: only Index-addressed access shards (SURVEY 8e)", i, rid);
                    HJ_TRY(local_of(rid, &s0, &cnt));
                    if (is_segment_state(sc->shards[rid].deferred)) {
                        // overwriting a segment (the index zero-fill in front of the next Compress) makes it a block again
                        if (access[b].read) HJ_TRY(materialise(dev, sc, rid, bufs[b], descs[rid]));  // fails, naming the resource
                        sc->shards[rid].deferred = HJ_SHARD_PLAIN;
                    }
                    if (sc->shards[rid].deferred) {
                        if (access[b].written || access[b].cond_gather) HJ_TRY(materialise(dev, sc, rid, bufs[b], descs[rid]));
                        else add_seed[b] = any_seed = true;
                    }
                }
                launch_size = (size_t)cnt;
                index_base = (uint32_t)s0;
                for (uint32_t b : replica_views) {
                    hj_buffer* view = nullptr;
                    const size_t es = descs[p.resources[b]].elem_bytes;
                    HJ_REQUIRE((s0 + cnt) * es <= bufs[b]->bytes, "kernel pass %u: replica %u is smaller than the pass extent", i, p.resources[b]);
                    HJ_TRY(hj_buffer_wrap(dev, (char*)bufs[b]->ptr + s0 * es, (size_t)cnt * es, &view));
                    views.push_back(view);
                    bufs[b] = view;
                }
                if (any_seed) {
                    build_seeded_ir(ir, add_seed, &seeded);
                    ir = &seeded.view;
                    for (uint32_t slot : seeded.seed_of_slot) bufs.push_back(sc->shards[p.resources[slot]].seed);
                    for (hj_buffer* sb : bufs) HJ_REQUIRE(sb, "kernel pass %u: a deferred resource carries no seed buffer", i);
                }
            }
            hj_kernel* k = nullptr;
            HJ_TRY(hj_kernel_get(dev, ir, &k));
            hj_status s = hj_kernel_launch(dev, k, launch_size, size_buf, bufs.data(), (uint32_t)bufs.size(), index_base);
            hj_kernel_release(k);
            HJ_TRY(s);
            break;
        }
        case HJ_PASS_REDUCE: {
            snprintf(name, sizeof(name), "Reduce");
            hj_buffer *dst = nullptr, *src = nullptr;
            const hj_buffer_desc *ddst = nullptr, *dsrc = nullptr;
            HJ_TRY(res(p, 0, &dst, &ddst));
            HJ_TRY(res(p, 1, &src, &dsrc));
            if (sharded(p.resources[1])) {
                uint64_t s0, cnt;
                HJ_TRY(local_of(p.resources[1], &s0, &cnt));
                HJ_TRY(materialise(dev, sc, p.resources[1], src, *dsrc));
                HJ_TRY(hj_sharded_reduce(sc->comm, (hj_reduce_op)p.arg, (hj_type_kind)ddst->ty, (size_t)cnt, src, dst));
            } else {
                HJ_TRY(hj_reduce(dev, (hj_reduce_op)p.arg, (hj_type_kind)ddst->ty, dsrc->size, src, dst));
            }
            break;
        }
        case HJ_PASS_PREFIX_SUM: {
            snprintf(name, sizeof(name), "Prefix Sum Large");
            hj_buffer *dst = nullptr, *src = nullptr;
            const hj_buffer_desc *ddst = nullptr, *dsrc = nullptr;
            HJ_TRY(res(p, 0, &dst, &ddst));
            HJ_TRY(res(p, 1, &src, &dsrc));
            if (sharded(p.resources[1])) {
                HJ_REQUIRE(sharded(p.resources[0]), "prefix-sum pass %u: the scan of a sharded resource is sharded", i);
                uint64_t s0, cnt;
                HJ_TRY(local_of(p.resources[1], &s0, &cnt));
                HJ_TRY(materialise(dev, sc, p.resources[1], src, *dsrc));
                hj_shard_desc& sd = sc->shards[p.resources[0]];
                static const bool no_defer = getenv("HJ_NO_DEFERRED_SEED") != nullptr;
                if (sd.seed && integer_kind(ddst->ty) && sc->world > 1 && !no_defer) {
                    // 8 bytes per element: local scan, the rank's offset stays in `seed` for the consumers
                    HJ_TRY(hj_sharded_prefix_sum_deferred(sc->comm, (hj_type_kind)ddst->ty, (size_t)cnt, (int32_t)p.arg, src, dst,
                                                          sd.seed));
                    sd.deferred = 1;
                    snprintf(name, sizeof(name), "Prefix Sum Large (deferred seed)");
                } else {
                    HJ_TRY(hj_sharded_prefix_sum(sc->comm, (hj_type_kind)ddst->ty, (size_t)cnt, (int32_t)p.arg, src, dst));
                    sd.deferred = 0;
                }
            } else {
                HJ_TRY(hj_prefix_sum(dev, (hj_type_kind)ddst->ty, dsrc->size, (int32_t)p.arg, src, dst, nullptr));
            }
            break;
        }
        case HJ_PASS_COMPRESS: {
            snprintf(name, sizeof(name), "Compress Large");
            hj_buffer *index_out = nullptr, *out_count = nullptr, *src = nullptr;
            const hj_buffer_desc* dsrc = nullptr;
            HJ_TRY(res(p, 0, &index_out, nullptr));
            HJ_TRY(res(p, 1, &out_count, nullptr));
            HJ_TRY(res(p, 2, &src, &dsrc));
            const bool zt = zero_tail_for == (int64_t)i;  // the pass in front left the zero-fill to us
            if (sharded(p.resources[2])) {
                HJ_REQUIRE(sharded(p.resources[0]) && !sharded(p.resources[1]),
                           "compress pass %u: the index segment of a sharded mask is sharded, the count a replica", i);
                uint64_t s0, cnt;
                HJ_TRY(local_of(p.resources[2], &s0, &cnt));
                HJ_REQUIRE(cnt <= src->bytes && cnt * 4 <= index_out->bytes && out_count->bytes >= 4,
                           "compress pass %u: buffer sizes do not match the %llu-element shard", i, (unsigned long long)cnt);
                // the rank's own count stays beside the segment: it sizes the DynSize kernels that run over it
                hj_shard_desc& seg = sc->shards[p.resources[0]];
                const hj_shard_desc& msk = sc->shards[p.resources[2]];
                hj_buffer* seg_seed = seg.seed && seg.seed->bytes >= 8 ? seg.seed : nullptr;
                if (size_buf) {
                    // nested compaction (jit/test.rs:976-1019): the mask is aligned with a segment and only its first
                    // seed[0] elements on this rank count.  The result holds positions in the RANK's part of the
                    // parent sequence (index_base 0): they address the rank's segments in place.
                    HJ_REQUIRE(msk.deferred == HJ_SHARD_SEGMENT && msk.seed && msk.seed->bytes >= 8,
                               "compress pass %u: a DynSize Compress over sharded data needs a mask that is aligned with a "
                               "compacted segment and carries its count", i);
                    snprintf(name, sizeof(name), "Compress Large (segment)");
                    HJ_TRY(sharded_compress_pass(sc->comm, (size_t)cnt, 0u, src, index_out, out_count, zt, seg_seed, msk.seed));
                    seg.deferred = seg_seed ? HJ_SHARD_SEGMENT_LOCAL : HJ_SHARD_PLAIN;
                } else {
                    HJ_REQUIRE(!is_segment_state(msk.deferred),
                               "compress pass %u: the mask is a per-rank segment, not a block of a sharded array", i);
                    HJ_TRY(sharded_compress_pass(sc->comm, (size_t)cnt, (uint32_t)s0, src, index_out, out_count, zt, seg_seed));
                    seg.deferred = seg_seed ? HJ_SHARD_SEGMENT : HJ_SHARD_PLAIN;
                }
            } else if (zt) {
                const size_t n = dsrc->size;
                HJ_REQUIRE(n >= 1 && n <= src->bytes && n * 4 <= index_out->bytes && out_count->bytes >= 4 &&
                               (!size_buf || size_buf->bytes >= 4),
                           "compress pass %u: buffer sizes do not match %zu elements", i, n);
                DeviceGuard g(dev);
                HJ_TRY(launch_compress(dev, n, size_buf ? (const uint32_t*)size_buf->ptr : nullptr, (uint32_t*)out_count->ptr,
                                       (const uint8_t*)src->ptr, (uint32_t*)index_out->ptr, 0, true));
            } else {
                HJ_TRY(hj_compress(dev, dsrc->size, size_buf, out_count, src, index_out, 0));
            }
            break;
        }
        default:
            return fail(HJ_ERR_UNSUPPORTED, "pass %u: device op %u is out of scope for the B200 backend "
                        "(MatMul / FusedMlp / textures / acceleration structures)", i, p.kind);
        }
        if (timed) {
            HJ_CUDA(cudaEventRecord(ev[i + 1], dev->stream));
            snprintf(report->passes[i].name, sizeof(report->passes[i].name), "%s", name);
        }
    }
    if (timed) {
        HJ_CUDA(cudaEventSynchronize(ev[n_passes]));
        for (uint32_t i = 0; i < n_passes; i++) {
            float start_ms = 0.f, dur_ms = 0.f;
            cudaEventElapsedTime(&start_ms, ev[0], ev[i]);
            cudaEventElapsedTime(&dur_ms, ev[i], ev[i + 1]);
            report->passes[i].start_us = start_ms * 1e3;
            report->passes[i].duration_us = dur_ms * 1e3;
        }
    }
    if (report) {
        report->n_passes = n_passes;
        report->cpu_duration_us =
            std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - cpu_start).count();
    }
    return HJ_OK;
}

}  // namespace

extern "C" hj_status hj_execute_graph(hj_device* dev, const hj_pass* passes, uint32_t n_passes,
                                      hj_buffer* const* env, const hj_buffer_desc* descs,
                                      uint32_t n_resources, hj_report* report) {
    HJ_REQUIRE(dev && (passes || n_passes == 0), "hj_execute_graph: null argument");
    HJ_REQUIRE((env && descs) || n_resources == 0, "hj_execute_graph: null environment");
    return execute_passes(dev, nullptr, passes, n_passes, env, descs, n_resources, report);
}

// ---- sharded execution ---------------------------------------------------------------------------
extern "C" void hj_shard_bounds(uint64_t n, int32_t world, int32_t rank, uint64_t* start, uint64_t* end) {
    uint64_t s = 0, e = 0;
    if (world >= 1 && rank >= 0 && rank < world) shard_bounds(n, world, rank, &s, &e);
    if (start) *start = s;
    if (end) *end = e;
}

// Placement propagation, host only.  Resources the caller has placed keep their placement; every
// HJ_RES_AUTO resource is placed by the passes that touch it, iterated to a fixed point so that
// demand also flows backwards (the zero-fill kernel in front of a Compress has no sharded input of
// its own, but its output is the index segment of a sharded mask):
//   kernel with a sharded resource -> every undecided resource of the pass extent that is only
//                                     addressed by the bare Index is SHARDED, the rest replicas;
//   Reduce                         -> dst replicated (every rank ends up with the global result);
//   PrefixSum                      -> dst like src (and src like a dst already sharded);
//   Compress                       -> index_out like the mask (and the mask like an index_out
//                                     already sharded), out_count replicated (global count).
// What no sharded pass ever touches stays a replica: the single-GPU program, run by every rank.
extern "C" hj_status hj_shard_plan(const hj_pass* passes, uint32_t n_passes, const hj_buffer_desc* descs, uint32_t n_resources,
                                   hj_shard_desc* shards) {
    HJ_REQUIRE((passes || n_passes == 0) && ((descs && shards) || n_resources == 0), "hj_shard_plan: null argument");
    std::vector<std::vector<SlotAccess>> access(n_passes);
    for (uint32_t i = 0; i < n_passes; i++) {
        const hj_pass& p = passes[i];
        for (uint32_t b = 0; b < p.n_resources; b++)
            HJ_REQUIRE(p.resources && p.resources[b] < n_resources, "pass %u: ResourceId out of range", i);
        switch (p.kind) {
        case HJ_PASS_KERNEL: {
            HJ_REQUIRE(p.ir, "kernel pass %u without IR", i);
            const std::string verr = validate_ir(p.ir);
            if (!verr.empty()) return fail(HJ_ERR_INVALID, "kernel pass %u: IR rejected: %s", i, verr.c_str());
            HJ_REQUIRE(p.ir->n_buffers == p.n_resources, "kernel pass %u: IR names %u buffers, the pass binds %u", i,
                       p.ir->n_buffers, p.n_resources);
            analyse_slot_access(p.ir, &access[i]);
            break;
        }
        case HJ_PASS_REDUCE: HJ_REQUIRE(p.n_resources >= 2, "reduce pass %u needs [dst, src]", i); break;
        case HJ_PASS_PREFIX_SUM: HJ_REQUIRE(p.n_resources >= 2, "prefix-sum pass %u needs [dst, src]", i); break;
        case HJ_PASS_COMPRESS: HJ_REQUIRE(p.n_resources >= 3, "compress pass %u needs [index_out, out_count, src]", i); break;
        default:
            return fail(HJ_ERR_UNSUPPORTED, "pass %u: device op %u is out of scope for the B200 backend", i, p.kind);
        }
    }
    // index segment -> the count its Compress pass writes (the size buffer of the DynSize kernels behind it)
    std::vector<int64_t> count_of_index(n_resources, -1);
    for (uint32_t i = 0; i < n_passes; i++)
        if (passes[i].kind == HJ_PASS_COMPRESS) count_of_index[passes[i].resources[0]] = (int64_t)passes[i].resources[1];
    bool changed = true;
    auto place = [&](uint32_t rid, uint32_t pl) {
        if (shards[rid].placement == HJ_RES_AUTO) {
            shards[rid].placement = pl;
            changed = true;
        }
    };
    auto is = [&](uint32_t rid, uint32_t pl) { return shards[rid].placement == pl; };
    for (uint32_t round = 0; changed && round <= n_passes + 1; round++) {
        changed = false;
        for (uint32_t i = 0; i < n_passes; i++) {
            const hj_pass& p = passes[i];
            switch (p.kind) {
            case HJ_PASS_KERNEL: {
                bool any = false;
                for (uint32_t b = 0; b < p.n_resources; b++) any = any || is(p.resources[b], HJ_RES_SHARDED);
                if (!any) break;  // undecided for now
                if (p.size_buffer >= 0) {
                    // DynSize: shards only as a SEGMENT kernel — one of its resources is the index segment of a
                    // sharded Compress whose count sizes the pass (or arrives as a segment); then what it reaches
                    // through the segment, or writes at Index, lives per rank as well
                    for (uint32_t c = 0; c < p.n_resources; c++) {
                        const uint32_t crid = p.resources[c];
                        const bool from_compress = count_of_index[crid] == (int64_t)p.size_buffer;
                        if (!is(crid, HJ_RES_SHARDED) || !(from_compress || is_segment_state(shards[crid].deferred)) ||
                            descs[crid].ty != HJ_U32 || descs[crid].size != p.size)
                            continue;
                        std::vector<SegmentAccess> acc;
                        if (!analyse_segment_access(p.ir, c, &acc, nullptr)) continue;
                        for (uint32_t b = 0; b < p.n_resources; b++) {
                            const uint32_t rid = p.resources[b];
                            const SegmentAccess& a = acc[b];
                            const bool per_rank = (a.read || a.written) && descs[rid].size == p.size &&
                                                  (a.through_segment || (a.index_only && !a.read));
                            place(rid, per_rank ? HJ_RES_SHARDED : HJ_RES_REPLICATED);
                        }
                        break;
                    }
                    break;
                }
                for (uint32_t b = 0; b < p.n_resources; b++) {
                    const uint32_t rid = p.resources[b];
                    const SlotAccess& a = access[i][b];
                    place(rid, a.index_only && (a.read || a.written) && descs[rid].size == p.size ? HJ_RES_SHARDED : HJ_RES_REPLICATED);
                }
                break;
            }
            case HJ_PASS_REDUCE: place(p.resources[0], HJ_RES_REPLICATED); break;
            case HJ_PASS_PREFIX_SUM:
                if (!is(p.resources[1], HJ_RES_AUTO)) place(p.resources[0], shards[p.resources[1]].placement);
                else if (is(p.resources[0], HJ_RES_SHARDED)) place(p.resources[1], HJ_RES_SHARDED);
                break;
            case HJ_PASS_COMPRESS:
                place(p.resources[1], HJ_RES_REPLICATED);
                if (!is(p.resources[2], HJ_RES_AUTO)) place(p.resources[0], shards[p.resources[2]].placement);
                else if (is(p.resources[0], HJ_RES_SHARDED)) place(p.resources[2], HJ_RES_SHARDED);
                break;
            default: break;
            }
        }
    }
    for (uint32_t r = 0; r < n_resources; r++)
        if (shards[r].placement == HJ_RES_AUTO) shards[r].placement = HJ_RES_REPLICATED;
    return HJ_OK;
}

extern "C" hj_status hj_execute_graph_sharded(hj_comm* comm, const hj_pass* passes, uint32_t n_passes, hj_buffer* const* env,
                                              const hj_buffer_desc* descs, uint32_t n_resources, hj_shard_desc* shards,
                                              hj_report* report) {
    HJ_REQUIRE(comm && (passes || n_passes == 0), "hj_execute_graph_sharded: null argument");
    HJ_REQUIRE((env && descs && shards) || n_resources == 0, "hj_execute_graph_sharded: null environment");
    ShardCtx sc;
    sc.comm = comm;
    sc.shards = shards;
    int32_t rank = 0, world = 1;
    HJ_TRY(hj_comm_info(comm, &rank, &world, nullptr, nullptr));
    sc.rank = rank;
    sc.world = world;
    for (uint32_t r = 0; r < n_resources; r++)
        HJ_REQUIRE(shards[r].placement == HJ_RES_REPLICATED || shards[r].placement == HJ_RES_SHARDED,
                   "resource %u has no placement (run hj_shard_plan first)", r);
    hj_device* dev = comm_device(comm);
    DeviceGuard g(dev);  // held throughout: the exchanges of one launch must not interleave with another thread's
    return execute_passes(dev, &sc, passes, n_passes, env, descs, n_resources, report);
}

namespace {
constexpr size_t MAX_INSTANCES_PER_KEY = 8;

void destroy_instance(GraphInstance& g) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
}
}  // namespace

namespace {
// The relaunch path shared by hj_execute_graph_cached and hj_execute_graph_sharded_cached.  For a
// sharded pass list (`sc`) an instance also belongs to the communicator, the placement, the seed
// buffers and the `deferred` state the inputs arrive in; the state the list leaves behind is recorded
// at capture and restored on every replay.  The sharded kernels read their exchange epoch from device
// memory (peer.cuh), so a replay needs no new launch parameters.
hj_status execute_cached(hj_device* dev, const ShardCtx* sc, uint64_t graph_key, const hj_pass* passes, uint32_t n_passes,
                         hj_buffer* const* env, const hj_buffer_desc* descs, uint32_t n_resources, uint32_t* how) {
    static const bool disabled = getenv("HJ_NO_CUDA_GRAPHS") != nullptr;
    if (how) *how = 0;
    DeviceGuard g(dev);  // held across the whole capture: nobody else may enqueue on the capturing stream
    if (!dev->gcache) dev->gcache = new GraphCache();
    GraphCache& gc = *dev->gcache;
    bool arriving = false;  // a buffer that is still being uploaded: its events cannot be waited for inside a capture
    if (dev->async_live.load(std::memory_order_acquire) != 0)
        for (uint32_t i = 0; i < n_resources; i++) arriving = arriving || (env[i] && env[i]->progress);
    if (disabled || n_passes == 0 || arriving) {
        gc.plain++;
        return execute_passes(dev, sc, passes, n_passes, env, descs, n_resources, nullptr);
    }
    std::vector<void*> sig(n_resources);
    for (uint32_t i = 0; i < n_resources; i++) sig[i] = env[i] ? env[i]->ptr : nullptr;
    if (sc) {
        sig.push_back((void*)sc->comm);
        for (uint32_t i = 0; i < n_resources; i++) {
            sig.push_back(sc->shards[i].seed ? sc->shards[i].seed->ptr : nullptr);
            sig.push_back((void*)(uintptr_t)(sc->shards[i].placement * 2u + (sc->shards[i].deferred ? 1u : 0u)));
        }
    }
    std::vector<GraphInstance>& insts = gc.by_key[graph_key];
    GraphInstance* inst = nullptr;
    for (GraphInstance& c : insts)
        if (c.ptrs == sig) inst = &c;
    auto run_plain = [&]() {
        gc.plain++;
        return execute_passes(dev, sc, passes, n_passes, env, descs, n_resources, nullptr);
    };

    if (!inst) {
        // first sight of these addresses: execute normally (compiles kernels, sizes every scratch)
        if (insts.size() >= MAX_INSTANCES_PER_KEY) {
            size_t oldest = 0;
            for (size_t i = 1; i < insts.size(); i++)
                if (insts[i].last_use < insts[oldest].last_use) oldest = i;
            destroy_instance(insts[oldest]);
            insts.erase(insts.begin() + (long)oldest);
        }
        GraphInstance fresh;
        fresh.ptrs = std::move(sig);
        fresh.last_use = ++gc.clock;
        insts.push_back(std::move(fresh));
        return run_plain();
    }
    inst->last_use = ++gc.clock;
    if (inst->exec && inst->generation != dev->lookback.generation) destroy_instance(*inst);  // stale scratch pointers
    if (inst->exec) {
        HJ_TRY(count_epoch(dev, inst->n_epochs));
        HJ_CUDA(cudaGraphLaunch(inst->exec, dev->stream));
        dev->launches.fetch_add(inst->n_launches, std::memory_order_relaxed);
        if (sc)
            for (uint32_t i = 0; i < n_resources; i++) sc->shards[i].deferred = inst->deferred_out[i];
        gc.replayed++;
        if (how) *how = 2;
        return HJ_OK;
    }
    if (inst->uncapturable) return run_plain();
    // second launch with these addresses: capture the pass list
    const uint32_t epochs0 = dev->lookback.epoch;
    const uint64_t launches0 = dev->launches.load(std::memory_order_relaxed);
    const uint64_t gen0 = dev->lookback.generation;
    std::vector<uint32_t> deferred_in;
    if (sc)
        for (uint32_t i = 0; i < n_resources; i++) deferred_in.push_back(sc->shards[i].deferred);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(dev->stream, cudaStreamCaptureModeThreadLocal);
    hj_status st = HJ_ERR_CUDA;
    if (e == cudaSuccess) {
        st = execute_passes(dev, sc, passes, n_passes, env, descs, n_resources, nullptr);
        e = cudaStreamEndCapture(dev->stream, &graph);
    }
    const uint32_t n_epochs = dev->lookback.epoch - epochs0;
    const bool ok = e == cudaSuccess && st == HJ_OK && graph && gen0 == dev->lookback.generation &&
                    dev->lookback.epoch >= epochs0;
    if (ok) e = cudaGraphInstantiate(&inst->exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (!ok || e != cudaSuccess) {
        // nothing ran (the work was only recorded): fall back to the pass-by-pass path for good
        cudaGetLastError();
        inst->exec = nullptr;
        inst->uncapturable = true;
        dev->lookback.epoch = epochs0;
        dev->launches.store(launches0, std::memory_order_relaxed);
        if (sc)
            for (uint32_t i = 0; i < n_resources; i++) sc->shards[i].deferred = deferred_in[i];
        return run_plain();
    }
    inst->generation = gen0;
    inst->n_epochs = n_epochs;
    inst->n_launches = dev->launches.load(std::memory_order_relaxed) - launches0;
    inst->deferred_out.clear();
    if (sc)
        for (uint32_t i = 0; i < n_resources; i++) inst->deferred_out.push_back(sc->shards[i].deferred);
    HJ_CUDA(cudaGraphLaunch(inst->exec, dev->stream));
    gc.captured++;
    if (how) *how = 1;
    return HJ_OK;
}
}  // namespace

extern "C" hj_status hj_execute_graph_cached(hj_device* dev, uint64_t graph_key, const hj_pass* passes,
                                             uint32_t n_passes, hj_buffer* const* env, const hj_buffer_desc* descs,
                                             uint32_t n_resources, uint32_t* how) {
    HJ_REQUIRE(dev && (passes || n_passes == 0), "hj_execute_graph_cached: null argument");
    HJ_REQUIRE((env && descs) || n_resources == 0, "hj_execute_graph_cached: null environment");
    return execute_cached(dev, nullptr, graph_key, passes, n_passes, env, descs, n_resources, how);
}

extern "C" hj_status hj_execute_graph_sharded_cached(hj_comm* comm, uint64_t graph_key, const hj_pass* passes, uint32_t n_passes,
                                                     hj_buffer* const* env, const hj_buffer_desc* descs, uint32_t n_resources,
                                                     hj_shard_desc* shards, uint32_t* how) {
    HJ_REQUIRE(comm && (passes || n_passes == 0), "hj_execute_graph_sharded_cached: null argument");
    HJ_REQUIRE((env && descs && shards) || n_resources == 0, "hj_execute_graph_sharded_cached: null environment");
    ShardCtx sc;
    sc.comm = comm;
    sc.shards = shards;
    int32_t rank = 0, world = 1, peer_memory = 0;
    HJ_TRY(hj_comm_info(comm, &rank, &world, &peer_memory, nullptr));
    sc.rank = rank;
    sc.world = world;
    for (uint32_t r = 0; r < n_resources; r++)
        HJ_REQUIRE(shards[r].placement == HJ_RES_REPLICATED || shards[r].placement == HJ_RES_SHARDED,
                   "resource %u has no placement (run hj_shard_plan first)", r);
    hj_device* dev = comm_device(comm);
    DeviceGuard g(dev);
    if (how) *how = 0;
    // without peer memory the exchanges are NCCL calls: those run pass by pass
    if (world > 1 && !peer_memory) return execute_passes(dev, &sc, passes, n_passes, env, descs, n_resources, nullptr);
    return execute_cached(dev, &sc, graph_key, passes, n_passes, env, descs, n_resources, how);
}

extern "C" hj_status hj_graph_cache_drop(hj_device* dev, uint64_t graph_key) {
    HJ_REQUIRE(dev, "null device");
    DeviceGuard g(dev);
    if (!dev->gcache) return HJ_OK;
    auto& map = dev->gcache->by_key;
    if (graph_key == 0) {
        for (auto& kv : map)
            for (GraphInstance& c : kv.second) destroy_instance(c);
        map.clear();
        return HJ_OK;
    }
    auto it = map.find(graph_key);
    if (it != map.end()) {
        for (GraphInstance& c : it->second) destroy_instance(c);
        map.erase(it);
    }
    return HJ_OK;
}

extern "C" hj_status hj_graph_cache_stats(hj_device* dev, uint64_t* captured, uint64_t* replayed, uint64_t* plain) {
    HJ_REQUIRE(dev, "null device");
    DeviceGuard g(dev);
    const GraphCache* gc = dev->gcache;
    if (captured) *captured = gc ? gc->captured : 0;
    if (replayed) *replayed = gc ? gc->replayed : 0;
    if (plain) *plain = gc ? gc->plain : 0;
    return HJ_OK;
}
