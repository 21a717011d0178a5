// tgraph_io.cpp — wire format of a compiled Graph.
//
// The reference keeps compiled graphs (`graph::Graph`, hephaestus-jit/src/graph.rs:145-151) and
// pipelines (vulkan_core/pipeline.rs:30-46) in memory only: every process start re-traces,
// re-schedules and re-compiles.  Here the cubins already persist on disk by IR hash (jit.cpp); this
// file adds the other half: a Graph — pass list with the flat IR of every kernel pass
// (ir.rs:40-46), resource table, inputs / outputs, and the CONTENTS of captured buffers — as one
// self-describing byte string, so a recorded function can be launched in a fresh process without
// tracing anything.
//
// Layout (little endian; every count is a u32 unless noted):
//   "HJGRAPH1" | abi version | type table | resources | inputs | outputs | passes | u64 checksum
//   type table : n, then per node {kind, elem, num, cols, rows, n_fields, fields[]}; ids are LOCAL
//                (position in this table, children before parents) and re-interned on load
//   resource   : kind u32 (Input / Captured / Internal), u64 element count, local type id,
//                u64 n_bytes + bytes (Captured only: the buffer's contents at serialise time)
//   pass       : resources[], i32 size_buffer, u32 is_kernel, u64 size, device op {code, arg},
//                and for kernel passes the IR arrays verbatim (hj_ir_var, deps, hj_type_desc,
//                struct_fields, n_buffers — the IR carries its own type table)
// The checksum detects corruption; structure is validated on load (ids in range, well-formed IR) so
// that a damaged file ends in an error, never in a crash of the parser
// (tests: test_graph_deserialize_survives_mutations_behind_a_valid_checksum).  It is a cache format for
// files the process wrote itself, not a sandbox: like any traced program, a graph can address its
// buffers out of range if its indices say so.
// Internal resources that were bound to LIVE variables when the graph was compiled (scheduled
// variables the caller still held) have no variable on the loading side: they become plain
// temporaries, and results leave through `outputs` (which is how record() returns them).
#include <atomic>
#include <cstring>
#include <memory>
#include <unordered_map>

#include "hj_internal.h"
#include "trace_internal.h"

namespace hj {
namespace tr {
namespace {

constexpr char MAGIC[8] = {'H', 'J', 'G', 'R', 'A', 'P', 'H', '1'};

struct Writer {
    std::vector<uint8_t> out;
    void raw(const void* p, size_t n) {
        const uint8_t* b = static_cast<const uint8_t*>(p);
        out.insert(out.end(), b, b + n);
    }
    void u32(uint32_t v) { raw(&v, 4); }
    void u64(uint64_t v) { raw(&v, 8); }
    template <typename T>
    void array(const std::vector<T>& v) {
        u32((uint32_t)v.size());
        if (!v.empty()) raw(v.data(), v.size() * sizeof(T));
    }
};

struct Reader {
    const uint8_t* p;
    size_t left;
    void raw(void* dst, size_t n) {
        if (n > left) throw TraceError("graph deserialise: truncated input");
        memcpy(dst, p, n);
        p += n;
        left -= n;
    }
    uint32_t u32() { uint32_t v; raw(&v, 4); return v; }
    uint64_t u64() { uint64_t v; raw(&v, 8); return v; }
    template <typename T>
    void array(std::vector<T>& v) {
        const uint32_t n = u32();
        if ((size_t)n * sizeof(T) > left) throw TraceError("graph deserialise: truncated input");
        v.resize(n);
        if (n) raw(v.data(), (size_t)n * sizeof(T));
    }
};

// local ids for the transitive closure of the types the resource table uses, children first
struct TypeTable {
    std::vector<TypeNode> nodes;
    std::unordered_map<TypeId, uint32_t> local;
    uint32_t add(TypeId t) {
        auto it = local.find(t);
        if (it != local.end()) return it->second;
        TypeNode n = type_node(t);
        if (n.kind == HJ_VEC || n.kind == HJ_ARRAY || n.kind == HJ_MAT) n.elem = add(n.elem);
        for (TypeId& f : n.fields) f = add(f);
        nodes.push_back(n);
        const uint32_t id = (uint32_t)nodes.size() - 1;
        local[t] = id;
        return id;
    }
};

}  // namespace

std::vector<uint8_t> serialize_graph(const Graph& g) {
    TypeTable tt;
    std::vector<uint32_t> res_ty(g.resource_descs.size());
    for (size_t i = 0; i < g.resource_descs.size(); i++) res_ty[i] = tt.add(g.resource_descs[i].ty);

    Writer w;
    w.raw(MAGIC, 8);
    w.u32(hj_abi_version());
    w.u32((uint32_t)tt.nodes.size());
    for (const TypeNode& n : tt.nodes) {
        w.u32(n.kind); w.u32(n.elem); w.u32(n.num); w.u32(n.cols); w.u32(n.rows);
        w.array(n.fields);
    }
    w.u32((uint32_t)g.resources.size());
    for (size_t i = 0; i < g.resources.size(); i++) {
        const GraphResource& r = g.resources[i];
        w.u32((uint32_t)r.kind);
        w.u64(g.resource_descs[i].size);
        w.u32(res_ty[i]);
        if (r.kind == GraphResource::Captured) {
            hj_buffer* buf = nullptr;
            {
                std::lock_guard<std::mutex> lock(g_trace_mu);
                const Var& v = g_trace.var(r.id);
                if (v.data.kind != Resource::Buffer) throw TraceError("graph serialise: captured resource holds no buffer");
                buf = v.data.buf;
                hj_buffer_retain(buf);
            }
            const size_t bytes = g.resource_descs[i].size * type_size(g.resource_descs[i].ty);
            std::vector<uint8_t> host(bytes);
            const hj_status s = bytes ? hj_buffer_to_host(buf, 0, bytes, host.data()) : HJ_OK;
            hj_buffer_release(buf);
            if (s != HJ_OK) throw TraceError(std::string("graph serialise: ") + hj_last_error());
            w.u64(bytes);
            w.raw(host.data(), bytes);
        }
    }
    w.array(g.inputs);
    w.array(g.outputs);
    w.u32((uint32_t)g.passes.size());
    for (const Pass& p : g.passes) {
        w.array(p.resources);
        w.u32((uint32_t)p.size_buffer);
        w.u32(p.is_kernel ? 1u : 0u);
        w.u64(p.size);
        w.u32(p.device_op.code);
        w.u32(p.device_op.arg);
        if (p.is_kernel) {
            w.array(p.ir.vars);
            w.array(p.ir.deps);
            w.array(p.ir.types);
            w.array(p.ir.struct_fields);
            w.u32(p.ir.n_buffers);
        }
    }
    w.u64(hash_bytes(w.out.data(), w.out.size()));
    return std::move(w.out);
}

Graph* deserialize_graph(hj_device* dev, const void* bytes, size_t n_bytes) {
    if (n_bytes < 8 + 4 + 8 || memcmp(bytes, MAGIC, 8) != 0) throw TraceError("graph deserialise: not a serialised graph");
    uint64_t want;
    memcpy(&want, static_cast<const uint8_t*>(bytes) + n_bytes - 8, 8);
    if (hash_bytes(bytes, n_bytes - 8) != want) throw TraceError("graph deserialise: checksum mismatch");
    Reader r{static_cast<const uint8_t*>(bytes) + 8, n_bytes - 16};
    const uint32_t abi = r.u32();
    if (abi != hj_abi_version()) throw TraceError("graph deserialise: written by another ABI version");

    const uint32_t n_types = r.u32();
    if ((size_t)n_types * 24 > r.left) throw TraceError("graph deserialise: truncated input");
    std::vector<TypeId> types(n_types);
    for (size_t i = 0; i < types.size(); i++) {
        TypeNode n;
        n.kind = r.u32(); n.elem = r.u32(); n.num = r.u32(); n.cols = r.u32(); n.rows = r.u32();
        r.array(n.fields);
        auto child = [&](uint32_t local) -> TypeId {
            if (local >= i) throw TraceError("graph deserialise: malformed type table");
            return types[local];
        };
        switch (n.kind) {
        case HJ_VEC: types[i] = type_vector(child(n.elem), n.num); break;
        case HJ_ARRAY: types[i] = type_array(child(n.elem), n.num); break;
        case HJ_MAT: types[i] = type_matrix(child(n.elem), n.cols, n.rows); break;
        case HJ_STRUCT: {
            std::vector<TypeId> f;
            for (TypeId l : n.fields) f.push_back(child(l));
            types[i] = type_struct(f.data(), (uint32_t)f.size());
            break;
        }
        default: types[i] = type_scalar(n.kind); break;
        }
    }

    std::unique_ptr<Graph> g(new Graph());
    static std::atomic<uint64_t> next_uid{1ull << 40};  // apart from the uids compile_graph hands out
    g->uid = next_uid.fetch_add(1);
    const uint32_t n_res = r.u32();
    for (uint32_t i = 0; i < n_res; i++) {
        GraphResource gr;
        const uint32_t kind = r.u32();
        if (kind > GraphResource::Internal) throw TraceError("graph deserialise: malformed resource table");
        gr.kind = (GraphResource::Kind)kind;
        BufferDesc d;
        d.size = r.u64();
        const uint32_t lt = r.u32();
        if (lt >= types.size()) throw TraceError("graph deserialise: malformed resource table");
        d.ty = types[lt];
        if (gr.kind == GraphResource::Captured) {
            const uint64_t nb = r.u64();
            if (nb > r.left || nb != d.size * type_size(d.ty)) throw TraceError("graph deserialise: malformed captured buffer");
            if (!dev) throw TraceError("graph deserialise: the graph captures buffers, a device is required");
            hj_buffer* buf = nullptr;
            if (hj_buffer_create_from_slice(dev, r.p, nb, &buf) != HJ_OK)
                throw TraceError(std::string("graph deserialise: ") + hj_last_error());
            r.p += nb;
            r.left -= nb;
            // the captured variable lives in the trace like any tr::array (trace.rs:647-663);
            // the graph owns the only reference
            Var v;
            v.op.kind = OpKind::Buffer;
            v.ty = d.ty;
            v.extent.n = d.size;
            v.data.kind = Resource::Buffer;
            v.data.buf = buf;  // takes over the reference create_from_slice returned
            std::lock_guard<std::mutex> lock(g_trace_mu);
            gr.id = g_trace.new_var_id(std::move(v));
        }
        g->resources.push_back(gr);
        g->resource_descs.push_back(d);
    }
    r.array(g->inputs);
    r.array(g->outputs);
    for (uint32_t rid : g->inputs)
        if (rid >= n_res) throw TraceError("graph deserialise: input out of range");
    for (uint32_t rid : g->outputs)
        if (rid >= n_res) throw TraceError("graph deserialise: output out of range");
    const uint32_t n_passes = r.u32();
    for (uint32_t i = 0; i < n_passes; i++) {
        Pass p;
        r.array(p.resources);
        for (uint32_t rid : p.resources)
            if (rid >= n_res) throw TraceError("graph deserialise: pass resource out of range");
        p.size_buffer = (int32_t)r.u32();
        if (p.size_buffer >= (int32_t)n_res) throw TraceError("graph deserialise: size buffer out of range");
        p.is_kernel = r.u32() != 0;
        p.size = r.u64();
        p.device_op.kind = p.is_kernel ? OpKind::Nop : OpKind::DeviceOp;
        p.device_op.code = r.u32();
        p.device_op.arg = r.u32();
        if (p.is_kernel) {
            r.array(p.ir.vars);
            r.array(p.ir.deps);
            r.array(p.ir.types);
            r.array(p.ir.struct_fields);
            p.ir.n_buffers = r.u32();
            if (p.resources.size() < p.ir.n_buffers) throw TraceError("graph deserialise: kernel pass binds too few resources");
            const hj_ir view = p.ir.view();
            const std::string problem = validate_ir(&view);
            if (!problem.empty()) throw TraceError("graph deserialise: pass " + std::to_string(i) + ": " + problem);
        } else if (p.device_op.code > DOP_COMPRESS) {
            throw TraceError("graph deserialise: unknown device op");
        }
        g->passes.push_back(std::move(p));
    }
    if (r.left != 0) throw TraceError("graph deserialise: trailing bytes");
    return g.release();
}

}  // namespace tr
}  // namespace hj
