/*
 * hj.h — C ABI of the B200-native (sm_100a) execution backend for hephaestus-jit.
 *
 * This is the drop-in boundary: exactly the entry points a Rust
 * `CudaDevice: BackendDevice` / `CudaBuffer: BackendBuffer` (the `todo!()` stubs at
 * hephaestus-jit/src/backend/cuda/mod.rs:16-73 and the `todo!()` arms at
 * hephaestus-jit/src/backend/mod.rs:104,112,118) would bind through `extern "C"`.
 * Plain pointers and sizes only; no torch / C++ types cross this line.
 * INTEGRATION.md shows the Rust-side binding.
 *
 * Conventions
 *   - every function returns hj_status (0 = ok); the message of the last failure on the
 *     calling thread is available through hj_last_error().
 *   - hj_device / hj_buffer / hj_kernel / hj_graph are intrusively ref-counted handles
 *     (the reference clones Arc<Buffer> freely: backend/vulkan/mod.rs:461-465).
 *   - all work is enqueued on the device's stream; calls that return data to the host
 *     (hj_buffer_to_host, hj_device_sync) block until it is visible, every other call is
 *     stream-ordered and asynchronous.  The reference blocks on a fence after every
 *     execute_graph (vulkan_core/device.rs:297-334); stream order gives the same
 *     observable behaviour to `to_host`.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *     with HJ_ERR_NO_DEVICE.
 */
#ifndef HJ_H
#define HJ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t hj_status;
enum {
    HJ_OK = 0,
    HJ_ERR_INVALID = 1,      /* bad argument (null handle, n == 0, size mismatch, ...)        */
    HJ_ERR_UNSUPPORTED = 2,  /* (op, type) pair the reference marks todo!() (reduce.rs:84-166) */
    HJ_ERR_CUDA = 3,         /* a CUDA runtime / driver call failed                            */
    HJ_ERR_NVRTC = 4,        /* runtime compilation of a fused kernel failed                   */
    HJ_ERR_NCCL = 5,         /* NCCL failure or libnccl not loadable                           */
    HJ_ERR_NO_DEVICE = 6,    /* no CUDA device / driver present                                */
    HJ_ERR_OOM = 7
};

/* Scalar element types; numbering follows the declaration order of `VarType`
 * (hephaestus-jit/src/vartype.rs:89-104). */
typedef enum {
    HJ_VOID = 0, HJ_BOOL = 1, HJ_I8 = 2, HJ_U8 = 3, HJ_I16 = 4, HJ_U16 = 5, HJ_I32 = 6,
    HJ_U32 = 7, HJ_I64 = 8, HJ_U64 = 9, HJ_F16 = 10, HJ_F32 = 11, HJ_F64 = 12,
    /* composite kinds, only used inside hj_type_desc */
    HJ_VEC = 13, HJ_ARRAY = 14, HJ_MAT = 15, HJ_STRUCT = 16
} hj_type_kind;

/* Reduction operators; numbering follows `ReduceOp` (hephaestus-jit/src/op.rs:90-99). */
typedef enum {
    HJ_REDUCE_MAX = 0, HJ_REDUCE_MIN = 1, HJ_REDUCE_SUM = 2, HJ_REDUCE_PROD = 3,
    HJ_REDUCE_OR = 4, HJ_REDUCE_AND = 5, HJ_REDUCE_XOR = 6
} hj_reduce_op;

typedef struct hj_device hj_device;
typedef struct hj_buffer hj_buffer;
typedef struct hj_kernel hj_kernel;

const char* hj_last_error(void);
/* Library/ABI version, bumped on every incompatible change. */
uint32_t hj_abi_version(void);
/* Number of visible CUDA devices (0 without a driver; never fails). */
int32_t hj_device_count(void);

/* ---- device --------------------------------------------------------------------------
 * replaces CudaDevice::create (backend/cuda/mod.rs:16-19); devices are process-global
 * singletons per ordinal like VulkanDevice::create (backend/vulkan/mod.rs:94-113). */
hj_status hj_device_create(int32_t ordinal, hj_device** out);
hj_status hj_device_retain(hj_device* dev);
hj_status hj_device_release(hj_device* dev);
hj_status hj_device_sync(hj_device* dev);
/* The cudaStream_t all work of this device is enqueued on. */
hj_status hj_device_stream(hj_device* dev, void** out_stream);
/* Enqueue on a caller-owned stream instead (e.g. torch's current stream); NULL restores
 * the device's own stream. */
hj_status hj_device_set_stream(hj_device* dev, void* stream);
hj_status hj_device_info(hj_device* dev, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor,
                         uint64_t* total_mem, uint64_t* l2_bytes);
/* Pool statistics (the reference's ResourcePool, vulkan_core/pool.rs:52-66). */
hj_status hj_device_pool_stats(hj_device* dev, uint64_t* bytes_live, uint64_t* bytes_cached,
                               uint64_t* n_alloc, uint64_t* n_free);
hj_status hj_device_pool_trim(hj_device* dev);
/* Number of kernel launches this library has enqueued on `dev` since creation. */
hj_status hj_device_launch_count(hj_device* dev, uint64_t* out);

/* ---- buffers -------------------------------------------------------------------------
 * replaces BackendDevice::create_buffer / create_buffer_from_slice
 * (backend/mod.rs:30-31, backend/vulkan/mod.rs:129-148,427-452) and
 * BackendBuffer::to_host (backend/mod.rs:39, backend/vulkan/mod.rs:478-509).
 * Contents of a fresh buffer are unspecified (recycled pool memory, pool.rs:36-41). */
hj_status hj_buffer_create(hj_device* dev, size_t bytes, hj_buffer** out);
hj_status hj_buffer_create_from_slice(hj_device* dev, const void* data, size_t bytes,
                                      hj_buffer** out);
/* create_buffer_from_slice that returns at once: `src` (pinned: hj_host_alloc) is copied chunk by chunk on a
 * side stream.  A single kernel pass over the bare Index that reads such buffers (hj_execute_graph with one
 * Kernel pass — what `x.fma(..).sin()..` of a traced program compiles to) is launched per chunk behind the
 * upload, and hj_buffer_to_host of its outputs drains per chunk: the reference's three blocking steps
 * `tr::array -> graph.launch -> to_vec` (trace.rs:647-663, graph.rs:315-323, trace.rs:1404-1438) overlap.
 * Every other use of the buffer first waits (on the device, not the host) for the whole upload.
 * `src` must stay valid and unchanged until a blocking call on a dependent result (hj_buffer_to_host,
 * hj_device_sync) has returned.  `elem_bytes`: element size of the array (chunks end on element boundaries). */
hj_status hj_buffer_create_from_host_async(hj_device* dev, const void* src, size_t bytes, size_t elem_bytes,
                                           hj_buffer** out);
/* The chunk schedule of the asynchronous upload for an array of `n` elements (host logic only, no device needed):
 * writes up to `capacity` (first, count) pairs and returns the number of chunks in *n_chunks (which may exceed
 * `capacity`).  Chunks are contiguous, cover [0, n), every boundary but the last is a multiple of 4096 elements;
 * a long array ramps 1/8, 1/4, 1/2 of a full chunk up at the front and down at the back (the fill and the drain of
 * the pipeline overlap nothing).  `chunk_elems` = 0: the default (2^24, or HJ_ASYNC_CHUNK_ELEMS). */
hj_status hj_async_chunk_schedule(uint64_t n, uint64_t chunk_elems, uint64_t* first, uint64_t* count, uint32_t capacity,
                                  uint32_t* n_chunks);
/* Non-owning view over device memory someone else allocated (e.g. a torch tensor). */
hj_status hj_buffer_wrap(hj_device* dev, void* device_ptr, size_t bytes, hj_buffer** out);
hj_status hj_buffer_retain(hj_buffer* buf);
hj_status hj_buffer_release(hj_buffer* buf);
hj_status hj_buffer_to_host(hj_buffer* buf, size_t offset_bytes, size_t nbytes, void* dst);
/* Stream-ordered host->device copy into an existing buffer (no allocation). */
hj_status hj_buffer_upload(hj_buffer* buf, size_t offset_bytes, const void* src, size_t nbytes);
hj_status hj_buffer_fill_zero(hj_buffer* buf);
hj_status hj_buffer_device_ptr(hj_buffer* buf, void** out);
hj_status hj_buffer_size(hj_buffer* buf, size_t* out_bytes);
hj_status hj_buffer_device(hj_buffer* buf, hj_device** out); /* borrowed, not retained */
/* Pinned host staging memory for the end-to-end path. */
hj_status hj_host_alloc(size_t bytes, void** out);
/* The same on the NUMA node `dev` is attached to (best effort; see runtime.cpp). */
hj_status hj_host_alloc_near(hj_device* dev, size_t bytes, void** out);
hj_status hj_host_free(void* ptr);

/* ---- device ops (hand-written sm_100a kernels) ---------------------------------------
 * Direct entry points for the three DeviceOp passes execute_graph dispatches to
 * (backend/vulkan/mod.rs:259-299).  `n` is the element count of `src`. */

/* dst[0] = fold(op, src[0..n)); replaces builtin::reduce::reduce (builtin/reduce.rs:22-314).
 * Integer results are exact (wrapping); f32/f64 sums are reordered (tolerance in DESIGN.md). */
hj_status hj_reduce(hj_device* dev, hj_reduce_op op, hj_type_kind ty, size_t n,
                    hj_buffer* src, hj_buffer* dst);
/* dst[i] = sum(src[0..=i]) (inclusive) or sum(src[0..i)) (exclusive); replaces
 * builtin::prefix_sum::prefix_sum_large (builtin/prefix_sum.rs:31-162).
 * `seed` (may be NULL) is a 1-element device buffer added to every output: the cross-GPU
 * offset of a sharded scan. */
hj_status hj_prefix_sum(hj_device* dev, hj_type_kind ty, size_t n, int32_t inclusive,
                        hj_buffer* src, hj_buffer* dst, hj_buffer* seed);
/* index_out[0..count) = ascending positions of non-zero mask bytes, out_count[0] = count;
 * replaces builtin::compress::compress_large (builtin/compress.rs:157-283).
 * Elements of index_out at and beyond `count` are left untouched, exactly like the
 * reference (they stay zero from the scheduler's zero-fill pass, trace.rs:1600-1601).
 * `size_buf` (may be NULL) holds a device-resident u32 element count <= n (DynSize).
 * `index_base` is added to every emitted index (shard offset, 0 on a single GPU). */
hj_status hj_compress(hj_device* dev, size_t n, hj_buffer* size_buf, hj_buffer* out_count,
                      hj_buffer* src_mask, hj_buffer* index_out, uint32_t index_base);
/* dst[idx[i]] = op(dst[idx[i]], value) for i in 0..n  — the ScatterReduce kernel op
 * (codegen/glsl/mod.rs:400-446) specialised for the histogram shape: `src` is either a
 * buffer of n values or NULL (then `literal` is the u64 bit pattern of the value).
 * Supported: op Sum/Min/Max/Or/And/Xor, ty U32/I32/F32(sum)/U64(sum). */
hj_status hj_scatter_reduce(hj_device* dev, hj_reduce_op op, hj_type_kind ty, size_t n,
                            hj_buffer* idx, hj_buffer* src, uint64_t literal,
                            hj_buffer* dst, size_t n_dst);
/* dst[i] = src[idx[i]] for i in 0..n (Gather kernel op, codegen/glsl/mod.rs:523-579),
 * elem_bytes in {1,2,4,8,16}. */
hj_status hj_gather(hj_device* dev, size_t elem_bytes, size_t n, hj_buffer* src,
                    hj_buffer* idx, hj_buffer* dst);

/* The same three device ops over HOST arrays (pinned memory for full PCIe speed): the array is
 * streamed through the GPU in chunks of `chunk_elems` elements (0 = default), upload, kernel and
 * download overlapped on three streams; what crosses a chunk boundary (partials, the running total,
 * the index base) stays on the device.  The pipelined form of the reference's blocking
 * tr::array -> launch -> to_vec sequence (trace.rs:647-663, 1404-1438) for a single device op;
 * results are complete on return.  hj_compress_host: host_index_out[0..*host_count) receives the
 * indices (+ index_base), entries beyond are not touched. */
hj_status hj_reduce_host(hj_device* dev, hj_reduce_op op, hj_type_kind ty, size_t n,
                         const void* host_src, void* host_dst /* one element */, size_t chunk_elems);
hj_status hj_prefix_sum_host(hj_device* dev, hj_type_kind ty, size_t n, int32_t inclusive,
                             const void* host_src, void* host_dst, size_t chunk_elems);
hj_status hj_compress_host(hj_device* dev, size_t n, const uint8_t* host_mask, uint32_t* host_index_out,
                           uint32_t* host_count, uint32_t index_base, size_t chunk_elems);

/* ---- fused-kernel IR (NVRTC path) ----------------------------------------------------
 * Flat mirror of `IR` / `ir::Var` (hephaestus-jit/src/ir.rs:12-46) and of the interned
 * `VarType` tree (vartype.rs:89-122). */

/* One node of the type table.  Scalars: kind only.  Vec/Array: elem + num.
 * Mat: elem + cols + rows.  Struct: `num` field types at struct_fields[first_field..]. */
typedef struct {
    uint32_t kind;        /* hj_type_kind */
    uint32_t elem;        /* index into types[] (Vec/Array/Mat) */
    uint32_t num;         /* Vec/Array length, Struct field count */
    uint32_t cols, rows;  /* Mat */
    uint32_t first_field; /* Struct: offset into struct_fields[] */
} hj_type_desc;

/* KernelOp tags (hephaestus-jit/src/op.rs:49-88), `arg` carries the payload. */
typedef enum {
    HJ_OP_NOP = 0, HJ_OP_SCATTER = 1,
    HJ_OP_SCATTER_REDUCE = 2,  /* arg = hj_reduce_op */
    HJ_OP_SCATTER_ATOMIC = 3,  /* arg = hj_reduce_op */
    HJ_OP_ATOMIC_INC = 4, HJ_OP_GATHER = 5, HJ_OP_INDEX = 6, HJ_OP_LITERAL = 7,
    HJ_OP_EXTRACT = 8,         /* arg = element */
    HJ_OP_DYN_EXTRACT = 9, HJ_OP_CONSTRUCT = 10, HJ_OP_SELECT = 11,
    HJ_OP_LOOP_START = 12, HJ_OP_LOOP_END = 13, HJ_OP_IF_START = 14, HJ_OP_IF_END = 15,
    HJ_OP_TEX_LOOKUP = 16, HJ_OP_TRACE_RAY = 17, /* out of scope: rejected */
    HJ_OP_BOP = 18,            /* arg = hj_bop */
    HJ_OP_UOP = 19,            /* arg = hj_uop */
    HJ_OP_FMA = 20, HJ_OP_BUFFER_REF = 21,
    HJ_OP_TEXTURE_REF = 22, HJ_OP_ACCEL_REF = 23 /* out of scope: rejected */
} hj_kernel_op;

/* Bop / Uop numbering follows op.rs:2-31 / op.rs:33-46. */
typedef enum {
    HJ_BOP_ADD = 0, HJ_BOP_SUB, HJ_BOP_MUL, HJ_BOP_DIV, HJ_BOP_MODULUS, HJ_BOP_MIN, HJ_BOP_MAX,
    HJ_BOP_INNER, HJ_BOP_AND, HJ_BOP_OR, HJ_BOP_XOR, HJ_BOP_SHL, HJ_BOP_SHR,
    HJ_BOP_EQ, HJ_BOP_NEQ, HJ_BOP_LT, HJ_BOP_LE, HJ_BOP_GT, HJ_BOP_GE
} hj_bop;
typedef enum {
    HJ_UOP_CAST = 0, HJ_UOP_BITCAST, HJ_UOP_NEG, HJ_UOP_SQRT, HJ_UOP_ABS, HJ_UOP_SIN,
    HJ_UOP_COS, HJ_UOP_EXP2, HJ_UOP_LOG2
} hj_uop;

typedef struct {
    uint32_t ty;        /* index into types[] */
    uint32_t op;        /* hj_kernel_op */
    uint32_t arg;       /* op payload */
    uint32_t dep_start; /* deps[dep_start..dep_end) */
    uint32_t dep_end;
    uint32_t _pad;
    uint64_t data;      /* literal bits / buffer slot (ir.rs:16) */
} hj_ir_var;

typedef struct {
    const hj_ir_var* vars;
    uint32_t n_vars;
    const uint32_t* deps;
    uint32_t n_deps;
    const hj_type_desc* types;
    uint32_t n_types;
    const uint32_t* struct_fields;
    uint32_t n_struct_fields;
    uint32_t n_buffers;
} hj_ir;

/* Content hash of an IR (the kernel-cache key; plays the role of Prehashed<IR>,
 * prehashed.rs:9-26).  Stable across processes. */
uint64_t hj_ir_hash(const hj_ir* ir);
/* Lower an IR to CUDA C++ (what the NVRTC stage compiles).  Returns a malloc'ed,
 * NUL-terminated string the caller frees with hj_free_string. Works without a GPU. */
hj_status hj_ir_codegen(const hj_ir* ir, char** out_source);
void hj_free_string(char* s);
/* Compile (or fetch from the per-device cache keyed by IR hash) the fused kernel for
 * `ir`; replaces VulkanDevice::compile_ir + Pipeline::create
 * (backend/vulkan/mod.rs:83-91, vulkan_core/pipeline.rs:30-46). */
hj_status hj_kernel_get(hj_device* dev, const hj_ir* ir, hj_kernel** out);
hj_status hj_kernel_release(hj_kernel* k);
/* Compile `ir` to an sm_100a cubin without a device (build check / cache warm-up).
 * The cubin is malloc'ed; free with hj_free_string((char*)cubin). */
hj_status hj_ir_compile_cubin(const hj_ir* ir, void** out_cubin, size_t* out_size);
/* Launch over `size` elements (or the device-resident u32 in `size_buf` if non-NULL,
 * capped by `size`), buffers in IR slot order; replaces the Kernel arm of execute_graph
 * (backend/vulkan/mod.rs:194-257).  `index_base` offsets KernelOp::Index (sharding); the value
 * 0xffffffff (never a block start: sizes fit u32) means "read it from size_buf[1]" — `size_buf` must then
 * hold two u32 (count, base), which is how the sharded pass interpreter runs DynSize kernels per segment. */
hj_status hj_kernel_launch(hj_device* dev, hj_kernel* k, size_t size, hj_buffer* size_buf,
                           hj_buffer* const* buffers, uint32_t n_buffers, uint32_t index_base);
/* Out-of-core elementwise map over HOST arrays (pinned memory for full PCIe speed): array i is
 * the kernel's buffer slot i and holds `n` elements; every slot must be accessed through the bare
 * Index only.  Chunks of `chunk_elems` elements (0 = default) are pipelined through upload /
 * kernel / download streams; results are visible on return.  The pipelined form of
 * tr::array -> launch -> to_vec (trace.rs:647-663, 1404-1438). */
hj_status hj_kernel_map_host(hj_device* dev, hj_kernel* k, size_t n, void* const* host_arrays,
                             uint32_t n_arrays, size_t chunk_elems);
/* cache statistics: compiled (misses) / hits, and on-disk cubin cache hits */
hj_status hj_device_kernel_cache_stats(hj_device* dev, uint64_t* n_compiled, uint64_t* n_hits,
                                       uint64_t* n_disk_hits);

/* ---- execute_graph -------------------------------------------------------------------
 * replaces BackendDevice::execute_graph (backend/mod.rs:33, backend/vulkan/mod.rs:151-383).
 * The pass list is what Graph::passes() exposes (graph.rs:409-425). */
typedef enum {
    HJ_PASS_KERNEL = 0, HJ_PASS_REDUCE = 1, HJ_PASS_PREFIX_SUM = 2, HJ_PASS_COMPRESS = 3
} hj_pass_kind;

typedef struct {
    uint32_t kind;            /* hj_pass_kind */
    uint32_t arg;             /* reduce: hj_reduce_op; prefix sum: inclusive flag */
    const uint32_t* resources; /* ResourceId list in the reference's order (see hj.h top of
                                  section): kernel = IR slot order; reduce/scan = [dst, src];
                                  compress = [index_out, out_count, src] */
    uint32_t n_resources;
    int32_t size_buffer;      /* ResourceId of the DynSize buffer, or -1 */
    const hj_ir* ir;          /* kernel passes */
    uint64_t size;            /* kernel passes: static element count */
} hj_pass;

typedef struct {
    uint64_t size; /* elements */
    uint32_t ty;   /* hj_type_kind of the element (scalars) or HJ_STRUCT etc. */
    uint32_t elem_bytes;
} hj_buffer_desc;

/* Per-pass timing, mirror of PassReport/ExecReport (backend/report.rs:2-19). */
typedef struct {
    char name[64];
    double start_us;
    double duration_us;
} hj_pass_report;
typedef struct {
    double cpu_duration_us;
    uint32_t n_passes;
    hj_pass_report* passes; /* caller-provided array of >= n_passes entries, or NULL */
    uint32_t passes_capacity;
} hj_report;

hj_status hj_execute_graph(hj_device* dev, const hj_pass* passes, uint32_t n_passes,
                           hj_buffer* const* env, const hj_buffer_desc* descs,
                           uint32_t n_resources, hj_report* report /* may be NULL */);

/* The zero-fill of a Compress pass's index buffer (`index = sized_literal(0, n)`, trace.rs:1600-1601)
 * is left to the compaction when the kernel pass in front holds it as a plain store: returns 1 if
 * buffer slot `slot` of `ir` is only ever written by a top-level, unconditional
 * Scatter(BufferRef(slot), Literal 0u32, Index) and never read; *scatter_var = that variable,
 * *nothing_else = 1 if the kernel has no other side effect (the pass is then skipped).  Host-only. */
int32_t hj_ir_index_zero_fill(const hj_ir* ir, uint32_t slot, uint32_t* scatter_var, int32_t* nothing_else);

/* The same call for a pass list the caller launches again and again — what FCache::call does
 * with a recorded function's Graph (hephaestus-jit/src/record.rs:120-210, graph.rs:192-400; the
 * reference re-records and re-submits a Vulkan command buffer on every launch,
 * backend/vulkan/mod.rs:151-383).  `graph_key` names the pass list (the caller guarantees that one
 * key always comes with the same passes and descs).  The first launch of a (key, buffer
 * addresses) pair executes normally, the second captures the whole pass list into ONE CUDA graph,
 * every later one replays it with a single cudaGraphLaunch.  Launches whose buffers sit at other
 * addresses than any captured instance take the normal path.  `how`, if not NULL, receives
 * 0 = executed pass by pass, 1 = captured and launched, 2 = replayed. */
hj_status hj_execute_graph_cached(hj_device* dev, uint64_t graph_key, const hj_pass* passes,
                                  uint32_t n_passes, hj_buffer* const* env,
                                  const hj_buffer_desc* descs, uint32_t n_resources,
                                  uint32_t* how /* may be NULL */);
/* Drops the captured instances of one key (0: of every key), e.g. when the caller frees a Graph. */
hj_status hj_graph_cache_drop(hj_device* dev, uint64_t graph_key);
/* Counters since device creation: graphs captured, graph replays, uncaptured launches through
 * hj_execute_graph_cached. */
hj_status hj_graph_cache_stats(hj_device* dev, uint64_t* captured, uint64_t* replayed,
                               uint64_t* plain);

/* ---- sharded (multi-GPU) ops ---------------------------------------------------------
 * One process per GPU; `hj_comm` wraps an NCCL communicator created from a unique id the
 * host distributes (torch.distributed store / broadcast).  The reference has no multi-GPU
 * code; these combine per-GPU partials exactly as BASELINE.json names. */
typedef struct hj_comm hj_comm;
#define HJ_UNIQUE_ID_BYTES 128
hj_status hj_comm_unique_id(uint8_t out_id[HJ_UNIQUE_ID_BYTES]);
hj_status hj_comm_create(hj_device* dev, const uint8_t id[HJ_UNIQUE_ID_BYTES], int32_t rank,
                         int32_t world, hj_comm** out);
/* The same communicator WITHOUT NCCL: the scalar / small-array exchanges of the sharded ops only need
 * every rank's mailbox mapped into every process (CUDA IPC).  hj_comm_create_local allocates this
 * rank's mailbox and returns its IPC handle; the host all-gathers the `world` handles over any
 * transport it likes (torch.distributed with gloo, MPI, a pipe) and passes them, in rank order, to
 * hj_comm_connect.  A communicator built this way supports hj_sharded_reduce / prefix_sum(_deferred) /
 * compress, hj_sharded_scatter_reduce for 4-byte arrays of <= 65536 elements and
 * hj_execute_graph_sharded; hj_sharded_rebalance and larger arrays need NCCL and fail with
 * HJ_ERR_NCCL.  Ranks may share one GPU (useful for testing: the kernels of different processes are
 * time-sliced, the exchange is the same code). */
#define HJ_IPC_HANDLE_BYTES 64
hj_status hj_comm_create_local(hj_device* dev, int32_t rank, int32_t world, hj_comm** out,
                               uint8_t handle_out[HJ_IPC_HANDLE_BYTES]);
hj_status hj_comm_connect(hj_comm* comm, const uint8_t* handles /* world x HJ_IPC_HANDLE_BYTES */);
hj_status hj_comm_info(hj_comm* comm, int32_t* rank, int32_t* world, int32_t* peer_memory, int32_t* has_nccl);
hj_status hj_comm_device(hj_comm* comm, hj_device** out); /* borrowed, not retained */
hj_status hj_comm_destroy(hj_comm* comm);
/* local reduce of this rank's shard, then all-reduce of the partials: dst[0] on every rank
 * holds the global result. */
hj_status hj_sharded_reduce(hj_comm* comm, hj_reduce_op op, hj_type_kind ty, size_t n_local,
                            hj_buffer* src, hj_buffer* dst);
/* MATERIALISED sharded scan: shard totals (a read-only pass with the exchange fused into its last
 * CTA) -> exclusive offset -> seeded local scan; dst holds this rank's slice of the global scan.
 * 3 * sizeof(T) bytes of HBM traffic per element. */
hj_status hj_sharded_prefix_sum(hj_comm* comm, hj_type_kind ty, size_t n_local,
                                int32_t inclusive, hj_buffer* src, hj_buffer* dst);
/* Sharded scan with a DEFERRED seed, 2 * sizeof(T) bytes per element, one kernel: dst receives the
 * LOCAL scan of the shard and seed_out[0] (device, one element of `ty`) this rank's exclusive offset
 * (sum of the totals of the ranks before it, exchanged over peer memory by the CTA that owns the
 * last tile).  The global scan is dst[i] + seed_out[0] — "segment + offset", like the offsets table
 * of hj_sharded_compress.  Fused kernels of hj_execute_graph_sharded add the seed when they load
 * the value; hj_apply_seed materialises it. */
hj_status hj_sharded_prefix_sum_deferred(hj_comm* comm, hj_type_kind ty, size_t n_local,
                                         int32_t inclusive, hj_buffer* src, hj_buffer* dst,
                                         hj_buffer* seed_out);
/* buf[i] += seed[0] for i in 0..n (in place, 2 * sizeof(T) bytes per element). */
hj_status hj_apply_seed(hj_device* dev, hj_type_kind ty, size_t n, hj_buffer* buf, hj_buffer* seed);
/* local compress with global indices (index_base = global start of the shard); all-gather
 * of the counts: counts_out[world] (u32, device) and out_count[0] = global count. */
hj_status hj_sharded_compress(hj_comm* comm, size_t n_local, uint32_t index_base,
                              hj_buffer* src_mask, hj_buffer* index_out, hj_buffer* out_count,
                              hj_buffer* counts_out);
/* privatised local histogram, then all-reduce (sum) of the n_dst bins. */
hj_status hj_sharded_scatter_reduce(hj_comm* comm, hj_reduce_op op, hj_type_kind ty,
                                    size_t n_local, hj_buffer* idx, hj_buffer* src,
                                    uint64_t literal, hj_buffer* dst, size_t n_dst);
/* Re-partition a sharded compacted sequence evenly over the ranks, order preserved: rank q holds
 * counts[q] elements of `elem_bytes` in `src` (counts: u32[world] on the device, as written by
 * hj_sharded_compress); afterwards `dst` holds this rank's block of the concatenated sequence,
 * out_count[0] (device u32, may be NULL) and *new_count_host (may be NULL) its length.  Block
 * boundaries are q*T/W + min(q, T%W).  Slices move GPU to GPU (grouped ncclSend/ncclRecv). */
hj_status hj_sharded_rebalance(hj_comm* comm, size_t elem_bytes, hj_buffer* src, hj_buffer* counts,
                               hj_buffer* dst, hj_buffer* out_count, uint64_t* new_count_host);

/* ---- sharded execute_graph ---------------------------------------------------------------
 * The pass interpreter over arrays partitioned across the ranks of `comm` (one process per GPU):
 * what BackendDevice::execute_graph (backend/vulkan/mod.rs:151-383) does for one device, pass for
 * pass, for a Graph whose large arrays are split into contiguous blocks.
 *   Kernel pass over sharded data  launched over the rank's block with KernelOp::Index = global index
 *                                  (index_base); replicas may be read (gather tables), not written;
 *                                  a `dst[keys[i]] += literal` pass is the privatised histogram + the
 *                                  combine over peer memory (hj_sharded_scatter_reduce);
 *   Reduce                         hj_sharded_reduce, result replicated;
 *   PrefixSum                      hj_sharded_prefix_sum_deferred when the destination carries a seed
 *                                  buffer (integer types): 2 * sizeof(T) bytes/element, the consumers add
 *                                  the offset; else hj_sharded_prefix_sum;
 *   Compress                       local compaction with global indices, counts exchanged in the kernel;
 *                                  index_out stays a per-rank segment, out_count is the global count.
 *   DynSize kernel behind a        the wavefront step (jit/test.rs:1020-1062: compress_dyn -> gather /
 *   sharded Compress               scatter through the compacted indices): every rank runs the kernel over
 *                                  ITS segment, sized by its own count on the device; the segment's indices
 *                                  are global and fall into the rank's block by construction, so sharded
 *                                  arrays of the mask's extent are addressed through them in place.  Needs
 *                                  a seed buffer on the index segment (it receives the rank's count).  A
 *                                  DynSize Compress whose mask is aligned with a segment (nested compaction)
 *                                  compacts the rank's part: HJ_SHARD_SEGMENT_LOCAL.
 * Not sharded (SURVEY 8e "replicas only"): access to a sharded resource through any other computed
 * index, writes to a replica from a sharded kernel, a replica read at the bare Index inside a segment
 * kernel, device ops over a segment -> HJ_ERR_UNSUPPORTED.  (KernelOp::Index as a VALUE inside a segment
 * kernel is the position in the GLOBAL compacted sequence, as on one GPU: the rank's offset sits behind
 * its count in the seed buffer.) */
typedef enum { HJ_RES_REPLICATED = 0, HJ_RES_SHARDED = 1, HJ_RES_AUTO = 2 } hj_placement;
/* values of hj_shard_desc.deferred */
#define HJ_SHARD_PLAIN 0u    /* the buffer holds the rank's block as it is */
#define HJ_SHARD_DEFERRED 1u /* a LOCAL scan: the global value of element i is buffer[i] + seed[0] */
#define HJ_SHARD_SEGMENT 2u  /* a per-rank compacted SEGMENT (Compress index output, or what a DynSize
                                kernel wrote at Index): only the first seed[0] (u32) entries are defined;
                                seed[1] (u32) = entries of the ranks before this one */
#define HJ_SHARD_SEGMENT_LOCAL 3u /* the index output of a DynSize Compress over a segment-aligned mask (nested
                                     compaction, jit/test.rs:976-1019): a segment whose entries are positions in
                                     the RANK's part of the parent sequence — they address the rank's segments in
                                     place; on one GPU the same entries are positions in the whole parent sequence
                                     (= these + the parent's seed[1]) */
typedef struct {
    uint32_t placement; /* hj_placement.  SHARDED: the buffer holds this rank's block
                           [start, end) = hj_shard_bounds(descs[i].size, world, rank) of the global array */
    uint32_t deferred;  /* in / out, SHARDED resources: HJ_SHARD_PLAIN / _DEFERRED / _SEGMENT / _SEGMENT_LOCAL */
    hj_buffer* seed;    /* >= 8 bytes (one element of the resource's type for a scan; two u32 for a segment) on the device, or
                           NULL (then PrefixSum results are always materialised and a Compress index
                           segment cannot size dependent DynSize kernels) */
} hj_shard_desc;
/* Block of rank `rank`: sizes differ by at most one, order preserved. */
void hj_shard_bounds(uint64_t n, int32_t world, int32_t rank, uint64_t* start, uint64_t* end);
/* Host only: fills in every HJ_RES_AUTO placement from the passes that write the resource. */
hj_status hj_shard_plan(const hj_pass* passes, uint32_t n_passes, const hj_buffer_desc* descs,
                        uint32_t n_resources, hj_shard_desc* shards);
hj_status hj_execute_graph_sharded(hj_comm* comm, const hj_pass* passes, uint32_t n_passes,
                                   hj_buffer* const* env, const hj_buffer_desc* descs,
                                   uint32_t n_resources, hj_shard_desc* shards,
                                   hj_report* report /* may be NULL */);

/* hj_execute_graph_cached for sharded pass lists: the first launch of a (key, buffers, placement,
 * seeds, incoming `deferred` state) combination executes pass by pass, the second is captured into
 * ONE CUDA graph, later ones replay it (the exchange epochs of the sharded kernels live in device
 * memory, so nothing changes between launches).  Every rank must make the same sequence of calls. */
hj_status hj_execute_graph_sharded_cached(hj_comm* comm, uint64_t graph_key, const hj_pass* passes,
                                          uint32_t n_passes, hj_buffer* const* env,
                                          const hj_buffer_desc* descs, uint32_t n_resources,
                                          hj_shard_desc* shards, uint32_t* how /* may be NULL */);

/* ---- trace / schedule / graph (host side) ------------------------------------------------
 * C++ restatement of the layers ABOVE the backend traits, so that programs written against
 * the reference's op vocabulary produce the same Graph / IR and drive this backend end to end:
 * hephaestus-jit/src/trace.rs (op constructors, new_var scheduling rules :364-399),
 * graph.rs (compile :436-614, launch_with :192-400), compiler.rs (:20-230),
 * record.rs (FCache :116-210).  A Rust build keeps the reference's own trace layer and binds
 * only the sections above; this section is what makes the backend testable end to end here.
 *
 * A variable handle (`uint64_t`) is an owned reference like the reference's VarRef: every
 * function that returns one hands out a new reference the caller releases with
 * hj_tr_var_release; handles passed in are borrowed.  0 is never a valid handle; `active`
 * arguments accept 0 for "unconditional".  The trace is process-global behind one mutex, the
 * schedule is thread-local (trace.rs:71-76). */
typedef struct hj_graph hj_graph;

/* interned VarType tree (vartype.rs:20-122); scalars: type id == hj_type_kind */
uint32_t hj_tr_type_scalar(uint32_t kind);
uint32_t hj_tr_type_vector(uint32_t elem, uint32_t num);
uint32_t hj_tr_type_array(uint32_t elem, uint32_t num);
uint32_t hj_tr_type_matrix(uint32_t elem, uint32_t cols, uint32_t rows);
uint32_t hj_tr_type_struct(const uint32_t* fields, uint32_t n);
size_t hj_tr_type_size(uint32_t ty);                      /* vartype.rs:125-155 */
size_t hj_tr_type_alignment(uint32_t ty);                 /* vartype.rs:169-189 */
size_t hj_tr_type_offset(uint32_t ty, uint32_t elem);     /* vartype.rs:156-168 */
uint32_t hj_tr_type_kind(uint32_t ty);

hj_status hj_tr_var_retain(uint64_t v);                   /* VarRef::clone, trace.rs:281-286 */
hj_status hj_tr_var_release(uint64_t v);                  /* VarRef::drop,  trace.rs:287-291 */
hj_status hj_tr_var_info(uint64_t v, uint32_t* ty, int32_t* dynamic, uint64_t* extent,
                         int32_t* evaluated, uint64_t* rc, int32_t* dirty);
uint64_t hj_tr_var_hash(uint64_t v);                      /* impl Hash for VarRef, trace.rs:250-265 */
int32_t hj_tr_is_empty(void);                             /* tr::is_empty, trace.rs:220-222 */
uint64_t hj_tr_n_live(void);
hj_status hj_tr_var_buffer(uint64_t v, hj_buffer** out);  /* borrowed; NULL if not evaluated */

hj_status hj_tr_index(uint64_t* out);                                             /* trace.rs:552-562 */
hj_status hj_tr_sized_index(uint64_t n, uint64_t* out);                           /* :567-577 */
hj_status hj_tr_dynamic_index(uint64_t capacity, uint64_t size_var, uint64_t* out); /* :578-597 */
hj_status hj_tr_literal(uint32_t ty, uint64_t bits, uint64_t* out);               /* :602-616 */
hj_status hj_tr_sized_literal(uint32_t ty, uint64_t bits, uint64_t n, uint64_t* out); /* :627-641 */
hj_status hj_tr_array(hj_device* dev, uint32_t ty, const void* data, uint64_t n, uint64_t* out); /* :647-663 */
hj_status hj_tr_from_buffer(hj_buffer* buf, uint32_t ty, uint64_t n, uint64_t* out);
/* Sharded arrays (no reference counterpart; BASELINE north star: "large arrays are partitioned across
 * the GPUs"): a variable of n_global elements whose buffer holds this rank's contiguous block
 * (hj_shard_bounds).  Everything traced from it is scheduled exactly like the reference schedules the
 * unsharded program; Graph launches that meet a sharded variable run through hj_execute_graph_sharded.
 * The communicator is borrowed and must outlive the variables. */
/* tr::array over pinned host memory with the asynchronous upload of hj_buffer_create_from_host_async */
hj_status hj_tr_array_async(hj_device* dev, uint32_t ty, const void* pinned_data, uint64_t n, uint64_t* out);
hj_status hj_tr_array_sharded(hj_comm* comm, uint32_t ty, const void* local_data, uint64_t n_global, uint64_t* out);
hj_status hj_tr_from_buffer_sharded(hj_comm* comm, hj_buffer* local_buf, uint32_t ty, uint64_t n_global, uint64_t* out);
/* *sharded = 1: the variable's buffer is the block [start, start + count) of its global extent;
 * *deferred = HJ_SHARD_DEFERRED: it is a scan result kept as (local scan, offset) — hj_tr_to_host adds
 * the offset, hj_tr_materialise adds it on the device; HJ_SHARD_SEGMENT: it is the rank's segment of a
 * compacted sequence (for a DynSize variable *count is then the rank's own count, read from the device).
 * hj_tr_to_host of a sharded variable addresses the BLOCK. */
hj_status hj_tr_var_shard(uint64_t v, int32_t* sharded, uint64_t* start, uint64_t* count, int32_t* deferred);
hj_status hj_tr_materialise(uint64_t v);
hj_status hj_tr_bop(uint32_t op /* hj_bop */, uint64_t a, uint64_t b, uint64_t* out);  /* :968-1046 */
hj_status hj_tr_uop(uint32_t op /* hj_uop */, uint64_t a, uint64_t* out);              /* :1048-1054 */
hj_status hj_tr_cast(uint64_t a, uint32_t ty, uint64_t* out);                     /* :1056-1068 */
hj_status hj_tr_bitcast(uint64_t a, uint32_t ty, uint64_t* out);                  /* :1070-1082 */
hj_status hj_tr_fma(uint64_t a, uint64_t b, uint64_t c, uint64_t* out);           /* :1335-1351 */
hj_status hj_tr_select(uint64_t true_val, uint64_t cond, uint64_t false_val, uint64_t* out); /* :1483-1504 */
hj_status hj_tr_extract(uint64_t a, uint32_t elem, uint64_t* out);                /* :1439-1459 */
hj_status hj_tr_extract_dyn(uint64_t a, uint64_t elem, uint64_t* out);            /* :1460-1477 */
hj_status hj_tr_composite(const uint64_t* refs, uint32_t n, uint64_t* out);       /* :675-693 */
hj_status hj_tr_vec(const uint64_t* refs, uint32_t n, uint64_t* out);             /* :717-737 */
hj_status hj_tr_mat(const uint64_t* columns, uint32_t n, uint64_t* out);         /* :734-756 */
hj_status hj_tr_arr(const uint64_t* refs, uint32_t n, uint64_t* out);             /* :694-712 */
hj_status hj_tr_gather(uint64_t src, uint64_t idx, uint64_t active, uint64_t* out); /* gather_if :1125-1164 */
hj_status hj_tr_scatter(uint64_t src, uint64_t dst, uint64_t idx, uint64_t active); /* :1167-1209 */
hj_status hj_tr_scatter_reduce(uint64_t src, uint64_t dst, uint64_t idx, uint64_t active, uint32_t op); /* :1210-1258 */
hj_status hj_tr_scatter_atomic(uint64_t src, uint64_t dst, uint64_t idx, uint64_t active, uint32_t op,
                               uint64_t* out);                                    /* :1259-1309 */
hj_status hj_tr_atomic_inc(uint64_t dst, uint64_t idx, uint64_t active, uint64_t* out); /* :1310-1334 */
hj_status hj_tr_prefix_sum(uint64_t a, int32_t inclusive, uint64_t* out);         /* :1623-1637 */
hj_status hj_tr_reduce(uint64_t a, uint32_t op /* hj_reduce_op */, uint64_t* out); /* :1641-1653 */
hj_status hj_tr_compress(uint64_t mask, uint64_t* out_count, uint64_t* out_index); /* :1595-1620 */
hj_status hj_tr_compress_dyn(uint64_t mask, uint64_t* out);                       /* :1583-1591 */
/* loop_start / if_start (:408-432, :466-490): state[0] is the bool condition; state_out
 * receives n new references to the state as seen inside the body */
hj_status hj_tr_scope_start(int32_t is_loop, const uint64_t* state, uint32_t n, uint64_t* out_scope,
                            uint64_t* state_out);
/* loop_end / if_end (:433-465, :492-521) */
hj_status hj_tr_scope_end(uint64_t scope, const uint64_t* state, uint32_t n, uint64_t* state_out);

hj_status hj_tr_schedule(uint64_t v);        /* VarRef::schedule, trace.rs:919-935 */
hj_status hj_tr_schedule_eval(void);         /* tr::schedule_eval, trace.rs:540-545 */
hj_status hj_tr_reset_schedule(void);        /* drops the calling thread's schedule */
hj_status hj_tr_var_size(uint64_t v, uint64_t* out); /* element count; DynSize reads the device count */
hj_status hj_tr_to_host(uint64_t v, uint64_t start_elem, uint64_t n_elem, void* dst); /* to_vec, :1404-1438 */

/* Report of Graph::launch_with (graph.rs:138-143) */
typedef struct {
    float aliasing_rate;
    double aliasing_duration_us;
    double backend_cpu_us;
    uint32_t n_passes;
    hj_pass_report* passes; /* caller-provided (>= n_passes entries) for per-pass GPU times, or NULL */
    uint32_t passes_capacity;
} hj_graph_report;

hj_status hj_tr_compile(hj_graph** out);     /* tr::compile, trace.rs:528-536 */
hj_status hj_tr_compile_fn(const uint64_t* inputs, uint32_t n_in, const uint64_t* outputs, uint32_t n_out,
                           hj_graph** out);  /* graph::compile with inputs/outputs, record.rs:168-193 */
hj_status hj_graph_retain(hj_graph* g);
hj_status hj_graph_release(hj_graph* g);
uint32_t hj_graph_n_passes(hj_graph* g);
uint32_t hj_graph_n_outputs(hj_graph* g);
/* The flat IR of kernel pass `pass` (valid while the graph is alive; *out = NULL for device-op
 * passes): inspect or pre-compile a graph's kernels with hj_ir_codegen / hj_ir_compile_cubin. */
hj_status hj_graph_pass_ir(hj_graph* g, uint32_t pass, const hj_ir** out);
/* Wire format of a compiled graph: passes with their kernel IR, resource table, inputs/outputs and
 * the contents of captured buffers (the reference keeps graphs in memory only, graph.rs:145-151;
 * together with the on-disk cubin cache a recorded function starts in a fresh process without
 * tracing or compiling).  *bytes_out is malloc'ed — free with hj_free_string((char*)bytes).
 * `dev` of hj_graph_deserialize may be NULL for graphs that capture no buffers. */
hj_status hj_graph_serialize(hj_graph* g, void** bytes_out, size_t* n_out);
hj_status hj_graph_deserialize(hj_device* dev, const void* bytes, size_t n, hj_graph** out);
/* `{:#?}` of the Graph, byte-identical to the reference's insta snapshots; free with hj_free_string */
hj_status hj_graph_debug_string(hj_graph* g, char** out);
/* Graph::launch / launch_with (graph.rs:180-400); outputs_out (may be NULL) receives
 * hj_graph_n_outputs new references */
hj_status hj_graph_launch(hj_graph* g, hj_device* dev, const uint64_t* inputs, uint32_t n_in,
                          uint64_t* outputs_out, hj_graph_report* report);
/* The same over arrays partitioned across the ranks of `comm` (may be NULL: then the communicator
 * of any sharded input / captured variable is used, and a graph without one runs on `dev` alone). */
hj_status hj_graph_launch_sharded(hj_graph* g, hj_device* dev, hj_comm* comm, const uint64_t* inputs,
                                  uint32_t n_in, uint64_t* outputs_out, hj_graph_report* report);

/* function cache of record() (record.rs:116-210): key = hash(function identity, input hashes) */
hj_status hj_fcache_get(uint64_t key, hj_graph** out); /* *out = NULL on a miss, retained on a hit */
hj_status hj_fcache_put(uint64_t key, hj_graph* g);
hj_status hj_fcache_clear(void);
uint64_t hj_fcache_size(void);

#ifdef __cplusplus
}
#endif
#endif /* HJ_H */
