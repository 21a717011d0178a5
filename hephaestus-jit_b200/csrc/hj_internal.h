// hj_internal.h — shared internals of libhj_b200.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/hj.h"

namespace hj {

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const char* fmt, ...);
hj_status fail(hj_status code, const char* fmt, ...);

#define HJ_CUDA(expr)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            cudaGetLastError();                                                           \
            return ::hj::fail(_e == cudaErrorMemoryAllocation ? HJ_ERR_OOM : HJ_ERR_CUDA, \
                              "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                              __FILE__, __LINE__);                                        \
        }                                                                                 \
    } while (0)

#define HJ_TRY(expr)                        \
    do {                                    \
        hj_status _s = (expr);              \
        if (_s != HJ_OK) return _s;         \
    } while (0)

#define HJ_REQUIRE(cond, ...)                                  \
    do {                                                       \
        if (!(cond)) return ::hj::fail(HJ_ERR_INVALID, __VA_ARGS__); \
    } while (0)

size_t type_size(hj_type_kind ty);
const char* type_name(hj_type_kind ty);
const char* reduce_op_name(hj_reduce_op op);

struct KernelCache;  // jit.cpp
struct GraphCache;   // graph_exec.cpp

// Persistent per-device scratch of the decoupled-look-back kernels (scan, compress).
// Status words carry an epoch so the buffer never has to be cleared between launches.
struct LookbackScratch {
    void* base = nullptr;   // layout: lookback.cuh (ticket | status words | aggregates | inclusives)
    size_t bytes = 0;
    size_t capacity_tiles = 0;
    uint32_t epoch = 0;     // epochs consumed by the launches enqueued since the scratch was last cleared
    uint64_t generation = 0;  // bumped whenever any per-device scratch is reallocated (captured graphs hold raw pointers)
};

// A buffer that is still being written (or read) chunk by chunk on one of the device's side streams:
// an asynchronous upload (hj_buffer_create_from_host_async) or the output of a kernel pass that was
// launched chunk-wise behind such an upload (jit.cpp: kernel_launch_streamed).  `done[c]` fires when
// chunk c — elements [first[c], first[c] + count[c]) — is complete, `done.back()` when everything is.
// A schedule-less progress (first.empty()) is a plain fence: the buffer is being READ on a side stream.
// Everything that touches the buffer on the device stream calls settle() first.
struct AsyncProgress {
    std::vector<size_t> first, count;
    std::vector<cudaEvent_t> done;
    ~AsyncProgress() {
        for (cudaEvent_t e : done)
            if (e) cudaEventDestroy(e);
    }
};

}  // namespace hj

struct hj_device {
    std::atomic<int> rc{1};
    int ordinal = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // the stream work is enqueued on (own or caller's)
    cudaMemPool_t pool = nullptr;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t total_mem = 0, l2_bytes = 0;
    int max_smem_optin = 0;
    std::recursive_mutex mu;  // serialises enqueue + scratch use (Device: Send + Sync in the reference);
                              // recursive: a graph capture holds it across the passes it enqueues
    hj::LookbackScratch lookback;
    void* reduce_scratch = nullptr;  // partials + ticket of the single-pass reduction
    size_t reduce_scratch_bytes = 0;
    void* hist_scratch = nullptr;    // per-CTA-group private histograms of the privatised scatter-reduce
    size_t hist_scratch_bytes = 0;
    std::atomic<uint64_t> launches{0};
    std::atomic<uint64_t> n_alloc{0}, n_free{0};
    hj::KernelCache* kcache = nullptr;
    hj::GraphCache* gcache = nullptr;  // captured CUDA graphs of hj_execute_graph_cached
    // side streams of the chunk-wise asynchronous path (upload | kernels | download), created on first use
    cudaStream_t side_up = nullptr, side_kernel = nullptr, side_down = nullptr;
    std::atomic<int> async_live{0};  // buffers that carry an AsyncProgress (settle() is free while this is 0)
};

struct hj_buffer {
    std::atomic<int> rc{1};
    hj_device* dev = nullptr;
    void* ptr = nullptr;
    size_t bytes = 0;
    bool owned = true;
    std::shared_ptr<hj::AsyncProgress> progress;  // guarded by the device lock
    uint32_t progress_elem_bytes = 0;
};

namespace hj {

// RAII guard: selects the device and holds its enqueue lock.
struct DeviceGuard {
    hj_device* dev;
    std::unique_lock<std::recursive_mutex> lock;
    explicit DeviceGuard(hj_device* d) : dev(d), lock(d->mu) { cudaSetDevice(d->ordinal); }
};

// ---- chunk-wise asynchronous buffers (runtime.cpp) ------------------------------------------
hj_status ensure_side_streams(hj_device* dev);  // device lock held
void attach_progress(hj_buffer* b, std::shared_ptr<AsyncProgress> p, uint32_t elem_bytes);  // device lock held
// Orders the device stream behind whatever a side stream still does with `b` and forgets the progress.
void settle_locked(hj_buffer* b);
inline void settle(hj_buffer* b) {
    if (!b || b->dev->async_live.load(std::memory_order_acquire) == 0) return;
    DeviceGuard g(b->dev);
    settle_locked(b);
}
inline void settle_all(hj_device* dev, std::initializer_list<hj_buffer*> bs) {  // device lock held
    if (dev->async_live.load(std::memory_order_acquire) == 0) return;
    for (hj_buffer* b : bs) settle_locked(b);
}
// first / count (in elements) of the chunks `n` elements are moved in: small chunks at both ends (the fill
// and the drain of the pipeline overlap nothing), full ones in between
void chunk_schedule(size_t n, size_t chunk_elems, size_t align_elems, std::vector<size_t>* first, std::vector<size_t>* count);

// jit.cpp: launches a kernel pass chunk by chunk behind asynchronous uploads; *done = false when not applicable
hj_status kernel_launch_streamed(hj_device* dev, hj_kernel* k, size_t size, hj_buffer* const* buffers, uint32_t n_buffers,
                                 bool* done);

// jit.cpp: hj_kernel_launch with slot i bound to buffers[i]->ptr - shift_bytes[i] (shift_bytes may be NULL)
hj_status kernel_launch_shifted(hj_device* dev, hj_kernel* k, size_t size, hj_buffer* size_buf, hj_buffer* const* buffers,
                                uint32_t n_buffers, uint32_t index_base, const uint64_t* shift_bytes);

// host_stream.cu: a PrefixSum pass over a source that is still arriving; *done = false when not applicable
hj_status prefix_sum_arriving(hj_device* dev, hj_type_kind ty, size_t n, bool inclusive, hj_buffer* src, hj_buffer* dst,
                              bool* done);

// Grow-only scratch helpers (called with the device lock held).
hj_status ensure_reduce_scratch(hj_device* dev, size_t bytes);
hj_status ensure_hist_scratch(hj_device* dev, size_t bytes);
hj_status ensure_lookback_scratch(hj_device* dev, size_t n_tiles);
// Accounts for `n` launches of epoch-consuming kernels (clears the scratch before the device-side
// epoch could wrap).
hj_status count_epoch(hj_device* dev, uint32_t n = 1);

// ---- kernel launchers (defined in the .cu files; device lock held by the caller) ---------
struct PeerView;  // peer.cuh
// `peers` / `mode` (optional): fold the result with the partials of the other ranks over peer memory
// inside the same kernel (comm.cu) — mode 1: all ranks (sharded reduce), mode 2: the ranks before
// this one (the seed of a sharded scan)
hj_status launch_reduce(hj_device* dev, hj_reduce_op op, hj_type_kind ty, size_t n,
                        const void* src, void* dst, const PeerView* peers = nullptr, uint32_t mode = 0);
// `seed_out` / `peers` (optional, ring kernel only — see prefix_sum_can_fuse_exchange):
// the kernel also exchanges the shard totals over peer memory and writes this rank's exclusive
// offset to seed_out[0] (the DEFERRED seed of a sharded scan, comm.cu)
hj_status launch_prefix_sum(hj_device* dev, hj_type_kind ty, size_t n, bool inclusive,
                            const void* src, void* dst, const void* seed, void* seed_out = nullptr,
                            const PeerView* peers = nullptr);
bool prefix_sum_can_fuse_exchange(hj_type_kind ty, size_t n, const void* src, const void* dst);
hj_status launch_compress_zero_tail(hj_device* dev, uint32_t* index_out, const uint32_t* count, size_t n);
// zero_tail: also leave index_out[count .. n) zeroed (in the ring kernel where it can, else by
// launch_compress_zero_tail afterwards)
// `counts_out` / `peers` (optional, ring kernel only — see compress_can_fuse_exchange): the
// kernel also exchanges the per-rank counts over peer memory; out_count[0] then holds the GLOBAL
// count and counts_out[q] (may be NULL) the count of rank q
hj_status launch_compress(hj_device* dev, size_t n, const uint32_t* size_buf, uint32_t* out_count,
                          const uint8_t* mask, uint32_t* index_out, uint32_t index_base, bool zero_tail = false,
                          uint32_t* counts_out = nullptr, const PeerView* peers = nullptr);
bool compress_can_fuse_exchange(size_t n, const uint8_t* mask);
struct ArrayPeerView;  // peer.cuh
// `ax` (optional): on the packed-16 histogram path the fold kernel also all-reduces the bins over
// peer memory; *exchanged tells the caller whether that happened (else it combines the copies itself)
hj_status launch_scatter_reduce(hj_device* dev, hj_reduce_op op, hj_type_kind ty, size_t n,
                                const uint32_t* idx, const void* src, uint64_t literal, void* dst,
                                size_t n_dst, const ArrayPeerView* ax = nullptr, bool* exchanged = nullptr);
hj_status launch_gather(hj_device* dev, size_t elem_bytes, size_t n, const void* src,
                        const uint32_t* idx, void* dst);
hj_status launch_fill(hj_device* dev, void* dst, size_t n, size_t elem_bytes, uint64_t pattern);

// comm.cu internals used by the sharded pass interpreter (graph_exec.cpp)
hj_device* comm_device(hj_comm* c);
// `local_count` (optional, >= 8 bytes): receives this rank's own count and the counts of the ranks before it
// (what sizes the DynSize kernels that run over the rank's segment, and their KernelOp::Index)
// `local_size` (optional): DynSize — only the first local_size[0] mask elements of this rank are compacted
hj_status sharded_compress_pass(hj_comm* c, size_t n_local, uint32_t index_base, hj_buffer* mask, hj_buffer* index_out,
                                hj_buffer* out_count, bool zero_tail, hj_buffer* local_count = nullptr,
                                hj_buffer* local_size = nullptr);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel, size) instead of on every
// launch (a driver call of a few microseconds on the relaunch path).  Device lock held by the caller.
hj_status ensure_dynamic_smem(hj_device* dev, const void* kernel, size_t smem);

// Launch with programmatic stream serialization (PDL): the kernel may be scheduled while the kernel in
// front of it on the stream is still draining — its CTAs run their prologue (shared-memory carve-out,
// mbarrier init, bin zeroing) and then block in pdl_wait() (common.cuh) until that kernel has completed
// and its memory is visible.  ONLY for kernels that call pdl_wait() before their first global access.
// HJ_NO_PDL=1 launches them fully serialised.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    static const bool no_pdl = getenv("HJ_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

inline hj_status check_launch(hj_device* dev, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(HJ_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    dev->launches.fetch_add(1, std::memory_order_relaxed);
    return HJ_OK;
}

}  // namespace hj
