// comm.cu — multi-GPU layer: one process per GPU, per-GPU partials combined over NCCL
// (NVLink 5 / NVSwitch).
//
// The reference has no multi-GPU code at all; this layer is what
// BASELINE.json's north star adds on top of the backend: arrays are partitioned into
// contiguous shards, each rank runs the single-GPU kernel on its shard, and only
//   * one scalar per rank      (reduce: partials;  scan / compress: shard totals / counts), or
//   * one histogram per rank   (scatter-reduce),
// crosses the fabric.  The payloads are tiny, so the exchange is an all-gather of `world`
// scalars followed by a fold IN RANK ORDER on every rank: deterministic for floats, and it
// covers the operators / types NCCL has no reduction for (and / or / xor, 16-bit integers).
// The scan and compress kernels take the cross-GPU carry as a device-resident seed / index
// base, so no second pass over the data is needed to apply it (scan.cu, compress.cu).
//
// NCCL is loaded with dlopen at first use (libnccl.so.2: the copy torch already mapped when
// running under torchrun, else the system one), so libhj_b200.so itself has no NCCL dependency.
#include <dlfcn.h>

#include <cstdlib>
#include <type_traits>
#include <vector>
#include <nccl.h>

#include "common.cuh"
#include "hj_internal.h"
#include "peer.cuh"

using hj::HJ_MAX_PEERS;

struct hj_comm {
    hj_device* dev = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    void* scratch = nullptr;  // [0,64): local scalar | [64, 64+8*world): gathered | then: seed / sum
    // peer-memory exchange (NVLink / NVSwitch): every rank's mailbox is mapped into every process
    // through CUDA IPC, so `world` scalars are all-gathered by direct remote stores + local polling
    void* mailbox = nullptr;              // own: 2 parities x world slots x 16 bytes
    void* peer_mailbox[HJ_MAX_PEERS] = {};  // [rank] = own mailbox, others IPC-mapped
    // the same IPC allocation also holds a 2-parity inbox for array payloads (privatised histograms):
    // peers PUSH (element, epoch) words into slot [their rank] over NVLink, the owner polls locally
    void* peer_arraybox[HJ_MAX_PEERS] = {};
    bool p2p = false;
    bool connected = true;    // false between hj_comm_create_local and hj_comm_connect
};

namespace hj {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("cannot load libnccl: ") + dlerror(); return; }
#define HJ_SYM(field, sym)                                                      \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym));  \
    if (!api.field) { api.error = std::string("libnccl lacks ") + sym; return; }
        HJ_SYM(GetUniqueId, "ncclGetUniqueId")
        HJ_SYM(CommInitRank, "ncclCommInitRank")
        HJ_SYM(CommDestroy, "ncclCommDestroy")
        HJ_SYM(AllGather, "ncclAllGather")
        HJ_SYM(AllReduce, "ncclAllReduce")
        HJ_SYM(Send, "ncclSend")
        HJ_SYM(Recv, "ncclRecv")
        HJ_SYM(GroupStart, "ncclGroupStart")
        HJ_SYM(GroupEnd, "ncclGroupEnd")
        HJ_SYM(GetErrorString, "ncclGetErrorString")
#undef HJ_SYM
    });
    return api;
}

#define HJ_NCCL(expr)                                                                        \
    do {                                                                                     \
        ncclResult_t _r = (expr);                                                            \
        if (_r != ncclSuccess)                                                               \
            return fail(HJ_ERR_NCCL, "%s failed: %s", #expr, nccl().GetErrorString(_r));     \
    } while (0)

hj_status need_nccl() {
    NcclApi& a = nccl();
    if (!a.error.empty()) return fail(HJ_ERR_NCCL, "%s", a.error.c_str());
    return HJ_OK;
}

__global__ void set_u32_kernel(uint32_t* p, uint32_t v) { *p = v; }

// seed[0] = sum of gathered[0 .. rank)  (exclusive scan of the per-rank totals at `rank`);
// total[0] = sum of gathered[0 .. world)
template <typename T>
__global__ void offsets_kernel(const T* gathered, int rank, int world, T* seed, T* total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        T s = (T)0, pre = (T)0;
        for (int q = 0; q < world; q++) {
            if (q == rank) pre = s;
            s = (T)(s + gathered[q]);
        }
        if (seed) seed[0] = pre;
        if (total) total[0] = s;
    }
}

// What sizes and places the DynSize kernels that run over this rank's segment of a sharded compaction:
// seg[0] = the rank's own count, seg[1] = the counts of the ranks before it (where the segment starts in
// the global compacted sequence — KernelOp::Index of those kernels, codegen.cpp: HJ_SIZE)
__global__ void segment_seed_kernel(const uint32_t* counts, int rank, uint32_t* seg) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t before = 0;
        for (int q = 0; q < rank; q++) before += counts[q];
        seg[0] = counts[rank];
        seg[1] = before;
    }
}

// dst[i] = fold over ranks (in rank order) of all[q * n + i]
template <typename T, int OP>
__global__ void fold_ranks_kernel(const T* all, size_t n, int world, T* dst) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T v = all[i];
    for (int q = 1; q < world; q++) {
        T o = all[(size_t)q * n + i];
        if (OP == HJ_REDUCE_OR) v |= o;
        else if (OP == HJ_REDUCE_AND) v &= o;
        else v ^= o;
    }
    dst[i] = v;
}

bool nccl_type(hj_type_kind ty, ncclDataType_t* out) {
    switch (ty) {
    case HJ_I8: *out = ncclInt8; return true;
    case HJ_U8: case HJ_BOOL: *out = ncclUint8; return true;
    case HJ_I32: *out = ncclInt32; return true;
    case HJ_U32: *out = ncclUint32; return true;
    case HJ_I64: *out = ncclInt64; return true;
    case HJ_U64: *out = ncclUint64; return true;
    case HJ_F32: *out = ncclFloat32; return true;
    case HJ_F64: *out = ncclFloat64; return true;
    default: return false;
    }
}

hj_status need_comm_nccl(hj_comm* c, const char* what) {
    if (!c->comm)
        return fail(HJ_ERR_NCCL, "%s needs NCCL, but this communicator was created without it (hj_comm_create_local)", what);
    return need_nccl();
}

// scratch layout: [0,64) local scalar | [64, 64+8*world) gathered | 64 bytes seed / sum | 64 bytes:
// the device-resident exchange epoch (u32) and the ticket of the multi-CTA exchange kernels (u32)
size_t scratch_bytes(int world) { return 64 + 8 * (size_t)world + 64 + 64; }
uint32_t* xepoch_slot(hj_comm* c) { return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(c->scratch) + 64 + 8 * (size_t)c->world + 64); }

char* local_slot(hj_comm* c) { return reinterpret_cast<char*>(c->scratch); }
char* gathered_slot(hj_comm* c) { return reinterpret_cast<char*>(c->scratch) + 64; }
char* extra_slot(hj_comm* c) { return reinterpret_cast<char*>(c->scratch) + 64 + 8 * (size_t)c->world; }

// ---- all-gather of one scalar per rank over peer memory ------------------------------------------
// Payloads on this path are one scalar per GPU, so the exchange is pure latency: an NCCL
// all-gather costs ~30-40 us per call on 8 GPUs (measured, profiles/r01_sharded.txt), more than the
// whole local reduction of a 2^27-element shard.  Instead every rank STORES its value straight
// into the mailbox of every peer (one 16-byte slot per source rank, mapped with CUDA IPC, the
// stores travel over NVLink / NVSwitch) and then polls its OWN mailbox until all `world` slots
// carry the current epoch.  A slot is two self-describing words, (low half | epoch) and
// (high half | epoch), so no ordering between the two stores is needed.  Two mailbox parities
// alternate: a rank can only be one exchange ahead of a peer (it needs the peer's value of the
// current exchange to finish it), so the slot it overwrites next is never one still being read.
__global__ void __launch_bounds__(32)
peer_allgather_kernel(PeerView pv, const void* __restrict__ local, int es, void* __restrict__ gathered) {
    const int lane = threadIdx.x;
    unsigned long long bits = 0;
    memcpy(&bits, local, es);  // es in {1, 2, 4, 8}
    const unsigned long long v = peer_allgather_warp(pv, bits, lane);
    if (lane < pv.world) memcpy(reinterpret_cast<char*>(gathered) + (size_t)lane * es, &v, es);
}

// ---- all-reduce of one small array per rank over peer memory (privatised histograms) -----------
// ncclAllReduce of 256 KiB costs ~50 us on 8 GPUs, twice the local histogram of a 2^25-key shard.
// Here the exchange is ONE kernel without any fence, flag or grid-wide wait: every thread owns four
// elements, PUSHES them into slot [its rank] of every peer's IPC-mapped inbox over NVLink /
// NVSwitch as self-validating words — each 8-byte word is (element, epoch), so a word is complete
// exactly when its epoch is the current one, however the stores were split or reordered on the
// way — then polls the `world - 1` slots of its OWN inbox (local memory, L1 bypassed) for the same
// four elements and folds them in rank order into dst: the result is bit-identical on every rank,
// floats included.  Half of the bytes on the wire are epochs; at 256 KiB per rank that is noise,
// the cost is one NVLink traversal.  Two inbox parities alternate with the epoch: a peer that is
// already one exchange ahead writes the other parity, and it cannot be two ahead because it needs
// this rank's words of the exchange in between.
constexpr size_t HJ_ARRAYBOX_OFFSET = 4096, HJ_ARRAYBOX_BYTES = 256 * 1024;
constexpr size_t HJ_ARRAYSLOT_BYTES = 2 * HJ_ARRAYBOX_BYTES;  // (element, epoch) pairs

template <typename T, int OP>
__device__ __forceinline__ T fold2(T a, T b) {
    if (OP == HJ_REDUCE_SUM) return (T)(a + b);
    if (OP == HJ_REDUCE_MAX) return a > b ? a : b;
    if (OP == HJ_REDUCE_MIN) return a < b ? a : b;
    if constexpr (!std::is_floating_point<T>::value) {
        if (OP == HJ_REDUCE_OR) return a | b;
        if (OP == HJ_REDUCE_AND) return a & b;
        if (OP == HJ_REDUCE_XOR) return a ^ b;
    }
    return a;
}

template <typename T, int OP>
__device__ __forceinline__ void array_allreduce_element(const ArrayPeerView& ax, uint32_t epoch, uint32_t i,
                                                        uint32_t* __restrict__ dst32, uint32_t n) {
    const int rank = ax.rank, world = ax.world;
    // a thread owns four elements; the last vector of an array whose length is not a multiple of four
    // (or whose base is not 16-byte aligned) is moved element by element, the padding travels as zeros
    const bool whole = 4 * i + 4 <= n && ((uintptr_t)dst32 & 15u) == 0;
    uint4 mine = make_uint4(0, 0, 0, 0);
    if (whole) {
        mine = reinterpret_cast<const uint4*>(dst32)[i];
    } else {
        uint32_t e[4] = {0, 0, 0, 0};
        for (uint32_t k = 0; k < 4 && 4 * i + k < n; k++) e[k] = dst32[4 * i + k];
        mine = make_uint4(e[0], e[1], e[2], e[3]);
    }
    const size_t slot_vecs = ax.slot_vecs;
    const size_t par = (size_t)(epoch & 1u) * ax.parity_vecs;
    const uint4 w0 = make_uint4(mine.x, epoch, mine.y, epoch), w1 = make_uint4(mine.z, epoch, mine.w, epoch);
    const uint4* own = ax.box[0];  // ax.box[rank] without indexing the parameter struct dynamically
#pragma unroll
    for (int q = 1; q < HJ_MAX_PEERS; q++)
        if (q == rank) own = ax.box[q];
    own += par;
#pragma unroll
    for (int q = 0; q < HJ_MAX_PEERS; q++)
        if (q < world && q != rank) {
            uint4* to = ax.box[q] + par + (size_t)rank * slot_vecs + 2 * (size_t)i;
            st_sys_v4(to, w0);
            st_sys_v4(to + 1, w1);
        }
    T acc[4];
    bool first = true;
#pragma unroll
    for (int q = 0; q < HJ_MAX_PEERS; q++)
        if (q < world) {  // rank order: the same association, hence the same bits, on every rank
            uint4 src = mine;
            if (q != rank) {
                const uint4* from = own + (size_t)q * slot_vecs + 2 * (size_t)i;
                uint4 a, b;
                unsigned ns = 20;
                while (true) {
                    a = ld_sys_v4(from);
                    b = ld_sys_v4(from + 1);
                    if (a.y == epoch && a.w == epoch && b.y == epoch && b.w == epoch) break;
                    __nanosleep(ns);
                    if (ns < 500) ns *= 2;
                }
                src = make_uint4(a.x, a.z, b.x, b.z);
            }
            T o[4];
            memcpy(o, &src, 16);
#pragma unroll
            for (int k = 0; k < 4; k++) acc[k] = first ? o[k] : fold2<T, OP>(acc[k], o[k]);
            first = false;
        }
    uint4 out;
    memcpy(&out, acc, 16);
    if (whole) {
        reinterpret_cast<uint4*>(dst32)[i] = out;
    } else {
        const uint32_t e[4] = {out.x, out.y, out.z, out.w};
        for (uint32_t k = 0; k < 4 && 4 * i + k < n; k++) dst32[4 * i + k] = e[k];
    }
}

template <typename T, int OP>
__global__ void __launch_bounds__(256)
array_allreduce_kernel(ArrayPeerView ax, uint32_t* __restrict__ dst32, uint32_t n) {
    static_assert(sizeof(T) == 4, "4-byte elements");
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) s_epoch = xepoch_begin(ax.xepoch);
    __syncthreads();
    const uint32_t epoch = s_epoch;
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (4 * i < n) array_allreduce_element<T, OP>(ax, epoch, i, dst32, n);
    array_exchange_commit(ax, epoch);  // the CTA that finishes last commits the epoch
}

PeerView peer_view(hj_comm* c) {
    PeerView pv;
    for (int i = 0; i < HJ_MAX_PEERS; i++) pv.box[i] = reinterpret_cast<unsigned long long*>(c->peer_mailbox[i]);
    pv.rank = c->rank;
    pv.world = c->world;
    pv.xepoch = xepoch_slot(c);
    return pv;
}

ArrayPeerView array_view(hj_comm* c) {
    ArrayPeerView ax;
    for (int r = 0; r < HJ_MAX_PEERS; r++) ax.box[r] = r < c->world ? reinterpret_cast<uint4*>(c->peer_arraybox[r]) : nullptr;
    ax.rank = c->rank;
    ax.world = c->world;
    ax.slot_vecs = (uint32_t)(HJ_ARRAYSLOT_BYTES / 16);
    ax.parity_vecs = (uint32_t)((size_t)c->world * HJ_ARRAYSLOT_BYTES / 16);
    ax.xepoch = xepoch_slot(c);
    ax.done = xepoch_slot(c) + 1;
    return ax;
}

// dst (n elements of a 4-byte type) = fold over ranks of their dst, through the peer inboxes
template <typename T>
hj_status peer_array_allreduce(hj_comm* c, hj_reduce_op op, void* dst, size_t n) {
    const uint32_t n_vec = (uint32_t)((n + 3) / 4);
    const ArrayPeerView ax = array_view(c);
    const unsigned grid = (n_vec + 255) / 256;
#define HJ_COMBINE(OP) array_allreduce_kernel<T, OP><<<grid, 256, 0, c->dev->stream>>>(ax, (uint32_t*)dst, (uint32_t)n)
    switch (op) {
    case HJ_REDUCE_SUM: HJ_COMBINE(HJ_REDUCE_SUM); break;
    case HJ_REDUCE_MAX: HJ_COMBINE(HJ_REDUCE_MAX); break;
    case HJ_REDUCE_MIN: HJ_COMBINE(HJ_REDUCE_MIN); break;
    case HJ_REDUCE_OR: HJ_COMBINE(HJ_REDUCE_OR); break;
    case HJ_REDUCE_AND: HJ_COMBINE(HJ_REDUCE_AND); break;
    default: HJ_COMBINE(HJ_REDUCE_XOR); break;
    }
#undef HJ_COMBINE
    return check_launch(c->dev, "array_allreduce_kernel");
}

// all-gather one element of `es` bytes per rank: local_slot -> gathered_slot
hj_status gather_scalars(hj_comm* c, size_t es) {
    if (c->p2p) {
        PeerView pv = peer_view(c);
        peer_allgather_kernel<<<1, 32, 0, c->dev->stream>>>(pv, local_slot(c), (int)es, gathered_slot(c));
        return check_launch(c->dev, "peer_allgather_kernel");
    }
    HJ_TRY(need_comm_nccl(c, "the all-gather fallback"));
    HJ_NCCL(nccl().AllGather(local_slot(c), gathered_slot(c), es, ncclUint8, c->comm, c->dev->stream));
    return HJ_OK;
}

// dst[i] += seed[0]: materialises a scan result that was left with a deferred seed
template <typename T>
__global__ void __launch_bounds__(256) apply_seed_kernel(T* __restrict__ dst, size_t n, const T* __restrict__ seed) {
    constexpr int VEC = 16 / sizeof(T);
    const T s = seed[0];
    const size_t nvec = n / VEC, stride = (size_t)gridDim.x * 256;
    uint4* v = reinterpret_cast<uint4*>(dst);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += stride) {
        uint4 raw = v[i];
        T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
        for (int j = 0; j < VEC; j++) e[j] = (T)(e[j] + s);
        v[i] = raw;
    }
    for (size_t i = nvec * VEC + (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) dst[i] = (T)(dst[i] + s);
}
template <typename T>
__global__ void __launch_bounds__(256) apply_seed_scalar_kernel(T* __restrict__ dst, size_t n, const T* __restrict__ seed) {
    const T s = seed[0];
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) dst[i] = (T)(dst[i] + s);
}
template <typename T>
hj_status run_apply_seed(hj_device* dev, size_t n, void* dst, const void* seed) {
    const size_t want = (n * sizeof(T) / 16 + 255) / 256, cap = (size_t)dev->sm_count * 16;
    const unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
    if (((uintptr_t)dst & 15u) == 0) apply_seed_kernel<T><<<grid, 256, 0, dev->stream>>>((T*)dst, n, (const T*)seed);
    else apply_seed_scalar_kernel<T><<<grid, 256, 0, dev->stream>>>((T*)dst, n, (const T*)seed);
    return check_launch(dev, "apply_seed_kernel");
}

// Peer mailboxes.  Every rank allocates one IPC-exportable block (scalar mailbox slots, then the
// array inbox), hands its cudaIpcMemHandle_t to every other rank — over NCCL (hj_comm_create) or
// through the host (hj_comm_create_local + hj_comm_connect: any transport, NCCL not needed) —
// and maps theirs.  Any failure leaves c->p2p false and the NCCL path in use.
size_t mailbox_bytes(int world) { return HJ_ARRAYBOX_OFFSET + 2 * (size_t)world * HJ_ARRAYSLOT_BYTES; }

bool alloc_mailbox(hj_comm* c, cudaIpcMemHandle_t* mine) {
    static_assert(2 * (size_t)HJ_MAX_PEERS * 16 <= HJ_ARRAYBOX_OFFSET, "mailbox slots overlap the array box");
    static_assert(sizeof(cudaIpcMemHandle_t) == HJ_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    const size_t box_bytes = mailbox_bytes(c->world);
    if (cudaMalloc(&c->mailbox, box_bytes) != cudaSuccess) { cudaGetLastError(); c->mailbox = nullptr; return false; }
    // epoch 0 = empty; completed before anybody can learn the handle
    if (cudaMemsetAsync(c->mailbox, 0, box_bytes, c->dev->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->dev->stream) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaIpcGetMemHandle(mine, c->mailbox) != cudaSuccess) { cudaGetLastError(); return false; }
    return true;
}

bool open_mailboxes(hj_comm* c, const cudaIpcMemHandle_t* all) {
    bool ok = true;
    for (int r = 0; ok && r < c->world; r++) {
        if (r == c->rank) { c->peer_mailbox[r] = c->mailbox; continue; }
        if (cudaIpcOpenMemHandle(&c->peer_mailbox[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            c->peer_mailbox[r] = nullptr;
            ok = false;
        }
    }
    return ok;
}

void finish_mailboxes(hj_comm* c, bool agreed) {
    c->p2p = agreed;
    for (int r = 0; c->p2p && r < c->world; r++)
        c->peer_arraybox[r] = reinterpret_cast<char*>(c->peer_mailbox[r]) + HJ_ARRAYBOX_OFFSET;
}

void setup_peer_mailboxes(hj_comm* c) {
    if (c->world < 2 || c->world > HJ_MAX_PEERS || getenv("HJ_NO_P2P")) return;
    cudaIpcMemHandle_t mine;
    bool ok = alloc_mailbox(c, &mine);
    // from here on every rank takes part in both collectives, whatever happened locally
    void* d_handles = nullptr;
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    if (cudaMalloc(&d_handles, hs * (c->world + 1)) != cudaSuccess) { cudaGetLastError(); return; }
    char* send = reinterpret_cast<char*>(d_handles) + hs * c->world;
    cudaMemcpyAsync(send, &mine, hs, cudaMemcpyHostToDevice, c->dev->stream);
    ok = nccl().AllGather(send, d_handles, hs, ncclUint8, c->comm, c->dev->stream) == ncclSuccess && ok;
    std::vector<cudaIpcMemHandle_t> all(c->world);
    ok = cudaMemcpyAsync(all.data(), d_handles, hs * c->world, cudaMemcpyDeviceToHost, c->dev->stream) == cudaSuccess && ok;
    ok = cudaStreamSynchronize(c->dev->stream) == cudaSuccess && ok;
    cudaFree(d_handles);
    ok = ok && open_mailboxes(c, all.data());
    // every rank must agree, or some would wait on mailboxes nobody writes: all-reduce the verdict
    int* d_ok = nullptr;
    if (cudaMalloc(&d_ok, sizeof(int)) != cudaSuccess) { cudaGetLastError(); return; }
    int h_ok = ok ? 1 : 0;
    cudaMemcpyAsync(d_ok, &h_ok, sizeof(int), cudaMemcpyHostToDevice, c->dev->stream);
    bool agreed = nccl().AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, c->comm, c->dev->stream) == ncclSuccess;
    cudaMemcpyAsync(&h_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, c->dev->stream);
    cudaStreamSynchronize(c->dev->stream);
    cudaFree(d_ok);
    finish_mailboxes(c, agreed && h_ok == 1);
}

}  // namespace
}  // namespace hj

namespace hj {
namespace {
// Local compaction with GLOBAL indices (index_base = global start of the shard); out_count[0] = global
// count, counts[q] = count of rank q (counts_out, else the communicator's scratch).  With peer memory the
// exchange of the counts runs inside the compaction kernel (compress.cu: CompressOp::finish).
hj_status sharded_compress(hj_comm* c, size_t n_local, uint32_t index_base, const uint8_t* mask, uint32_t* index_out,
                           uint32_t* out_count, uint32_t* counts_out, bool zero_tail, const uint32_t* size_buf = nullptr) {
    // size_buf (DynSize: only the first size_buf[0] mask elements of this rank count) takes the unfused path
    uint32_t* counts = counts_out ? counts_out : (uint32_t*)gathered_slot(c);
    if (c->world == 1) {
        HJ_TRY(launch_compress(c->dev, n_local, size_buf, out_count, mask, index_out, index_base, zero_tail));
        HJ_CUDA(cudaMemcpyAsync(counts, out_count, 4, cudaMemcpyDeviceToDevice, c->dev->stream));
        return HJ_OK;
    }
    if (!size_buf && c->p2p && compress_can_fuse_exchange(n_local, mask)) {
        PeerView pv = peer_view(c);
        static const bool zt_kernel = getenv("HJ_ZERO_TAIL_KERNEL") != nullptr;
        const bool zt_in_kernel = zero_tail && ((uintptr_t)index_out & 15u) == 0 && !zt_kernel;
        HJ_TRY(launch_compress(c->dev, n_local, nullptr, out_count, mask, index_out, index_base, zt_in_kernel, counts, &pv));
        if (zero_tail && !zt_in_kernel) return launch_compress_zero_tail(c->dev, index_out, counts + c->rank, n_local);
        return HJ_OK;
    }
    // small / misaligned shards, or no peer memory: compaction, then the exchange as its own step
    HJ_TRY(launch_compress(c->dev, n_local, size_buf, (uint32_t*)local_slot(c), mask, index_out, index_base, zero_tail));
    HJ_TRY(gather_scalars(c, 4));
    offsets_kernel<uint32_t><<<1, 32, 0, c->dev->stream>>>((const uint32_t*)gathered_slot(c), c->rank, c->world, nullptr,
                                                          out_count);
    HJ_TRY(check_launch(c->dev, "offsets_kernel"));
    if (counts_out)
        HJ_CUDA(cudaMemcpyAsync(counts_out, gathered_slot(c), 4 * (size_t)c->world, cudaMemcpyDeviceToDevice, c->dev->stream));
    return HJ_OK;
}
}  // namespace

namespace {
// seed[0] = sum of the shard totals of the ranks before this one, without a fused kernel: totals
// pass (+ exchange in its last CTA over peer memory, else all-gather + offsets kernel)
hj_status exclusive_offset_of_totals(hj_comm* c, hj_type_kind ty, size_t n_local, const void* src, void* seed) {
    const size_t es = type_size(ty);
    if (c->p2p) {
        PeerView pv = peer_view(c);
        return launch_reduce(c->dev, HJ_REDUCE_SUM, ty, n_local, src, seed, &pv, 2u);
    }
    HJ_TRY(launch_reduce(c->dev, HJ_REDUCE_SUM, ty, n_local, src, local_slot(c)));
    HJ_TRY(gather_scalars(c, es));
    switch (es) {
    case 1: offsets_kernel<uint8_t><<<1, 32, 0, c->dev->stream>>>((const uint8_t*)gathered_slot(c), c->rank, c->world, (uint8_t*)seed, nullptr); break;
    case 2: offsets_kernel<uint16_t><<<1, 32, 0, c->dev->stream>>>((const uint16_t*)gathered_slot(c), c->rank, c->world, (uint16_t*)seed, nullptr); break;
    case 4:
        if (ty == HJ_F32) offsets_kernel<float><<<1, 32, 0, c->dev->stream>>>((const float*)gathered_slot(c), c->rank, c->world, (float*)seed, nullptr);
        else offsets_kernel<uint32_t><<<1, 32, 0, c->dev->stream>>>((const uint32_t*)gathered_slot(c), c->rank, c->world, (uint32_t*)seed, nullptr);
        break;
    default:
        if (ty == HJ_F64) offsets_kernel<double><<<1, 32, 0, c->dev->stream>>>((const double*)gathered_slot(c), c->rank, c->world, (double*)seed, nullptr);
        else offsets_kernel<unsigned long long><<<1, 32, 0, c->dev->stream>>>((const unsigned long long*)gathered_slot(c), c->rank, c->world, (unsigned long long*)seed, nullptr);
        break;
    }
    return check_launch(c->dev, "offsets_kernel");
}
}  // namespace

}  // namespace hj

namespace hj {
hj_status sharded_compress_pass(hj_comm* c, size_t n_local, uint32_t index_base, hj_buffer* mask, hj_buffer* index_out,
                                hj_buffer* out_count, bool zero_tail, hj_buffer* local_count, hj_buffer* local_size) {
    HJ_REQUIRE(c->connected, "communicator is not connected yet (hj_comm_connect)");
    HJ_REQUIRE(!local_size || local_size->bytes >= 4, "sharded compress: the local size buffer is smaller than 4 bytes");
    HJ_REQUIRE(!local_count || local_count->bytes >= 8, "sharded compress: the segment's seed buffer is smaller than 8 bytes");
    DeviceGuard g(c->dev);
    HJ_TRY(sharded_compress(c, n_local, index_base, (const uint8_t*)mask->ptr, (uint32_t*)index_out->ptr,
                            (uint32_t*)out_count->ptr, nullptr, zero_tail, local_size ? (const uint32_t*)local_size->ptr : nullptr));
    // every path of sharded_compress leaves the per-rank counts in the communicator's scratch
    if (local_count) {
        segment_seed_kernel<<<1, 32, 0, c->dev->stream>>>((const uint32_t*)gathered_slot(c), c->rank, (uint32_t*)local_count->ptr);
        HJ_TRY(check_launch(c->dev, "segment_seed_kernel"));
    }
    return HJ_OK;
}
}  // namespace hj

namespace hj {
hj_device* comm_device(hj_comm* c) { return c->dev; }
}  // namespace hj

using namespace hj;

extern "C" {

hj_status hj_comm_unique_id(uint8_t out_id[HJ_UNIQUE_ID_BYTES]) {
    HJ_REQUIRE(out_id, "null argument");
    HJ_TRY(need_nccl());
    static_assert(sizeof(ncclUniqueId) == HJ_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    HJ_NCCL(nccl().GetUniqueId(&id));
    memcpy(out_id, &id, sizeof(id));
    return HJ_OK;
}

hj_status hj_comm_create(hj_device* dev, const uint8_t id[HJ_UNIQUE_ID_BYTES], int32_t rank, int32_t world,
                         hj_comm** out) {
    HJ_REQUIRE(dev && id && out, "null argument");
    HJ_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    HJ_TRY(need_nccl());
    DeviceGuard g(dev);
    auto c = new hj_comm();
    c->dev = dev;
    c->rank = rank;
    c->world = world;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, uid, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(HJ_ERR_NCCL, "ncclCommInitRank failed: %s", nccl().GetErrorString(r));
    }
    cudaError_t e = cudaMalloc(&c->scratch, scratch_bytes(world));
    if (e != cudaSuccess) {
        nccl().CommDestroy(c->comm);
        delete c;
        return fail(HJ_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    cudaMemsetAsync(c->scratch, 0, scratch_bytes(world), dev->stream);
    setup_peer_mailboxes(c);
    dev->rc.fetch_add(1);
    *out = c;
    return HJ_OK;
}

hj_status hj_comm_create_local(hj_device* dev, int32_t rank, int32_t world, hj_comm** out,
                               uint8_t handle_out[HJ_IPC_HANDLE_BYTES]) {
    HJ_REQUIRE(dev && out && handle_out, "null argument");
    HJ_REQUIRE(world >= 1 && world <= HJ_MAX_PEERS && rank >= 0 && rank < world, "bad rank %d / world %d (at most %d ranks)",
               rank, world, HJ_MAX_PEERS);
    DeviceGuard g(dev);
    auto c = new hj_comm();
    c->dev = dev;
    c->rank = rank;
    c->world = world;
    c->connected = world == 1;
    cudaError_t e = cudaMalloc(&c->scratch, scratch_bytes(world));
    if (e != cudaSuccess) {
        delete c;
        return fail(HJ_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    cudaMemsetAsync(c->scratch, 0, scratch_bytes(world), dev->stream);
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (world > 1 && !alloc_mailbox(c, &mine)) {
        if (c->mailbox) cudaFree(c->mailbox);
        cudaFree(c->scratch);
        delete c;
        return fail(HJ_ERR_CUDA, "cannot allocate / export the peer mailbox (CUDA IPC)");
    }
    memcpy(handle_out, &mine, sizeof(mine));
    dev->rc.fetch_add(1);
    *out = c;
    return HJ_OK;
}

hj_status hj_comm_connect(hj_comm* c, const uint8_t* handles) {
    HJ_REQUIRE(c && (handles || c->world == 1), "null argument");
    HJ_REQUIRE(!c->comm, "hj_comm_connect: this communicator was bootstrapped over NCCL");
    if (c->connected) return HJ_OK;
    DeviceGuard g(c->dev);
    std::vector<cudaIpcMemHandle_t> all(c->world);
    memcpy(all.data(), handles, sizeof(cudaIpcMemHandle_t) * (size_t)c->world);
    if (!open_mailboxes(c, all.data())) return fail(HJ_ERR_CUDA, "cudaIpcOpenMemHandle failed for a peer mailbox");
    finish_mailboxes(c, true);
    c->connected = true;
    return HJ_OK;
}

hj_status hj_comm_info(hj_comm* c, int32_t* rank, int32_t* world, int32_t* peer_memory, int32_t* has_nccl) {
    HJ_REQUIRE(c, "null comm");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (peer_memory) *peer_memory = c->p2p ? 1 : 0;
    if (has_nccl) *has_nccl = c->comm ? 1 : 0;
    return HJ_OK;
}

hj_status hj_comm_device(hj_comm* c, hj_device** out) {
    HJ_REQUIRE(c && out, "null argument");
    *out = c->dev;
    return HJ_OK;
}

hj_status hj_comm_destroy(hj_comm* c) {
    HJ_REQUIRE(c, "null comm");
    {
        DeviceGuard g(c->dev);
        cudaStreamSynchronize(c->dev->stream);
        for (int r = 0; r < c->world && r < HJ_MAX_PEERS; r++)
            if (r != c->rank && c->peer_mailbox[r]) cudaIpcCloseMemHandle(c->peer_mailbox[r]);
        if (c->comm) nccl().CommDestroy(c->comm);
        if (c->mailbox) cudaFree(c->mailbox);
        if (c->scratch) cudaFree(c->scratch);
    }
    c->dev->rc.fetch_sub(1);
    delete c;
    return HJ_OK;
}

#define HJ_COMM_READY(c) HJ_REQUIRE((c)->connected, "communicator is not connected yet (hj_comm_connect)")

hj_status hj_sharded_reduce(hj_comm* c, hj_reduce_op op, hj_type_kind ty, size_t n_local, hj_buffer* src,
                            hj_buffer* dst) {
    HJ_REQUIRE(c && src && dst, "hj_sharded_reduce: null argument");
    HJ_COMM_READY(c);
    size_t es = type_size(ty);
    HJ_REQUIRE(es && n_local >= 1 && n_local * es <= src->bytes && es <= dst->bytes, "hj_sharded_reduce: bad sizes");
    DeviceGuard g(c->dev);
    settle_all(c->dev, {src, dst});
    if (c->p2p && c->world > 1) {
        // ONE kernel: the CTA that finishes the local reduction sends the partial to every peer's
        // mailbox over NVLink, collects theirs and folds them in rank order (reduce.cu)
        PeerView pv = peer_view(c);
        return launch_reduce(c->dev, op, ty, n_local, src->ptr, dst->ptr, &pv, 1u);
    }
    // local partial -> all-gather -> fold in rank order with the same reduction kernel
    HJ_TRY(launch_reduce(c->dev, op, ty, n_local, src->ptr, local_slot(c)));
    if (c->world == 1) {
        HJ_CUDA(cudaMemcpyAsync(dst->ptr, local_slot(c), es, cudaMemcpyDeviceToDevice, c->dev->stream));
        return HJ_OK;
    }
    HJ_TRY(gather_scalars(c, es));
    return launch_reduce(c->dev, op, ty, (size_t)c->world, gathered_slot(c), dst->ptr);
}

hj_status hj_sharded_prefix_sum(hj_comm* c, hj_type_kind ty, size_t n_local, int32_t inclusive, hj_buffer* src,
                                hj_buffer* dst) {
    HJ_REQUIRE(c && src && dst, "hj_sharded_prefix_sum: null argument");
    HJ_COMM_READY(c);
    size_t es = type_size(ty);
    HJ_REQUIRE(es && n_local >= 1 && n_local * es <= src->bytes && n_local * es <= dst->bytes,
               "hj_sharded_prefix_sum: bad sizes");
    DeviceGuard g(c->dev);
    settle_all(c->dev, {src, dst});
    if (c->world == 1) return launch_prefix_sum(c->dev, ty, n_local, inclusive != 0, src->ptr, dst->ptr, nullptr);
    // MATERIALISED result: shard total (a read-only pass, sizeof(T) bytes/element, the exchange fused
    // into its last CTA) -> scan seeded with this rank's exclusive offset.  12 bytes/element for 4-byte
    // types; hj_sharded_prefix_sum_deferred is the 8-byte form.
    HJ_TRY(exclusive_offset_of_totals(c, ty, n_local, src->ptr, extra_slot(c)));
    return launch_prefix_sum(c->dev, ty, n_local, inclusive != 0, src->ptr, dst->ptr, extra_slot(c));
}

hj_status hj_sharded_prefix_sum_deferred(hj_comm* c, hj_type_kind ty, size_t n_local, int32_t inclusive, hj_buffer* src,
                                         hj_buffer* dst, hj_buffer* seed_out) {
    HJ_REQUIRE(c && src && dst && seed_out, "hj_sharded_prefix_sum_deferred: null argument");
    HJ_COMM_READY(c);
    size_t es = type_size(ty);
    HJ_REQUIRE(es && n_local >= 1 && n_local * es <= src->bytes && n_local * es <= dst->bytes && seed_out->bytes >= es,
               "hj_sharded_prefix_sum_deferred: bad sizes");
    DeviceGuard g(c->dev);
    settle_all(c->dev, {src, dst, seed_out});
    if (c->world == 1) {
        HJ_CUDA(cudaMemsetAsync(seed_out->ptr, 0, es, c->dev->stream));
        return launch_prefix_sum(c->dev, ty, n_local, inclusive != 0, src->ptr, dst->ptr, nullptr);
    }
    if (c->p2p && prefix_sum_can_fuse_exchange(ty, n_local, src->ptr, dst->ptr)) {
        // ONE kernel, 2 * sizeof(T) bytes/element: the local scan; the CTA that owns the last tile sends
        // the shard total to every peer's mailbox and leaves the sum of the lower ranks' totals in seed_out
        PeerView pv = peer_view(c);
        return launch_prefix_sum(c->dev, ty, n_local, inclusive != 0, src->ptr, dst->ptr, nullptr, seed_out->ptr, &pv);
    }
    // small / misaligned shards, or no peer memory: totals pass + exchange, then the unseeded scan
    HJ_TRY(exclusive_offset_of_totals(c, ty, n_local, src->ptr, seed_out->ptr));
    return launch_prefix_sum(c->dev, ty, n_local, inclusive != 0, src->ptr, dst->ptr, nullptr);
}

hj_status hj_apply_seed(hj_device* dev, hj_type_kind ty, size_t n, hj_buffer* buf, hj_buffer* seed) {
    HJ_REQUIRE(dev && buf && seed, "hj_apply_seed: null argument");
    const size_t es = type_size(ty);
    HJ_REQUIRE(es && n * es <= buf->bytes && seed->bytes >= es, "hj_apply_seed: bad sizes");
    if (n == 0) return HJ_OK;
    DeviceGuard g(dev);
    settle_all(dev, {buf, seed});
    switch (ty) {
    case HJ_I8: case HJ_U8: return run_apply_seed<uint8_t>(dev, n, buf->ptr, seed->ptr);
    case HJ_I16: case HJ_U16: return run_apply_seed<uint16_t>(dev, n, buf->ptr, seed->ptr);
    case HJ_I32: case HJ_U32: return run_apply_seed<uint32_t>(dev, n, buf->ptr, seed->ptr);
    case HJ_I64: case HJ_U64: return run_apply_seed<unsigned long long>(dev, n, buf->ptr, seed->ptr);
    case HJ_F32: return run_apply_seed<float>(dev, n, buf->ptr, seed->ptr);
    case HJ_F64: return run_apply_seed<double>(dev, n, buf->ptr, seed->ptr);
    default: return fail(HJ_ERR_UNSUPPORTED, "hj_apply_seed: unsupported element type %s", type_name(ty));
    }
}

hj_status hj_sharded_compress(hj_comm* c, size_t n_local, uint32_t index_base, hj_buffer* src_mask,
                              hj_buffer* index_out, hj_buffer* out_count, hj_buffer* counts_out) {
    HJ_REQUIRE(c && src_mask && index_out && out_count, "hj_sharded_compress: null argument");
    HJ_COMM_READY(c);
    HJ_REQUIRE(n_local >= 1 && n_local <= src_mask->bytes && n_local * 4 <= index_out->bytes && out_count->bytes >= 4,
               "hj_sharded_compress: bad sizes");
    HJ_REQUIRE(!counts_out || counts_out->bytes >= 4 * (size_t)c->world, "hj_sharded_compress: counts_out too small");
    DeviceGuard g(c->dev);
    settle_all(c->dev, {src_mask, index_out, out_count, counts_out});
    return sharded_compress(c, n_local, index_base, (const uint8_t*)src_mask->ptr, (uint32_t*)index_out->ptr,
                            (uint32_t*)out_count->ptr, counts_out ? (uint32_t*)counts_out->ptr : nullptr, false);
}

hj_status hj_sharded_scatter_reduce(hj_comm* c, hj_reduce_op op, hj_type_kind ty, size_t n_local, hj_buffer* idx,
                                    hj_buffer* src, uint64_t literal, hj_buffer* dst, size_t n_dst) {
    HJ_REQUIRE(c && idx && dst, "hj_sharded_scatter_reduce: null argument");
    HJ_COMM_READY(c);
    size_t es = type_size(ty);
    HJ_REQUIRE(es && n_local * 4 <= idx->bytes && n_dst * es <= dst->bytes && (!src || n_local * es <= src->bytes),
               "hj_sharded_scatter_reduce: bad sizes");
    DeviceGuard g(c->dev);
    settle_all(c->dev, {idx, src, dst});
    // privatised per GPU: every rank reduces its keys into its own copy of dst (which the
    // caller initialised with the operator's identity), then the copies are combined.  The
    // packed-16 histogram (BASELINE: 2^16 u32 bins) folds its private counters AND runs the
    // exchange over peer memory in one kernel (scatter.cu: hist_fold_exchange_kernel).
    bool exchanged = false;
    static const bool no_peer_array = getenv("HJ_NO_PEER_ARRAY") != nullptr;
    const bool small_array = c->p2p && c->world > 1 && es == 4 && n_dst * 4 <= HJ_ARRAYBOX_BYTES && !no_peer_array;
    if (small_array && op == HJ_REDUCE_SUM && (ty == HJ_U32 || ty == HJ_I32) && !src && n_local >= 1) {
        ArrayPeerView ax = array_view(c);
        // a rank whose shard is too small for the ring kernel runs the plain exchange kernel instead of
        // the fused one: same wire format, same (device-resident) epoch, so the ranks may differ
        HJ_TRY(launch_scatter_reduce(c->dev, op, ty, n_local, (const uint32_t*)idx->ptr, nullptr, literal, dst->ptr, n_dst, &ax,
                                     &exchanged));
        if (exchanged) return HJ_OK;
        return peer_array_allreduce<uint32_t>(c, op, dst->ptr, n_dst);
    } else {
        HJ_TRY(launch_scatter_reduce(c->dev, op, ty, n_local, (const uint32_t*)idx->ptr, src ? src->ptr : nullptr, literal,
                                     dst->ptr, n_dst));
    }
    if (c->world == 1) return HJ_OK;
    // small 4-byte arrays: exchange over peer memory
    const bool float_bits = ty == HJ_F32 && (op == HJ_REDUCE_OR || op == HJ_REDUCE_AND || op == HJ_REDUCE_XOR);
    if (small_array && ((uintptr_t)dst->ptr & 3u) == 0 && !float_bits) {
        if (ty == HJ_F32) return peer_array_allreduce<float>(c, op, dst->ptr, n_dst);
        if (ty == HJ_I32) return peer_array_allreduce<int32_t>(c, op, dst->ptr, n_dst);
        if (ty == HJ_U32) return peer_array_allreduce<uint32_t>(c, op, dst->ptr, n_dst);
    }
    HJ_TRY(need_comm_nccl(c, "hj_sharded_scatter_reduce of an array too large for the peer inbox"));
    ncclDataType_t dt;
    HJ_REQUIRE(nccl_type(ty, &dt), "hj_sharded_scatter_reduce: no NCCL type for %s", type_name(ty));
    if (op == HJ_REDUCE_SUM || op == HJ_REDUCE_MAX || op == HJ_REDUCE_MIN) {
        ncclRedOp_t rop = op == HJ_REDUCE_SUM ? ncclSum : op == HJ_REDUCE_MAX ? ncclMax : ncclMin;
        HJ_NCCL(nccl().AllReduce(dst->ptr, dst->ptr, n_dst, dt, rop, c->comm, c->dev->stream));
        return HJ_OK;
    }
    // and / or / xor: all-gather the copies and fold in rank order
    HJ_REQUIRE(es == 4 || es == 8, "hj_sharded_scatter_reduce: bitwise ops need a 4- or 8-byte type");
    void* all = nullptr;
    HJ_CUDA(cudaMallocAsync(&all, n_dst * es * (size_t)c->world, c->dev->stream));
    HJ_NCCL(nccl().AllGather(dst->ptr, all, n_dst * es, ncclUint8, c->comm, c->dev->stream));
    unsigned grid = (unsigned)((n_dst + 255) / 256);
#define HJ_FOLD(T, OP) fold_ranks_kernel<T, OP><<<grid, 256, 0, c->dev->stream>>>((const T*)all, n_dst, c->world, (T*)dst->ptr)
    if (es == 4) {
        if (op == HJ_REDUCE_OR) HJ_FOLD(uint32_t, HJ_REDUCE_OR);
        else if (op == HJ_REDUCE_AND) HJ_FOLD(uint32_t, HJ_REDUCE_AND);
        else HJ_FOLD(uint32_t, HJ_REDUCE_XOR);
    } else {
        if (op == HJ_REDUCE_OR) HJ_FOLD(unsigned long long, HJ_REDUCE_OR);
        else if (op == HJ_REDUCE_AND) HJ_FOLD(unsigned long long, HJ_REDUCE_AND);
        else HJ_FOLD(unsigned long long, HJ_REDUCE_XOR);
    }
#undef HJ_FOLD
    hj_status s = check_launch(c->dev, "fold_ranks_kernel");
    cudaFreeAsync(all, c->dev->stream);
    return s;
}

// Re-partition a sharded, compacted sequence evenly (the step after compress_dyn in a
// multi-GPU wavefront loop).  Rank q holds counts[q] elements; the global sequence is their
// concatenation in rank order.  Afterwards rank r holds the contiguous block
// [r*T/W + min(r, T%W), ...) of it (sizes differ by at most one, order preserved) in `dst`.
// The counts are read back to the host once (W x 4 bytes, the one synchronisation of this call: NCCL
// needs the slice sizes on the host, and the caller sizes its next launches from *new_count_host anyway); the payload moves GPU to GPU over NVLink as grouped ncclSend /
// ncclRecv of exactly the overlapping slices — every element crosses the fabric at most once
// and elements that stay on their rank are one device-to-device copy.
hj_status hj_sharded_rebalance(hj_comm* c, size_t elem_bytes, hj_buffer* src, hj_buffer* counts, hj_buffer* dst,
                               hj_buffer* out_count, uint64_t* new_count_host) {
    HJ_REQUIRE(c && src && counts && dst, "hj_sharded_rebalance: null argument");
    HJ_REQUIRE(elem_bytes >= 1 && elem_bytes <= 64, "hj_sharded_rebalance: bad element size");
    HJ_REQUIRE(counts->bytes >= 4 * (size_t)c->world, "hj_sharded_rebalance: counts too small");
    HJ_REQUIRE(!out_count || out_count->bytes >= 4, "hj_sharded_rebalance: out_count too small");
    DeviceGuard g(c->dev);
    settle_all(c->dev, {src, counts, dst, out_count});
    const int W = c->world, me = c->rank;
    std::vector<uint32_t> cnt(W);
    HJ_CUDA(cudaMemcpyAsync(cnt.data(), counts->ptr, 4 * (size_t)W, cudaMemcpyDeviceToHost, c->dev->stream));
    HJ_CUDA(cudaStreamSynchronize(c->dev->stream));
    std::vector<uint64_t> src0(W + 1, 0), dst0(W + 1, 0);  // source / target block boundaries
    for (int q = 0; q < W; q++) src0[q + 1] = src0[q] + cnt[q];
    const uint64_t T = src0[W], base = T / W, rem = T % W;
    for (int q = 0; q <= W; q++) dst0[q] = (uint64_t)q * base + ((uint64_t)q < rem ? (uint64_t)q : rem);
    const uint64_t mine = dst0[me + 1] - dst0[me];
    HJ_REQUIRE((uint64_t)cnt[me] * elem_bytes <= src->bytes, "hj_sharded_rebalance: counts[rank] exceeds src");
    HJ_REQUIRE(mine * elem_bytes <= dst->bytes, "hj_sharded_rebalance: dst holds %zu bytes, the balanced block needs %llu",
               dst->bytes, (unsigned long long)(mine * elem_bytes));
    if (W > 1) HJ_TRY(need_comm_nccl(c, "hj_sharded_rebalance"));
    auto overlap = [](uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1, uint64_t* lo, uint64_t* hi) {
        *lo = a0 > b0 ? a0 : b0;
        *hi = a1 < b1 ? a1 : b1;
        return *lo < *hi;
    };
    uint64_t lo, hi;
    if (overlap(src0[me], src0[me + 1], dst0[me], dst0[me + 1], &lo, &hi))  // what stays here
        HJ_CUDA(cudaMemcpyAsync((char*)dst->ptr + (lo - dst0[me]) * elem_bytes, (const char*)src->ptr + (lo - src0[me]) * elem_bytes,
                                (hi - lo) * elem_bytes, cudaMemcpyDeviceToDevice, c->dev->stream));
    if (W > 1) {
        HJ_NCCL(nccl().GroupStart());
        ncclResult_t r = ncclSuccess;
        for (int q = 0; q < W && r == ncclSuccess; q++) {
            if (q == me) continue;
            if (overlap(src0[me], src0[me + 1], dst0[q], dst0[q + 1], &lo, &hi))  // my elements in q's block
                r = nccl().Send((const char*)src->ptr + (lo - src0[me]) * elem_bytes, (hi - lo) * elem_bytes, ncclUint8, q, c->comm,
                                c->dev->stream);
            if (r == ncclSuccess && overlap(src0[q], src0[q + 1], dst0[me], dst0[me + 1], &lo, &hi))  // q's elements in mine
                r = nccl().Recv((char*)dst->ptr + (lo - dst0[me]) * elem_bytes, (hi - lo) * elem_bytes, ncclUint8, q, c->comm,
                                c->dev->stream);
        }
        const ncclResult_t e = nccl().GroupEnd();
        HJ_NCCL(r);
        HJ_NCCL(e);
    }
    if (out_count) {  // the value travels as a kernel argument: no second synchronisation for a stack variable
        set_u32_kernel<<<1, 1, 0, c->dev->stream>>>((uint32_t*)out_count->ptr, (uint32_t)mine);
        HJ_TRY(check_launch(c->dev, "set_u32_kernel"));
    }
    if (new_count_host) *new_count_host = mine;
    return HJ_OK;
}

}  // extern "C"
