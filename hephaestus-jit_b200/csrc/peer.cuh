// peer.cuh — all-gather of one scalar per rank over peer memory (NVLink / NVSwitch).
//
// Every rank's mailbox (2 parities x world slots x 16 bytes) is mapped into every process with
// CUDA IPC (comm.cu).  A rank STORES its value straight into slot [its rank] of every peer's
// mailbox and then polls its OWN mailbox until all `world` slots carry the current epoch.  A slot
// is two self-describing words, (low half | epoch) and (high half | epoch): no ordering between
// the two stores is needed.  Two parities alternate: a rank can only be one exchange ahead of a
// peer (it needs that peer's value of the current exchange to finish it), so the slot it
// overwrites next is never one that is still being read.
#pragma once
#include "common.cuh"

namespace hj {

constexpr int HJ_MAX_PEERS = 16;

struct PeerView {
    unsigned long long* box[HJ_MAX_PEERS];
    int rank, world;
    // The exchange epoch lives ON THE DEVICE (one counter per communicator, in its scratch): a kernel
    // that runs an exchange reads it, uses the next value and commits it when the exchange is done.
    // Every rank performs the same sequence of exchanges, so the counters advance in lock step — and
    // no launch parameter changes from one launch to the next, which is what lets a whole sharded
    // pass list be replayed from a captured CUDA graph (graph_exec.cpp).
    uint32_t* xepoch;
};

// exchange epochs are 31-bit and never 0 (0 is what a cleared mailbox holds)
__device__ __forceinline__ uint32_t xepoch_next(uint32_t v) {
    v = (v + 1u) & 0x7fffffffu;
    return v ? v : 1u;
}
__device__ __forceinline__ uint32_t xepoch_begin(const uint32_t* counter) {
    return xepoch_next(*reinterpret_cast<const volatile uint32_t*>(counter));
}

// pv.box[peer] without indexing the kernel-parameter struct dynamically (which would make every
// thread copy the whole struct to local memory in the kernel prologue): a select chain over
// constant-bank loads.
__device__ __forceinline__ unsigned long long* peer_box(const PeerView& pv, int peer) {
    unsigned long long* p = pv.box[0];
#pragma unroll
    for (int q = 1; q < HJ_MAX_PEERS; q++)
        if (peer == q) p = pv.box[q];
    return p;
}

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// The two halves of an exchange, for callers that split them over different threads (a thread that
// only polls has no remote store outstanding when it fences afterwards).
__device__ __forceinline__ void peer_send(const PeerView& pv, uint32_t epoch, unsigned long long bits, int peer) {
    const size_t parity_off = (size_t)(epoch & 1u) * pv.world * 2;
    unsigned long long* remote = peer_box(pv, peer) + parity_off + (size_t)pv.rank * 2;
    st_sys_u64(remote, (bits << 32) | epoch);
    st_sys_u64(remote + 1, (bits & 0xffffffff00000000ull) | epoch);
}
__device__ __forceinline__ unsigned long long peer_wait(const PeerView& pv, uint32_t epoch, int peer) {
    const size_t parity_off = (size_t)(epoch & 1u) * pv.world * 2;
    const unsigned long long* mine = peer_box(pv, pv.rank) + parity_off + (size_t)peer * 2;
    unsigned long long w0, w1;
    unsigned ns = 20;
    while (true) {
        w0 = ld_sys_u64(mine);
        w1 = ld_sys_u64(mine + 1);
        if ((uint32_t)w0 == epoch && (uint32_t)w1 == epoch) break;
        __nanosleep(ns);
        if (ns < 1000) ns *= 2;
    }
    return (w0 >> 32) | (w1 & 0xffffffff00000000ull);
}

// Executed by threads 0 .. world-1 of one CTA: thread `peer` sends `bits` (this rank's value) to
// rank `peer` and returns the value rank `peer` sent here.
__device__ __forceinline__ unsigned long long peer_exchange(const PeerView& pv, uint32_t epoch, unsigned long long bits,
                                                            int peer) {
    const size_t parity_off = (size_t)(epoch & 1u) * pv.world * 2;
    unsigned long long* remote = peer_box(pv, peer) + parity_off + (size_t)pv.rank * 2;
    st_sys_u64(remote, (bits << 32) | epoch);
    st_sys_u64(remote + 1, (bits & 0xffffffff00000000ull) | epoch);
    const unsigned long long* mine = peer_box(pv, pv.rank) + parity_off + (size_t)peer * 2;
    unsigned long long w0, w1;
    unsigned ns = 20;
    while (true) {
        w0 = ld_sys_u64(mine);
        w1 = ld_sys_u64(mine + 1);
        if ((uint32_t)w0 == epoch && (uint32_t)w1 == epoch) break;
        __nanosleep(ns);
        if (ns < 1000) ns *= 2;
    }
    return (w0 >> 32) | (w1 & 0xffffffff00000000ull);
}

// ---- small-array exchange (privatised histograms): every rank owns an IPC-mapped inbox of `world`
// slots; a rank PUSHES its elements as self-validating 8-byte words (element, epoch) into slot
// [its rank] of every peer's inbox and polls its own inbox for the peers' words (comm.cu).
struct ArrayPeerView {
    uint4* box[HJ_MAX_PEERS];  // inbox of every rank, parity 0 (box[rank] = own, local memory)
    int rank, world;
    uint32_t slot_vecs;        // uint4 per slot: one uint4 carries two (element, epoch) pairs
    uint32_t parity_vecs;      // distance between the two parities of an inbox, in uint4
    uint32_t* xepoch;          // device-resident exchange counter (see PeerView)
    uint32_t* done;            // ticket of the multi-CTA exchange kernels: the last CTA commits the epoch
};
// Called by every thread of a CTA of a multi-CTA exchange kernel, after its own part of the exchange
// (contains a __syncthreads): the CTA that finishes last commits the epoch — by then every CTA has read it.
__device__ __forceinline__ void array_exchange_commit(const ArrayPeerView& ax, uint32_t epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(ax.done, 1u);
        if (t == gridDim.x * gridDim.y - 1) {
            *ax.done = 0;
            *ax.xepoch = epoch;
        }
    }
}
__device__ __forceinline__ void st_sys_v4(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_sys_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// Pair `pair` of the array (elements 2*pair, 2*pair+1): send (a, b) to every peer, return in
// (*out_a, *out_b) the sum over all ranks in rank order (u32, wrapping) — bit-identical on every rank.
__device__ __forceinline__ void array_pair_allreduce_add(const ArrayPeerView& ax, uint32_t epoch, uint32_t pair, uint32_t a,
                                                         uint32_t b, uint32_t* out_a, uint32_t* out_b) {
    const uint4 w = make_uint4(a, epoch, b, epoch);
    const size_t par = (size_t)(epoch & 1u) * ax.parity_vecs;
    const uint4* own = ax.box[0];
#pragma unroll
    for (int q = 0; q < HJ_MAX_PEERS; q++) {
        if (q == ax.rank) own = ax.box[q];
        if (q < ax.world && q != ax.rank) st_sys_v4(ax.box[q] + par + (size_t)ax.rank * ax.slot_vecs + pair, w);
    }
    own += par;
    uint32_t sa = 0, sb = 0;
#pragma unroll
    for (int q = 0; q < HJ_MAX_PEERS; q++)
        if (q < ax.world) {
            uint4 v = w;
            if (q != ax.rank) {
                const uint4* from = own + (size_t)q * ax.slot_vecs + pair;
                unsigned ns = 20;
                while (true) {
                    v = ld_sys_v4(from);
                    if (v.y == epoch && v.w == epoch) break;
                    __nanosleep(ns);
                    if (ns < 500) ns *= 2;
                }
            }
            sa += v.x;
            sb += v.z;
        }
    *out_a = sa;
    *out_b = sb;
}

// Executed by one full warp (e.g. the prefix warp of the ring kernels, in their `finish`): all-gather
// of one value per rank — lane q (< world) returns the bits rank q contributed, other lanes 0.
// Draws the next exchange epoch from the device-resident counter and commits it afterwards.
__device__ __forceinline__ unsigned long long peer_allgather_warp(const PeerView& pv, unsigned long long bits, int lane) {
    uint32_t epoch = lane == 0 ? xepoch_begin(pv.xepoch) : 0u;
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    unsigned long long v = 0;
    if (lane < pv.world) v = peer_exchange(pv, epoch, bits, lane);
    __syncwarp();
    if (lane == 0) *pv.xepoch = epoch;
    return v;
}

}  // namespace hj
