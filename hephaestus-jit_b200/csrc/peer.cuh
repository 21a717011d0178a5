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
};

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
    uint4* box[HJ_MAX_PEERS];  // inbox of every rank at the current parity (box[rank] = own, local memory)
    int rank, world;
    uint32_t epoch;
    uint32_t slot_vecs;        // uint4 per slot: one uint4 carries two (element, epoch) pairs
};
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
__device__ __forceinline__ void array_pair_allreduce_add(const ArrayPeerView& ax, uint32_t pair, uint32_t a, uint32_t b,
                                                         uint32_t* out_a, uint32_t* out_b) {
    const uint4 w = make_uint4(a, ax.epoch, b, ax.epoch);
    const uint4* own = ax.box[0];
#pragma unroll
    for (int q = 0; q < HJ_MAX_PEERS; q++) {
        if (q == ax.rank) own = ax.box[q];
        if (q < ax.world && q != ax.rank) st_sys_v4(ax.box[q] + (size_t)ax.rank * ax.slot_vecs + pair, w);
    }
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
                    if (v.y == ax.epoch && v.w == ax.epoch) break;
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
__device__ __forceinline__ unsigned long long peer_allgather_warp(const PeerView& pv, uint32_t epoch,
                                                                  unsigned long long bits, int lane) {
    unsigned long long v = 0;
    if (lane < pv.world) v = peer_exchange(pv, epoch, bits, lane);
    return v;
}

}  // namespace hj
