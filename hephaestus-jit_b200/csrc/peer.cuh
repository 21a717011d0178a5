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
    unsigned long long* remote = pv.box[peer] + parity_off + (size_t)pv.rank * 2;
    st_sys_u64(remote, (bits << 32) | epoch);
    st_sys_u64(remote + 1, (bits & 0xffffffff00000000ull) | epoch);
}
__device__ __forceinline__ unsigned long long peer_wait(const PeerView& pv, uint32_t epoch, int peer) {
    const size_t parity_off = (size_t)(epoch & 1u) * pv.world * 2;
    const unsigned long long* mine = pv.box[pv.rank] + parity_off + (size_t)peer * 2;
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
    unsigned long long* remote = pv.box[peer] + parity_off + (size_t)pv.rank * 2;
    st_sys_u64(remote, (bits << 32) | epoch);
    st_sys_u64(remote + 1, (bits & 0xffffffff00000000ull) | epoch);
    const unsigned long long* mine = pv.box[pv.rank] + parity_off + (size_t)peer * 2;
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

}  // namespace hj
