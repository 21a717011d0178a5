// lookback.cuh — decoupled look-back (Merrill & Garland) tile-status protocol shared by the
// scan and compress kernels.
//
// The reference (kernels/prefix_sum_large.glsl:268-327, compress_large.glsl:150-216) keeps
// one u64 per partition, `(value << 32) | flag`, for EVERY type — which silently corrupts
// carries of f32/u64/f64 scans — and clears the whole array with an extra
// dispatch before every scan (prefix_sum_large_init.glsl).  Here:
//   * 4-byte prefixes use the same packed 64-bit word (one relaxed 64-bit store/load is
//     atomic), the value is BIT-cast, not value-converted;
//   * 8-byte prefixes use two such words (low half + flag, high half + flag) that a reader
//     only accepts when both flags agree, so nothing is truncated and nothing needs ordering;
//   * the flag carries a 30-bit launch epoch, so the scratch is never cleared between
//     launches: a word written by an earlier launch simply reads as "not ready".  The epoch
//     lives in the scratch itself (next to the ticket): every CTA reads it before its first
//     ticket draw and the CTA that makes the launch's last draw advances it.  No launch
//     parameter changes from one launch to the next, so the kernels can be replayed from a
//     captured CUDA graph (graph_exec.cpp);
//   * tiles take their index from an atomic ticket (like the reference,
//     prefix_sum_large.glsl:177-186), so a tile only ever waits on tiles that already run.
#pragma once
#include "common.cuh"

namespace hj {

enum : uint32_t { TILE_INVALID = 0, TILE_AGGREGATE = 1, TILE_INCLUSIVE = 2 };

// Scratch layout (bytes): [0,4) ticket counter | [4,8) epoch of the previous launch | words:
// max_tiles * 8 | agg: max_tiles * 8 | incl: max_tiles * 8.  4-byte prefixes only touch `words`.
struct LookbackView {
    unsigned* ticket;  // ticket[1] = epoch counter
    unsigned long long* words;
    unsigned long long* agg;
    unsigned long long* incl;
    uint32_t epoch;
};
static inline size_t lookback_bytes(size_t n_tiles) { return 64 + n_tiles * 24; }
static inline LookbackView lookback_view(void* base, size_t n_tiles, uint32_t epoch) {
    LookbackView v;
    char* p = reinterpret_cast<char*>(base);
    v.ticket = reinterpret_cast<unsigned*>(p);
    v.words = reinterpret_cast<unsigned long long*>(p + 64);
    v.agg = v.words + n_tiles;
    v.incl = v.agg + n_tiles;
    v.epoch = epoch;
    return v;
}
constexpr uint32_t EPOCH_MASK = 0x3fffffffu;
// Epoch of this launch; to be read by ONE thread per CTA BEFORE the CTA's first ticket draw (acquire:
// the draw cannot move ahead of it), so it is read before the launch's last draw, whose CTA calls
// epoch_advance.
__device__ __forceinline__ uint32_t epoch_begin(const LookbackView& lb) {
    return (ld_acquire_u32(lb.ticket + 1) + 1u) & EPOCH_MASK;
}
__device__ __forceinline__ void epoch_advance(const LookbackView& lb, uint32_t epoch) { lb.ticket[1] = epoch; }

template <typename P>
__device__ __forceinline__ void tile_publish(const LookbackView& lb, uint32_t tile, uint32_t state,
                                             P value) {
    const uint32_t flag = (lb.epoch << 2) | state;
    if constexpr (sizeof(P) == 4) {
        uint32_t bits = *reinterpret_cast<uint32_t*>(&value);
        st_relaxed_u64(lb.words + tile, ((unsigned long long)bits << 32) | flag);
    } else {
        // two self-describing words: (low half, flag) and (high half, flag).  A reader accepts
        // the pair only when both flags agree, so no store ordering is needed and both halves
        // travel in one round trip.
        unsigned long long bits = *reinterpret_cast<unsigned long long*>(&value);
        st_relaxed_u64(lb.words + tile, (bits << 32) | flag);
        st_relaxed_u64(lb.agg + tile, (bits & 0xffffffff00000000ull) | flag);
    }
}

// Returns the state of `tile` for this launch's epoch (TILE_INVALID if not yet written).
template <typename P>
__device__ __forceinline__ uint32_t tile_read(const LookbackView& lb, uint32_t tile, P* value) {
    if constexpr (sizeof(P) == 4) {
        unsigned long long w = ld_relaxed_u64(lb.words + tile);
        uint32_t flag = (uint32_t)w;
        if ((flag >> 2) != lb.epoch) return TILE_INVALID;
        uint32_t bits = (uint32_t)(w >> 32);
        *value = *reinterpret_cast<P*>(&bits);
        return flag & 3u;
    } else {
        unsigned long long w0 = ld_relaxed_u64(lb.words + tile);
        unsigned long long w1 = ld_relaxed_u64(lb.agg + tile);
        uint32_t flag = (uint32_t)w0;
        if ((flag >> 2) != lb.epoch || (uint32_t)w1 != flag) return TILE_INVALID;  // absent or torn
        unsigned long long bits = (w0 >> 32) | (w1 & 0xffffffff00000000ull);
        *value = *reinterpret_cast<P*>(&bits);
        return flag & 3u;
    }
}

// Executed by one full warp of tile `tile` (> 0): returns (in every lane) the sum of the
// aggregates of all predecessor tiles.
//
// Window width matters on B200: with ~900 tiles in flight the tiles of a wave are
// contemporaries, so a tile has to walk back over every in-flight predecessor until it meets
// one whose inclusive prefix is published, and the "inclusive frontier" advances by one
// window per L2 round trip.  With the classic 32-wide window (one status word per lane, as in
// the reference, prefix_sum_large.glsl:281-309) that caps the whole kernel at
// 32 tiles x tile bytes / round trip (measured: 2.8 TB/s for 32 KiB of traffic per tile).
// Here every lane inspects M consecutive predecessors per round (M independent loads in
// flight), i.e. a 32*M-wide window; lane l covers tiles win - l*M - m, m = 0..M-1.
template <typename P, int M = 4>
__device__ __forceinline__ P tile_lookback(const LookbackView& lb, uint32_t tile) {
    const int lane = lane_id();
    P prefix = (P)0;
    long long win = (long long)tile - 1;  // nearest tile of the current window
    // Tiles publish roughly in ticket order, so first wait (ONE polled word per tile, with
    // back-off) until the nearest predecessor has published anything; only then read windows.
    // Polling whole windows from ~600 resident tiles at once floods the L2 slices that hold
    // the frontier's status lines and delays the very stores everybody is waiting for.
    if (lane == 0) {
        P dummy;
        unsigned ns = 32;
        while (tile_read<P>(lb, tile - 1, &dummy) == TILE_INVALID) {
            __nanosleep(ns);
            if (ns < 256) ns *= 2;
        }
    }
    __syncwarp();
    while (true) {
        P v[M];
        uint32_t st[M];
#pragma unroll
        for (int m = 0; m < M; m++) {
            long long idx = win - (long long)lane * M - m;
            v[m] = (P)0;
            st[m] = TILE_INCLUSIVE;  // virtual tiles before tile 0: inclusive, value 0
            if (idx >= 0) st[m] = tile_read<P>(lb, (uint32_t)idx, &v[m]);
        }
        // fold this lane's tiles nearest -> farthest: stop at the first inclusive (done) or the
        // first not-yet-published one (blocked)
        P sum = (P)0;
        int kind = 0;  // 0: all aggregates, 1: reached an inclusive prefix, 2: blocked
#pragma unroll
        for (int m = 0; m < M; m++) {
            if (kind == 0) {
                if (st[m] == TILE_INVALID) kind = 2;
                else {
                    sum = (P)(sum + v[m]);
                    if (st[m] == TILE_INCLUSIVE) kind = 1;
                }
            }
        }
        const unsigned incl_mask = __ballot_sync(0xffffffffu, kind == 1);
        const unsigned blocked_mask = __ballot_sync(0xffffffffu, kind == 2);
        const int first_incl = incl_mask ? __ffs(incl_mask) - 1 : 32;
        const int first_blocked = blocked_mask ? __ffs(blocked_mask) - 1 : 32;
        if (first_incl < first_blocked) {  // everything up to the nearest inclusive is known
            P c = lane <= first_incl ? sum : (P)0;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) c = (P)(c + shfl_xor(c, s));
            return (P)(prefix + c);
        }
        if (first_blocked == 32) {  // a full window of aggregates: accumulate, step back
            P c = sum;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) c = (P)(c + shfl_xor(c, s));
            prefix = (P)(prefix + c);
            win -= 32 * M;
        }
        else {
            // a needed predecessor has not published yet -> poll the same window again, later
            __nanosleep(128);
        }
    }
}

}  // namespace hj
