// lookback.cuh — decoupled look-back (Merrill & Garland) tile-status protocol shared by the
// scan and compress kernels.
//
// The reference (kernels/prefix_sum_large.glsl:268-327, compress_large.glsl:150-216) keeps
// one u64 per partition, `(value << 32) | flag`, for EVERY type — which silently corrupts
// carries of f32/u64/f64 scans — and clears the whole array with an extra
// dispatch before every scan (prefix_sum_large_init.glsl).  Here:
//   * 4-byte prefixes use the same packed 64-bit word (one relaxed 64-bit store/load is
//     atomic), the value is BIT-cast, not value-converted;
//   * 8-byte prefixes use a flag array plus separate aggregate / inclusive arrays with
//     release/acquire ordering, so nothing is truncated;
//   * the flag carries a 30-bit launch epoch, so the scratch is never cleared between
//     launches: a word written by an earlier launch simply reads as "not ready";
//   * tiles take their index from an atomic ticket (like the reference,
//     prefix_sum_large.glsl:177-186), so a tile only ever waits on tiles that already run.
#pragma once
#include "common.cuh"

namespace hj {

enum : uint32_t { TILE_INVALID = 0, TILE_AGGREGATE = 1, TILE_INCLUSIVE = 2 };

// Scratch layout (bytes): [0,64) ticket counter | words: max_tiles * 8 | agg: max_tiles * 8 |
// incl: max_tiles * 8.  4-byte prefixes only touch `words`.
struct LookbackView {
    unsigned* ticket;
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

template <typename P>
__device__ __forceinline__ void tile_publish(const LookbackView& lb, uint32_t tile, uint32_t state,
                                             P value) {
    const uint32_t flag = (lb.epoch << 2) | state;
    if constexpr (sizeof(P) == 4) {
        uint32_t bits = *reinterpret_cast<uint32_t*>(&value);
        st_relaxed_u64(lb.words + tile, ((unsigned long long)bits << 32) | flag);
    } else {
        unsigned long long bits = *reinterpret_cast<unsigned long long*>(&value);
        unsigned long long* slot = (state == TILE_AGGREGATE ? lb.agg : lb.incl) + tile;
        st_relaxed_u64(slot, bits);
        st_release_u32(reinterpret_cast<unsigned*>(lb.words + tile), flag);
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
        uint32_t flag = ld_acquire_u32(reinterpret_cast<const unsigned*>(lb.words + tile));
        if ((flag >> 2) != lb.epoch) return TILE_INVALID;
        uint32_t state = flag & 3u;
        unsigned long long bits = ld_relaxed_u64((state == TILE_AGGREGATE ? lb.agg : lb.incl) + tile);
        *value = *reinterpret_cast<P*>(&bits);
        return state;
    }
}

// Executed by one full warp of tile `tile` (> 0): returns (in every lane) the sum of the
// aggregates of all predecessor tiles.  Lane l inspects tile - 1 - l - 32*k in round k.
template <typename P>
__device__ __forceinline__ P tile_lookback(const LookbackView& lb, uint32_t tile) {
    const int lane = lane_id();
    P prefix = (P)0;
    long long idx = (long long)tile - 1 - lane;
    while (true) {
        P v = (P)0;
        uint32_t state = TILE_INCLUSIVE;  // virtual tiles before tile 0: inclusive, value 0
        if (idx >= 0) {
            do {
                state = tile_read<P>(lb, (uint32_t)idx, &v);
            } while (state == TILE_INVALID);
        }
        // every lane has a valid predecessor now
        unsigned incl_mask = __ballot_sync(0xffffffffu, state == TILE_INCLUSIVE);
        if (incl_mask) {
            int first = __ffs(incl_mask) - 1;  // nearest predecessor whose full prefix is known
            P c = lane <= first ? v : (P)0;
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) c = (P)(c + shfl_xor(c, m));
            return (P)(prefix + c);
        }
        P c = v;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) c = (P)(c + shfl_xor(c, m));
        prefix = (P)(prefix + c);
        idx -= 32;
    }
}

}  // namespace hj
