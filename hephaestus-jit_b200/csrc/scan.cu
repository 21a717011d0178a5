// scan.cu — single-pass inclusive/exclusive prefix sum (decoupled look-back) for sm_100a.
//
// Replaces builtin::prefix_sum::prefix_sum_large
// (hephaestus-jit/src/backend/vulkan/builtin/prefix_sum.rs:31-162 +
// kernels/prefix_sum_large.glsl + prefix_sum_large_init.glsl).  Differences by design:
//   * tile = 256 threads x NLOADS 128-bit vectors (16 KiB for 4/8-byte types, 8x the
//     reference's 2048-item partition) so that enough bytes are in flight per SM for HBM3e;
//   * vectors are consumed in the order they are loaded (vector-striped layout), so there is
//     no shared-memory transpose: per-vector sums are scanned with warp shuffles, the
//     NLOADS x 8 warp totals by one warp;
//   * no separate init dispatch (epoch-tagged status words, lookback.cuh);
//   * the tail is masked in the kernel; nothing is read or written beyond n (reference D7);
//   * true exclusive and inclusive variants (reference D10), correct for f32/u64/f64 (D3);
//   * `seed` adds a device-resident offset to every output — the cross-GPU carry of the
//     sharded scan — at no extra pass.
// Algorithmic bytes: 2 * sizeof(T) per element (read once, write once); HBM-bound.
#include <type_traits>

#include "hj_internal.h"
#include "lookback.cuh"

namespace hj {
namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

template <typename T, typename P, int NLOADS, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const T* __restrict__ src, T* __restrict__ dst, size_t n, const T* __restrict__ seed,
            LookbackView lb, int vec_ok) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int TILE = SCAN_THREADS * NLOADS * VEC;
    static_assert(NLOADS * SCAN_WARPS <= 32, "warp totals must fit one warp");
    __shared__ uint32_t s_tile;
    __shared__ P s_warp[NLOADS * SCAN_WARPS];
    __shared__ P s_prefix;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        uint32_t t = atomicAdd(lb.ticket, 1u);
        if (t == gridDim.x - 1) *lb.ticket = 0;  // last ticket: re-arm for the next launch
        s_tile = t;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t base = (size_t)tile * TILE;
    const bool full = vec_ok && base + TILE <= n;

    // ---- load: vector (i, tid) sits at element offset (i*THREADS + tid)*VEC of the tile
    P x[NLOADS][VEC];
    if (full) {
        uint4 raw[NLOADS];
        const uint4* vsrc = reinterpret_cast<const uint4*>(src + base);
#pragma unroll
        for (int i = 0; i < NLOADS; i++) raw[i] = ld_stream_v4(vsrc + i * SCAN_THREADS + tid);
#pragma unroll
        for (int i = 0; i < NLOADS; i++) {
            const T* e = reinterpret_cast<const T*>(&raw[i]);
#pragma unroll
            for (int j = 0; j < VEC; j++) x[i][j] = (P)e[j];
        }
    } else {
#pragma unroll
        for (int i = 0; i < NLOADS; i++)
#pragma unroll
            for (int j = 0; j < VEC; j++) {
                size_t e = base + (size_t)(i * SCAN_THREADS + tid) * VEC + j;
                x[i][j] = e < n ? (P)src[e] : (P)0;
            }
    }

    // ---- per-vector sums, warp scans, warp totals
    P excl_in_warp[NLOADS];
#pragma unroll
    for (int i = 0; i < NLOADS; i++) {
        P s = x[i][0];
#pragma unroll
        for (int j = 1; j < VEC; j++) s = (P)(s + x[i][j]);
        P inc = warp_inclusive_sum(s);
        if (lane == 31) s_warp[i * SCAN_WARPS + warp] = inc;
        P up = shfl_up(inc, 1);
        excl_in_warp[i] = lane == 0 ? (P)0 : up;
    }
    __syncthreads();

    // ---- warp 0: scan the NLOADS*WARPS totals, then resolve the tile prefix by look-back
    if (warp == 0) {
        constexpr int NT = NLOADS * SCAN_WARPS;
        P v = lane < NT ? s_warp[lane] : (P)0;
        P inc = warp_inclusive_sum(v);
        P aggregate = shfl_idx(inc, 31);
        P up = shfl_up(inc, 1);
        if (lane < NT) s_warp[lane] = lane == 0 ? (P)0 : up;
        P exclusive;
        if (tile == 0) {
            exclusive = seed ? (P)seed[0] : (P)0;
            if (lane == 0) tile_publish<P>(lb, 0, TILE_INCLUSIVE, (P)(exclusive + aggregate));
        } else {
            if (lane == 0) tile_publish<P>(lb, tile, TILE_AGGREGATE, aggregate);
            exclusive = tile_lookback<P>(lb, tile);
            if (lane == 0) tile_publish<P>(lb, tile, TILE_INCLUSIVE, (P)(exclusive + aggregate));
        }
        if (lane == 0) s_prefix = exclusive;
    }
    __syncthreads();
    const P tile_prefix = s_prefix;

    // ---- finish: running sum inside each vector, store
#pragma unroll
    for (int i = 0; i < NLOADS; i++) {
        P run = (P)(tile_prefix + (P)(s_warp[i * SCAN_WARPS + warp] + excl_in_warp[i]));
        T out[VEC];
#pragma unroll
        for (int j = 0; j < VEC; j++) {
            if (INCLUSIVE) { run = (P)(run + x[i][j]); out[j] = (T)run; }
            else { out[j] = (T)run; run = (P)(run + x[i][j]); }
        }
        if (full) {
            st_stream_v4(reinterpret_cast<uint4*>(dst + base) + i * SCAN_THREADS + tid,
                         *reinterpret_cast<const uint4*>(out));
        } else {
#pragma unroll
            for (int j = 0; j < VEC; j++) {
                size_t e = base + (size_t)(i * SCAN_THREADS + tid) * VEC + j;
                if (e < n) dst[e] = out[j];
            }
        }
    }
}

template <typename T, typename P, int NLOADS>
hj_status run(hj_device* dev, size_t n, bool inclusive, const void* src, void* dst, const void* seed) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr size_t TILE = (size_t)SCAN_THREADS * NLOADS * VEC;
    size_t n_tiles = (n + TILE - 1) / TILE;
    HJ_REQUIRE(n_tiles < (1ull << 31), "prefix_sum: too many tiles");
    HJ_TRY(ensure_lookback_scratch(dev, n_tiles));
    uint32_t epoch;
    HJ_TRY(next_epoch(dev, &epoch));
    LookbackView lb = lookback_view(dev->lookback.base, dev->lookback.capacity_tiles, epoch);
    int vec_ok = (((uintptr_t)src | (uintptr_t)dst) & 15u) == 0;
    if (inclusive)
        scan_kernel<T, P, NLOADS, true><<<(unsigned)n_tiles, SCAN_THREADS, 0, dev->stream>>>(
            (const T*)src, (T*)dst, n, (const T*)seed, lb, vec_ok);
    else
        scan_kernel<T, P, NLOADS, false><<<(unsigned)n_tiles, SCAN_THREADS, 0, dev->stream>>>(
            (const T*)src, (T*)dst, n, (const T*)seed, lb, vec_ok);
    return check_launch(dev, "scan_kernel");
}

}  // namespace

hj_status launch_prefix_sum(hj_device* dev, hj_type_kind ty, size_t n, bool inclusive, const void* src,
                            void* dst, const void* seed) {
    // Integer sums wrap, so signed types run on the unsigned kernel of the same width.
    switch (ty) {
    case HJ_I8: case HJ_U8: return run<uint8_t, uint32_t, 1>(dev, n, inclusive, src, dst, seed);
    case HJ_I16: case HJ_U16: return run<uint16_t, uint32_t, 2>(dev, n, inclusive, src, dst, seed);
    case HJ_I32: case HJ_U32: return run<uint32_t, uint32_t, 4>(dev, n, inclusive, src, dst, seed);
    case HJ_I64: case HJ_U64: return run<uint64_t, uint64_t, 4>(dev, n, inclusive, src, dst, seed);
    case HJ_F32: return run<float, float, 4>(dev, n, inclusive, src, dst, seed);
    case HJ_F64: return run<double, double, 4>(dev, n, inclusive, src, dst, seed);
    default:
        return fail(HJ_ERR_UNSUPPORTED, "prefix_sum: unsupported element type %s", type_name(ty));
    }
}

}  // namespace hj
