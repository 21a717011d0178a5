// compress.cu — stream compaction (mask -> ascending indices + count) for sm_100a.
//
// Replaces builtin::compress::compress_large
// (hephaestus-jit/src/backend/vulkan/builtin/compress.rs:157-283 +
// kernels/compress_large.glsl:76-229).  Same contract: index_out[0..count) receives the
// positions of the set mask bytes in ascending order, out_count[0] the count, entries at and
// beyond count are not touched.  Differences by design:
//   * tile = 256 threads x 2 x 16 mask bytes = 8192 elements (4x the reference partition);
//   * each 16-byte vector is turned into a 16-bit lane mask with SIMD-in-word compares and
//     a multiply-gather, counted with popc — no per-byte scan;
//   * the selected indices are first compacted into shared memory and then written with
//     fully coalesced 128-byte warp stores.  The reference's per-thread scattered stores
//     (compress_large.glsl:224-228) touch one 32-byte sector per 4-byte index;
//   * the tail (and a device-resident DynSize count) is masked in the kernel (reference D4);
//   * `index_base` is added to every index: the shard's global offset on multi-GPU runs.
// Algorithmic bytes: n (mask) + 4 * count (indices); HBM-bound.
#include "hj_internal.h"
#include "lookback.cuh"

namespace hj {
namespace {

constexpr int CMP_THREADS = 256;
constexpr int CMP_WARPS = CMP_THREADS / 32;
constexpr int CMP_NLOADS = 2;
constexpr int CMP_TILE = CMP_THREADS * CMP_NLOADS * 16;

// 4 mask bytes -> 4-bit mask of the non-zero ones (bit k = byte k).
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t w) {
    uint32_t y = __vcmpne4(w, 0u) & 0x01010101u;  // bit 8k = (byte k != 0)
    // multiply-gather: 2^21 + 2^14 + 2^7 + 1 moves bit 8k to bit 21+k without collisions
    return (y * 0x00204081u >> 21) & 0xFu;
}

__global__ void __launch_bounds__(CMP_THREADS)
compress_kernel(const uint8_t* __restrict__ mask, size_t n, const uint32_t* __restrict__ size_buf,
                uint32_t* __restrict__ out_count, uint32_t* __restrict__ index_out,
                uint32_t index_base, LookbackView lb, int vec_ok) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[CMP_NLOADS * CMP_WARPS];
    __shared__ uint32_t s_prefix, s_total;
    __shared__ uint32_t s_idx[CMP_TILE];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        uint32_t t = atomicAdd(lb.ticket, 1u);
        if (t == gridDim.x - 1) *lb.ticket = 0;
        s_tile = t;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t base = (size_t)tile * CMP_TILE;
    size_t n_eff = n;
    if (size_buf) {  // DynSize: device-resident element count (graph.rs:503-508)
        size_t dyn = size_buf[0];
        n_eff = dyn < n ? dyn : n;
    }
    const bool full = vec_ok && base + CMP_TILE <= n_eff;

    uint32_t m[CMP_NLOADS];
    if (full) {
        uint4 raw[CMP_NLOADS];
        const uint4* vsrc = reinterpret_cast<const uint4*>(mask + base);
#pragma unroll
        for (int i = 0; i < CMP_NLOADS; i++) raw[i] = ld_stream_v4(vsrc + i * CMP_THREADS + tid);
#pragma unroll
        for (int i = 0; i < CMP_NLOADS; i++)
            m[i] = nonzero_nibble(raw[i].x) | (nonzero_nibble(raw[i].y) << 4) |
                   (nonzero_nibble(raw[i].z) << 8) | (nonzero_nibble(raw[i].w) << 12);
    } else {
#pragma unroll
        for (int i = 0; i < CMP_NLOADS; i++) {
            uint32_t bits = 0;
            size_t e0 = base + (size_t)(i * CMP_THREADS + tid) * 16;
            if (e0 < n_eff) {
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (e0 + j < n_eff && mask[e0 + j] != 0) bits |= 1u << j;
            }
            m[i] = bits;
        }
    }

    uint32_t excl_in_warp[CMP_NLOADS];
#pragma unroll
    for (int i = 0; i < CMP_NLOADS; i++) {
        uint32_t c = __popc(m[i]);
        uint32_t inc = warp_inclusive_sum(c);
        if (lane == 31) s_warp[i * CMP_WARPS + warp] = inc;
        excl_in_warp[i] = inc - c;
    }
    __syncthreads();

    if (warp == 0) {
        constexpr int NT = CMP_NLOADS * CMP_WARPS;
        uint32_t v = lane < NT ? s_warp[lane] : 0u;
        uint32_t inc = warp_inclusive_sum(v);
        uint32_t aggregate = __shfl_sync(0xffffffffu, inc, 31);
        if (lane < NT) s_warp[lane] = inc - v;
        uint32_t exclusive = 0;
        if (tile == 0) {
            if (lane == 0) tile_publish<uint32_t>(lb, 0, TILE_INCLUSIVE, aggregate);
        } else {
            if (lane == 0) tile_publish<uint32_t>(lb, tile, TILE_AGGREGATE, aggregate);
            exclusive = tile_lookback<uint32_t>(lb, tile);
            if (lane == 0) tile_publish<uint32_t>(lb, tile, TILE_INCLUSIVE, exclusive + aggregate);
        }
        if (lane == 0) {
            s_prefix = exclusive;
            s_total = aggregate;
            // compress_large.glsl:214-216: the last partition publishes the count
            if (tile == gridDim.x - 1) out_count[0] = exclusive + aggregate;
        }
    }

    // compact this tile's indices into shared memory (ranks are tile-local)
    // (s_warp is rewritten by warp 0 above, so wait for it first)
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CMP_NLOADS; i++) {
        uint32_t r = s_warp[i * CMP_WARPS + warp] + excl_in_warp[i];
        uint32_t first = index_base + (uint32_t)base + (uint32_t)(i * CMP_THREADS + tid) * 16u;
        uint32_t bits = m[i];
        while (bits) {
            int b = __ffs(bits) - 1;
            bits &= bits - 1;
            s_idx[r++] = first + b;
        }
    }
    __syncthreads();
    const uint32_t total = s_total;
    uint32_t* out = index_out + s_prefix;
    for (uint32_t k = tid; k < total; k += CMP_THREADS) out[k] = s_idx[k];
}

}  // namespace

hj_status launch_compress(hj_device* dev, size_t n, const uint32_t* size_buf, uint32_t* out_count,
                          const uint8_t* mask, uint32_t* index_out, uint32_t index_base) {
    HJ_REQUIRE(n <= 0xffffffffull, "compress: n does not fit the u32 index type");
    size_t n_tiles = (n + CMP_TILE - 1) / CMP_TILE;
    HJ_TRY(ensure_lookback_scratch(dev, n_tiles));
    uint32_t epoch;
    HJ_TRY(next_epoch(dev, &epoch));
    LookbackView lb = lookback_view(dev->lookback.base, dev->lookback.capacity_tiles, epoch);
    int vec_ok = ((uintptr_t)mask & 15u) == 0;
    compress_kernel<<<(unsigned)n_tiles, CMP_THREADS, 0, dev->stream>>>(mask, n, size_buf, out_count,
                                                                       index_out, index_base, lb, vec_ok);
    return check_launch(dev, "compress_kernel");
}

}  // namespace hj
