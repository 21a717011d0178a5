// compress.cu — stream compaction (mask -> ascending indices + count) for sm_100a.
//
// Replaces builtin::compress::compress_large
// (hephaestus-jit/src/backend/vulkan/builtin/compress.rs:157-283 +
// kernels/compress_large.glsl:76-229).  Same contract: index_out[0..count) receives the
// positions of the set mask bytes in ascending order, out_count[0] the count, entries at and
// beyond count are not touched.  Differences by design:
//   * two sweeps per 32 KiB super-tile, like the scan kernel (scan.cu explains why the
//     register-resident single sweep stalls on B200): sweep 1 only COUNTS the set bytes of
//     each warp's contiguous 4 KiB segment; block aggregate -> publish -> 128-wide look-back;
//     sweep 2 re-reads the mask rows (1 byte/element, L2-resident) and emits the indices;
//   * each 16-byte vector is turned into a 16-bit lane mask with SIMD-in-word compares and
//     a multiply-gather, counted with popc — no per-byte scan;
//   * the indices of a 512-element warp row are compacted in a per-warp shared-memory stage
//     and written with coalesced 128-byte warp stores; sweep 2 needs no block barrier.  The
//     reference's per-thread scattered stores (compress_large.glsl:224-228) touch one 32-byte
//     sector per 4-byte index;
//   * the tail (and a device-resident DynSize count) is masked in the kernel (reference D4);
//   * `index_base` is added to every index: the shard's global offset on multi-GPU runs.
// Algorithmic bytes: n (mask) + 4 * count (indices); HBM-bound.
#include "hj_internal.h"
#include "lookback.cuh"

namespace hj {
namespace {

constexpr int CMP_THREADS = 256;
constexpr int CMP_WARPS = CMP_THREADS / 32;
constexpr int CMP_ROWS = 8;                         // 512-byte mask rows per warp
constexpr int CMP_ROW = 32 * 16;                    // mask bytes per coalesced warp row
constexpr int CMP_SEG = CMP_ROWS * CMP_ROW;         // per warp: 4 KiB
constexpr int CMP_TILE = CMP_WARPS * CMP_SEG;       // 32768 elements per CTA
constexpr int CMP_CTAS_PER_SM = 8;

// 4 mask bytes -> 4-bit mask of the non-zero ones (bit k = byte k).
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t w) {
    uint32_t y = __vcmpne4(w, 0u) & 0x01010101u;  // bit 8k = (byte k != 0)
    // multiply-gather: 2^21 + 2^14 + 2^7 + 1 moves bit 8k to bit 21+k without collisions
    return (y * 0x00204081u >> 21) & 0xFu;
}
__device__ __forceinline__ uint32_t mask16(const uint4& v) {
    return nonzero_nibble(v.x) | (nonzero_nibble(v.y) << 4) | (nonzero_nibble(v.z) << 8) |
           (nonzero_nibble(v.w) << 12);
}
// guarded variant for the ragged last tile: bytes at and beyond n_eff count as 0
__device__ __forceinline__ uint32_t mask16_guarded(const uint8_t* mask, size_t e0, size_t n_eff) {
    uint32_t bits = 0;
    if (e0 < n_eff) {
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (e0 + j < n_eff && mask[e0 + j] != 0) bits |= 1u << j;
    }
    return bits;
}

__global__ void __launch_bounds__(CMP_THREADS, CMP_CTAS_PER_SM)
compress_kernel(const uint8_t* __restrict__ mask, size_t n, const uint32_t* __restrict__ size_buf,
                uint32_t* __restrict__ out_count, uint32_t* __restrict__ index_out,
                uint32_t index_base, LookbackView lb, int vec_ok) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[CMP_WARPS];
    __shared__ uint32_t s_stage[CMP_WARPS][CMP_ROW];  // 16 KiB: one row of indices per warp

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
    const size_t lane_base = base + (size_t)warp * CMP_SEG + lane * 16;  // this lane's vector, row 0

    // ---- sweep 1: number of selected elements in this warp's segment
    uint32_t cnt = 0;
    if (full) {
        const uint4* vsrc = reinterpret_cast<const uint4*>(mask + lane_base);
        uint4 raw[CMP_ROWS];
#pragma unroll
        for (int r = 0; r < CMP_ROWS; r++) raw[r] = __ldg(vsrc + r * 32);
#pragma unroll
        for (int r = 0; r < CMP_ROWS; r++) cnt += __popc(mask16(raw[r]));
    } else {
        for (int r = 0; r < CMP_ROWS; r++) cnt += __popc(mask16_guarded(mask, lane_base + (size_t)r * CMP_ROW, n_eff));
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) s_warp[warp] = cnt;
    __syncthreads();

    if (warp == 0) {
        uint32_t v = lane < CMP_WARPS ? s_warp[lane] : 0u;
        uint32_t inc = warp_inclusive_sum(v);
        uint32_t aggregate = __shfl_sync(0xffffffffu, inc, 31);
        uint32_t exclusive = 0;
        if (tile == 0) {
            if (lane == 0) tile_publish<uint32_t>(lb, 0, TILE_INCLUSIVE, aggregate);
        } else {
            if (lane == 0) tile_publish<uint32_t>(lb, tile, TILE_AGGREGATE, aggregate);
            exclusive = tile_lookback<uint32_t>(lb, tile);
            if (lane == 0) tile_publish<uint32_t>(lb, tile, TILE_INCLUSIVE, exclusive + aggregate);
        }
        // global rank of the first selected element of each warp's segment
        if (lane < CMP_WARPS) s_warp[lane] = exclusive + inc - v;
        // compress_large.glsl:214-216: the last partition publishes the count
        if (lane == 0 && tile == gridDim.x - 1) out_count[0] = exclusive + aggregate;
    }
    __syncthreads();

    // ---- sweep 2: re-read the rows, compact each through the warp's stage, write coalesced
    uint32_t carry = s_warp[warp];
    uint32_t* stage = s_stage[warp];
    for (int r = 0; r < CMP_ROWS; r++) {
        const size_t e0 = lane_base + (size_t)r * CMP_ROW;
        uint32_t bits = full ? mask16(__ldg(reinterpret_cast<const uint4*>(mask + e0)))
                             : mask16_guarded(mask, e0, n_eff);
        const uint32_t c = __popc(bits);
        const uint32_t inc = warp_inclusive_sum(c);
        const uint32_t row_total = __shfl_sync(0xffffffffu, inc, 31);
        uint32_t k = inc - c;
        const uint32_t first = index_base + (uint32_t)e0;
        while (bits) {
            int b = __ffs(bits) - 1;
            bits &= bits - 1;
            stage[k++] = first + b;
        }
        __syncwarp();
        for (uint32_t q = lane; q < row_total; q += 32) index_out[carry + q] = stage[q];
        __syncwarp();
        carry += row_total;
    }
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
