// compress.cu — stream compaction (mask -> ascending indices + count) for sm_100a.
//
// Replaces builtin::compress::compress_large
// (hephaestus-jit/src/backend/vulkan/builtin/compress.rs:157-283 +
// kernels/compress_large.glsl:76-229).  Same contract: index_out[0..count) receives the
// positions of the set mask bytes in ascending order, out_count[0] the count, entries at and
// beyond count are not touched (unless the caller asks for the zero tail, see ZT below).
//
// Two kernels live here:
//   * compress_ring_kernel — THE DEFAULT for 16-byte aligned masks of >= 64 KiB: the persistent
//     ring of ring.cuh with 56 KiB tiles x 3 TMA stages, 28 consumer warps and EARLY release.
//     Phase 1 turns every 16-byte vector into a 16-bit lane mask (IDP.4A: for 0/1 bytes the byte dot
//     product with (1,2,4,8) IS the mask nibble), counts it and parks the mask in a side plane, so
//     the data stage returns to the producer at once; phase 2 picks an emission path per slice by its
//     density: sparse slices store straight from the lanes' bit walks, everything else is staged as u16
//     offsets in output order (lane-major ownership below p ~ 0.5, "quads" above) and leaves as
//     aligned 128-bit stores (CompressOp::SliceOut).  DESIGN.md 3.3 has the measurements.
//     On a sharded launch `finish` also exchanges the per-rank counts over peer memory (comm.cu).
//   * compress_kernel — the fallback for small or misaligned inputs: two sweeps per 32 KiB tile
//     with decoupled look-back (sweep 1 counts, sweep 2 re-reads the L2-resident rows and emits).
// Differences from the reference in both: the tail and a device-resident DynSize count are masked
// in the kernel (D4), `index_base` is added to every index (the shard's global offset), indices
// leave as coalesced warp stores instead of one 32-byte sector per 4-byte index
// (compress_large.glsl:224-228).  Algorithmic bytes: n (mask) + 4 * count (indices).
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "hj_internal.h"
#include "lookback.cuh"
#include "peer.cuh"
#include "ring.cuh"

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
    __shared__ uint32_t s_tile, s_epoch;
    __shared__ uint32_t s_warp[CMP_WARPS];
    __shared__ uint32_t s_stage[CMP_WARPS][CMP_ROW];  // 16 KiB: one row of indices per warp

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const uint32_t e = epoch_begin(lb);
        uint32_t t = atomicAdd(lb.ticket, 1u);
        if (t == gridDim.x - 1) {
            *lb.ticket = 0;
            epoch_advance(lb, e);
        }
        s_tile = t;
        s_epoch = e;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    lb.epoch = s_epoch;
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

// ---- ring pipeline variant (ring.cuh): the default for 16-byte aligned masks ------------------
// Tile = 28 KiB of mask bytes, 28 consumer warps with a 1 KiB slice (two 512-element rows) each.
// Phase 1 turns every 16-byte vector into a 16-bit lane mask (bit tricks, no per-byte work),
// counts it, and parks the mask in the first two bytes of the vector it came from.  Phase 2
// reads the masks back, ranks both rows with ONE packed warp scan, and stages the selected
// positions of a row as u16 offsets in the slice's own bytes (dead by then) before a
// coalesced store of base + offset.  The kernel is instruction-issue bound, not HBM bound, at
// low selectivity, hence the 28 warps and the lean inner loops.
constexpr int CR_WARPS = 28;
constexpr int CR_STAGE_BYTES = 1040;  // per consumer warp: 512 u16 offsets of one row + 3 carried over, 16-byte multiple

// 16 mask bytes -> 16-bit lane mask (bit i = byte i is non-zero).
// Fast path: `bool` buffers only ever hold 0 or 1 (codegen stores bool as u8 0/1,
// codegen/glsl/mod.rs:256-270), so the four-way byte dot product with (1, 2, 4, 8) IS the 4-bit
// mask of a word (IDP.4A, 5 instructions per 16 bytes instead of 13 with multiply-gathers).
// Any other byte value is first normalised to 0/1 (exact "non-zero" test).
__device__ __forceinline__ uint32_t nonzero_lo(uint32_t w) {
    return ((((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u) >> 7;
}
__device__ __forceinline__ uint32_t lane_mask16(uint4 v) {
    if (((v.x | v.y | v.z | v.w) & 0xfefefefeu) != 0u) {
        v.x = nonzero_lo(v.x); v.y = nonzero_lo(v.y); v.z = nonzero_lo(v.z); v.w = nonzero_lo(v.w);
    }
    const uint32_t lo = __dp4a(v.y, 0x80402010u, __dp4a(v.x, 0x08040201u, 0u));
    const uint32_t hi = __dp4a(v.w, 0x80402010u, __dp4a(v.z, 0x08040201u, 0u));
    return lo + (hi << 8);
}

unsigned long long* g_compress_trace = nullptr;  // set through hj_debug_compress_trace()

// ZT ("zero tail"): the kernel also leaves index_out[count .. n) zeroed, so the scheduler's zero-fill
// of the whole index buffer in front of the Compress pass can go (graph_exec.cpp).  Nobody has to
// know `count` for that: the slice that ends at element e, with r selected elements up to e, proves
// that every position >= r + (n - e) lies in the tail (all n - e remaining elements could still be
// selected, no more).  These bounds telescope — slice s owns [B(s), B(s-1)), exactly as many
// positions as it has UNselected elements — so every slice writes 4 bytes per element in total,
// indices below `count` and zeros above, the ranges are disjoint and end at `count`.
template <int SLICE, int TILE, int TSLOTS, bool ZT = false>
struct CompressOp {
    using P = uint32_t;
    static constexpr int ROWS = SLICE / 512;  // 512-element rows per warp slice, ranked in pairs
    static_assert(ROWS % 2 == 0, "rows are ranked in pairs");
    struct Args {
        uint32_t* index_out;
        uint32_t* out_count;
        uint32_t index_base;
        size_t n;  // ZT: number of mask elements = length of index_out
        // sharded compaction (comm.cu): `exchange` makes the CTA that owns the last tile exchange the
        // per-rank counts over peer memory — out_count[0] = global count, counts_out[q] = count of rank q
        uint32_t* counts_out;
        PeerView pv;
        uint32_t exchange;
        // how a slice is emitted, by its number of selected elements: below staged_lo the sparse path (direct
        // stores), [staged_lo, staged_mid) staged with lane-major ownership, from staged_mid on with quads
        // (letting half of the warps of a CTA switch earlier than the other half — lane-major is bound by the
        // shared-memory pipe, the quads by the issue slots — was measured and is slower: a tile is done when its
        // slowest warp is)
        uint32_t staged_lo, staged_mid;
    };
    // aux layout (shared memory behind the ring): TSLOTS mask planes of TILE/8 bytes (one u16 per
    // 16 mask bytes), then one 1 KiB index stage per consumer warp.  Phase 1 leaves only the masks
    // behind, so the data stage goes straight back to the producer (EARLY release).
    static __device__ __forceinline__ uint16_t* stage_of(char* aux, int cw) {
        return reinterpret_cast<uint16_t*>(aux + TSLOTS * MASK_PLANE + cw * CR_STAGE_BYTES);
    }
    static constexpr int MASK_PLANE = TILE / 8;
    static __device__ __forceinline__ uint16_t* masks_of(char* aux, int slot, int cw) {
        return reinterpret_cast<uint16_t*>(aux + slot * MASK_PLANE + cw * (SLICE / 8));
    }
    static __device__ __forceinline__ P total(const char* slice, char* aux, int slot, int cw, int lane) {
        const char* mine = slice + lane * 16;
        uint16_t* m = masks_of(aux, slot, cw);
        uint32_t cnt = 0;
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            const uint32_t b = lane_mask16(lds_v4(mine + r * 512));
            m[r * 32 + lane] = (uint16_t)b;
            cnt += __popc(b);
        }
        return __reduce_add_sync(0xffffffffu, cnt);
    }
    // Staged slices (round 2, variant E of profiles/r02_compress_medium_variants.md).  The selected positions of a
    // slice are staged as u16 offsets relative to the SLICE, in output order, in a per-warp stage that is shifted
    // by the position of the slice's first output inside its 16-byte line: four consecutive entries always are
    // one aligned uint4 of the output.  After every row the complete vectors leave as LDS.64 + STG.128 and the up
    // to three entries behind them move to the front of the stage as the head of the next row; only the first
    // vector of a slice (entries in front of the slice's first output) and its last entries go out as scalar
    // stores.  (Round 1 copied every row out with LDS.U16 + STG.32 and a 64-bit address per index: 94 of the 233
    // warp instructions of a row.)
    struct SliceOut {
        uint16_t* stage;
        uint4* line;     // output line of stage vector 0
        uint32_t base;   // index of the slice's first element
        uint32_t hg;     // stage entries [0, hg) are not ours (only before the first flush)
        uint32_t pos;    // stage entries [hg, pos) are pending
        int lane;
        __device__ __forceinline__ SliceOut(uint16_t* st, uint32_t b, uint32_t* out, int l) : stage(st), base(b), lane(l) {
            hg = (uint32_t)((uintptr_t)out >> 2) & 3u;
            line = reinterpret_cast<uint4*>(out - hg);
            pos = hg;
        }
        // entries [pos, S) have just been staged by the whole warp
        __device__ __forceinline__ void flush(uint32_t S) {
            const uint32_t nv = S >> 2;  // vectors [0, nv) are complete
            __syncwarp();
            if (nv) {
                uint32_t v = lane;
                if (hg) {
                    if ((uint32_t)lane >= hg && lane < 4) reinterpret_cast<uint32_t*>(line)[lane] = base + stage[lane];
                    hg = 0;
                    if (lane == 0) v = 32;
                }
#pragma unroll 1
                for (; v < nv; v += 32) {
                    const uint2 w = *reinterpret_cast<const uint2*>(stage + 4 * v);
                    line[v] = make_uint4(base + (w.x & 0xffffu), base + (w.x >> 16), base + (w.y & 0xffffu), base + (w.y >> 16));
                }
                const uint32_t rem = S & 3u;
                const uint16_t x = (uint32_t)lane < rem ? stage[4 * nv + lane] : (uint16_t)0;
                __syncwarp();
                if ((uint32_t)lane < rem) stage[lane] = x;
                line += nv;
                pos = rem;
            } else {
                pos = S;
            }
            __syncwarp();
        }
        __device__ __forceinline__ void finish() {  // fewer than four entries are left (or no vector was ever complete)
            if ((uint32_t)lane >= hg && (uint32_t)lane < pos) reinterpret_cast<uint32_t*>(line)[lane] = base + stage[lane];
            __syncwarp();
        }
    };
    // "Quad" ownership, for the denser slices: a row is four groups of 128 elements and lane j owns elements
    // 4j .. 4j+3 of EVERY group, so the staging stores of a step ascend by at most 4 entries per lane — 1.2
    // wavefronts per store where lane-major ownership (16 consecutive elements per lane) has 2.9 at p = 0.5.  The
    // four group ranks come from ONE warp scan of the nibble popcounts packed into bytes (inclusive sums <= 128).
    static __device__ __forceinline__ void emit_slice_quads(const uint32_t* words, uint16_t* stage, uint32_t base,
                                                            uint32_t* out, int lane) {
        const uint32_t sh = (lane & 7) * 4, wi = lane >> 3;
        SliceOut so(stage, base, out, lane);
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            uint32_t q[4];
#pragma unroll
            for (int g = 0; g < 4; g++) q[g] = (words[r * 16 + 4 * g + wi] >> sh) & 0xFu;
            uint32_t c = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
            c = c - ((c >> 1) & 0x05050505u);
            c = (c & 0x03030303u) + ((c >> 2) & 0x03030303u);  // byte g = popc(q[g])
            const uint32_t inc = warp_inclusive_sum(c);
            const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
            const uint32_t ex = inc - c;
            uint32_t gb = so.pos;
#pragma unroll
            for (int g = 0; g < 4; g++) {
                uint16_t* p = stage + gb + ((ex >> (8 * g)) & 0xffu);
                const uint32_t val = r * 512 + g * 128 + lane * 4;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (q[g] & (1u << k)) *p++ = (uint16_t)(val + k);
                gb += (tot >> (8 * g)) & 0xffu;
            }
            so.flush(gb);
        }
        so.finish();
    }
    // Lane-major ownership (lane L owns elements 16L .. 16L+15 of a row), for the sparser slices: no nibble
    // extraction, one packed scan per pair of rows; the bank conflicts of its staging stores grow with the
    // density, which is why the quads take over above ~0.4.
    static __device__ __forceinline__ void emit_slice_lanes(const uint16_t* m, uint16_t* stage, uint32_t base, uint32_t* out,
                                                            int lane) {
        SliceOut so(stage, base, out, lane);
#pragma unroll
        for (int r = 0; r < ROWS; r += 2) {
            const uint32_t b0 = m[r * 32 + lane], b1 = m[(r + 1) * 32 + lane];
            const uint32_t c = __popc(b0) | (__popc(b1) << 16);
            const uint32_t inc = warp_inclusive_sum(c);
            const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
            const uint32_t ex = inc - c;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t b = h ? b1 : b0;
                const uint32_t t = h ? tot >> 16 : tot & 0xffffu;
                if (t) {
                    uint16_t* p = stage + so.pos + (h ? ex >> 16 : ex & 0xffffu);
                    const uint32_t val = (r + h) * 512 + lane * 16;
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (b & (1u << j)) *p++ = (uint16_t)(val + j);
                    so.flush(so.pos + t);
                }
            }
        }
        so.finish();
    }
    static __device__ __forceinline__ void emit(const char*, char* aux, int slot, size_t byte_off, uint32_t valid, P carry,
                                                P slice_total, int lane, int cw, const Args& a) {
        if (ZT && valid > slice_total) {  // this slice's share of the zero tail (index_out is 16-byte aligned)
            const uint32_t nz = valid - slice_total;
            const size_t first = (size_t)carry + slice_total + (a.n - (byte_off + valid));
            uint32_t* z = a.index_out + first;
            const uint32_t head = min(nz, (uint32_t)((4u - (first & 3u)) & 3u));
            if ((uint32_t)lane < head) z[lane] = 0;
            const uint32_t n_vec = (nz - head) / 4;
            uint4* zv = reinterpret_cast<uint4*>(z + head);
            for (uint32_t v = lane; v < n_vec; v += 32) st_stream_v4(zv + v, make_uint4(0, 0, 0, 0));
            const uint32_t done = head + 4 * n_vec;
            if ((uint32_t)lane < nz - done) z[done + lane] = 0;
        }
        if (slice_total == 0) return;
        const uint16_t* m = masks_of(aux, slot, cw);
        if (slice_total >= a.staged_lo) {
            // measured over p = 0.01 .. 0.99 (profiles/r02_compress_medium_variants.md): lane-major staging wins
            // below p ~ 0.5, the quads above — all the way to p = 0.99, where they beat the word-broadcast
            // path of round 1 (direct STG.32 runs, no stage) by 9 %
            if (slice_total >= a.staged_mid)
                emit_slice_quads(reinterpret_cast<const uint32_t*>(m), stage_of(aux, cw), a.index_base + (uint32_t)byte_off,
                                 a.index_out + carry, lane);
            else
                emit_slice_lanes(m, stage_of(aux, cw), a.index_base + (uint32_t)byte_off, a.index_out + carry, lane);
            return;
        }
        // Sparse slice (p <~ 0.09; the total is the one phase 1 counted).  The mask plane of the slice IS its
        // bitmap in element order (ROWS * 16 words), so a lane may just as well own WPL consecutive words: ONE
        // warp scan ranks the whole slice and every lane walks its own bits straight to HBM — no per-row scans,
        // no staging.  At p = 0.01 this is 31 M instead of 49 M warp instructions for 2^28 elements; from p ~ 0.09
        // on the scattered 4-byte stores (every lane its own run) cost more than staging does.
        constexpr int WPL = ROWS / 2;
        uint32_t w[WPL];
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            w[i] = reinterpret_cast<const uint32_t*>(m)[lane * WPL + i];
            c += __popc(w[i]);
        }
        uint32_t* out = a.index_out + carry + (warp_inclusive_sum(c) - c);
        const uint32_t e = a.index_base + (uint32_t)byte_off + lane * (32 * WPL);
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            uint32_t b = w[i];
            while (b) {
                const int j = __ffs(b) - 1;
                b &= b - 1;
                *out++ = e + 32 * i + j;
            }
        }
    }
    // compress_large.glsl:214-216: the last partition publishes the count
    static __device__ __forceinline__ void finish(P total, const Args& a, int lane) {
        if (a.exchange == 0) {
            if (lane == 0) a.out_count[0] = total;
            return;
        }
        const uint32_t c = (uint32_t)peer_allgather_warp(a.pv, total, lane);  // lane q: count of rank q
        if (a.counts_out && lane < a.pv.world) a.counts_out[lane] = c;
        const uint32_t all = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) a.out_count[0] = all;
    }
};

template <int CR_TILE, int CR_STAGES, int CR_TSLOTS, int CR_AHEAD, bool TRACE = false, bool ZT = false>
__global__ void __launch_bounds__((CR_WARPS + 3) * 32, 1)
compress_ring_kernel(const uint8_t* __restrict__ mask, size_t n, const uint32_t* __restrict__ size_buf,
                     uint32_t* __restrict__ out_count, uint32_t* __restrict__ index_out, uint32_t index_base,
                     LookbackView lb, uint32_t G, unsigned long long* trace, uint32_t* __restrict__ counts_out,
                     PeerView pv, uint32_t exchange, uint32_t staged_lo, uint32_t staged_mid) {
    extern __shared__ __align__(128) char smem[];
    size_t n_eff = n;
    if (size_buf) {  // DynSize: device-resident element count (graph.rs:503-508)
        pdl_wait();  // written by a kernel in front
        const size_t dyn = size_buf[0];
        n_eff = dyn < n ? dyn : n;
    }
    const uint32_t n_tiles = (uint32_t)((n_eff + CR_TILE - 1) / CR_TILE);
    if (n_tiles == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        pdl_wait();
        out_count[0] = 0;
    }
    using Op = CompressOp<CR_TILE / CR_WARPS, CR_TILE, CR_TSLOTS, ZT>;
    typename Op::Args args{index_out, out_count, index_base, n_eff, counts_out, pv, exchange, staged_lo, staged_mid};
    ring_pipeline<Op, CR_TILE, CR_STAGES, CR_WARPS, CR_AHEAD, CR_TSLOTS, true, 8, TRACE>(
        reinterpret_cast<const char*>(mask), n_eff, n_tiles, 0u, lb, G, args, smem, trace);
}

// index_out[count .. n) = 0, `count` read from the device: lets execute_graph drop the zero-fill
// of the whole index buffer that the scheduler places in front of every Compress
// (trace.rs:1600-1601) — the part below `count` is overwritten by the compaction anyway.
__global__ void __launch_bounds__(256)
compress_zero_tail_kernel(uint32_t* __restrict__ index_out, const uint32_t* __restrict__ count, size_t n) {
    const size_t c = min((size_t)__ldg(count), n);
    const size_t first_vec = (c + 3) / 4, n_vec = n / 4;  // vectors that lie entirely in the tail
    // One contiguous 32 KiB chunk per CTA, eight plain 128-bit stores per thread, no loop over
    // chunks: measured at 2^28 (profiles/r01c_traced_compress.txt), CTAs that stride over chunks
    // write at 5.0 TB/s, one chunk per CTA at 6.5 TB/s.  The grid is sized for count = 0; CTAs beyond
    // the tail leave at once (0.5 ns each).
    uint4* v = reinterpret_cast<uint4*>(index_out);
    const size_t chunk = first_vec + (size_t)blockIdx.x * 2048;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const size_t i = chunk + (size_t)k * 256 + threadIdx.x;
        if (i < n_vec) v[i] = make_uint4(0, 0, 0, 0);
    }
    if (blockIdx.x == 0 && threadIdx.x < 4) {  // the ragged ends: [c, 4 * first_vec) and [4 * n_vec, n)
        const size_t head = c + threadIdx.x;
        if (head < min(first_vec * 4, n)) index_out[head] = 0;
        const size_t tail = max(n_vec * 4, first_vec * 4) + threadIdx.x;
        if (tail < n) index_out[tail] = 0;
    }
}

}  // namespace

hj_status launch_compress_zero_tail(hj_device* dev, uint32_t* index_out, const uint32_t* count, size_t n) {
    if (n == 0) return HJ_OK;
    if (((uintptr_t)index_out & 15u) != 0) {  // foreign, misaligned buffer: plain fill of everything beyond count is not possible without the vector path
        return fail(HJ_ERR_INVALID, "compress: index buffer is not 16-byte aligned");
    }
    const size_t grid = (n / 4 + 2047) / 2048 + 1;
    compress_zero_tail_kernel<<<(unsigned)grid, 256, 0, dev->stream>>>(index_out, count, n);
    return check_launch(dev, "compress_zero_tail_kernel");
}

bool compress_can_fuse_exchange(size_t n, const uint8_t* mask) {
    static const int cfg = getenv("HJ_COMPRESS_CFG") ? atoi(getenv("HJ_COMPRESS_CFG")) : 1;
    return cfg != 0 && ((uintptr_t)mask & 15u) == 0 && n >= (64u << 10) && !g_compress_trace;
}

hj_status launch_compress(hj_device* dev, size_t n, const uint32_t* size_buf, uint32_t* out_count,
                          const uint8_t* mask, uint32_t* index_out, uint32_t index_base, bool zero_tail,
                          uint32_t* counts_out, const PeerView* peers) {
    HJ_REQUIRE(n <= 0xffffffffull, "compress: n does not fit the u32 index type");
    HJ_REQUIRE(!peers || (compress_can_fuse_exchange(n, mask) && !size_buf),
               "compress: the fused exchange needs the ring kernel and a static size");
    if (zero_tail) {  // only the ring kernel on a statically sized, aligned buffer zeroes the tail itself
        static const int c = getenv("HJ_COMPRESS_CFG") ? atoi(getenv("HJ_COMPRESS_CFG")) : 1;
        const bool in_kernel = c == 1 && !size_buf && !g_compress_trace && ((uintptr_t)mask & 15u) == 0 &&
                               ((uintptr_t)index_out & 15u) == 0 && n >= (64u << 10) && !getenv("HJ_ZERO_TAIL_KERNEL");
        if (!in_kernel) {
            HJ_REQUIRE(!peers, "compress: zero tail outside the ring kernel cannot be combined with the fused exchange");
            HJ_TRY(launch_compress(dev, n, size_buf, out_count, mask, index_out, index_base, false));
            return launch_compress_zero_tail(dev, index_out, out_count, n);
        }
    }
    static const int cfg = getenv("HJ_COMPRESS_CFG") ? atoi(getenv("HJ_COMPRESS_CFG")) : 1;  // 0: look-back kernel, 3: 28 KiB tiles
    if (cfg != 0 && ((uintptr_t)mask & 15u) == 0 && n >= (64u << 10)) {
        // tile geometry: 28 KiB tiles, 5 data stages (all but one in flight), 8 tile slots with
        // phase 1 six tiles ahead of phase 2; or 56 KiB tiles x 3 stages, 4 slots, three ahead
        const bool big = cfg != 3;  // HJ_COMPRESS_CFG=3 selects the 28 KiB geometry
        const size_t tile_bytes = big ? 57344 : 28672;
        const size_t tiles = (n + tile_bytes - 1) / tile_bytes;
        HJ_TRY(ensure_lookback_scratch(dev, tiles));
        HJ_TRY(count_epoch(dev));
        LookbackView view = lookback_view(dev->lookback.base, dev->lookback.capacity_tiles, 0);
        const unsigned grid = (unsigned)(tiles < (size_t)dev->sm_count ? tiles : (size_t)dev->sm_count);
        // emission path by the number of selected elements of a 2048-element slice (halved for the 1024-element
        // slices of the 28 KiB geometry): sparse | lane-major staged | quads staged.  HJ_COMPRESS_STAGED=lo,mid moves them.
        static const std::array<uint32_t, 2> staged = [] {
            std::array<uint32_t, 2> t = {176u, 960u};
            if (const char* e = getenv("HJ_COMPRESS_STAGED")) {
                unsigned a = 0, b = 0;
                if (sscanf(e, "%u,%u", &a, &b) == 2) t = {a, b};
            }
            return t;
        }();
        auto launch = [&](auto kernel, size_t smem) -> hj_status {
            HJ_TRY(ensure_dynamic_smem(dev, (const void*)kernel, smem));
            HJ_CUDA(launch_pdl(kernel, dim3(grid), dim3((CR_WARPS + 3) * 32), smem, dev->stream, mask, n, size_buf, out_count,
                               index_out, index_base, view, (grid + 31u) & ~31u, g_compress_trace, counts_out,
                               peers ? *peers : PeerView(), peers ? 1u : 0u, staged[0] / (big ? 1u : 2u), staged[1] / (big ? 1u : 2u)));
            return HJ_OK;
        };
        const size_t smem_small = ring_smem_bytes<uint32_t, 28672, 5, CR_WARPS, 8>(8 * (28672 / 8) + CR_WARPS * CR_STAGE_BYTES);
        const size_t smem_big = ring_smem_bytes<uint32_t, 57344, 3, CR_WARPS, 4>(4 * (57344 / 8) + CR_WARPS * CR_STAGE_BYTES);
        static_assert(ring_smem_bytes<uint32_t, 57344, 3, CR_WARPS, 4>(4 * (57344 / 8) + CR_WARPS * CR_STAGE_BYTES) <= 232448,
                      "compress ring: over the 227 KiB of shared memory a CTA may opt in to");
        // measured on B200 (profiles/r01_ring_sweeps.txt): the 56 KiB geometry wins at low
        // selectivity (fewer status-word sweeps per byte) and ties elsewhere
        if (g_compress_trace) HJ_TRY(launch(compress_ring_kernel<57344, 3, 4, 3, true>, smem_big));  // tools/ring_timeline.py
        else if (big && zero_tail) HJ_TRY(launch(compress_ring_kernel<57344, 3, 4, 3, false, true>, smem_big));
        else if (big) HJ_TRY(launch(compress_ring_kernel<57344, 3, 4, 3>, smem_big));
        else HJ_TRY(launch(compress_ring_kernel<28672, 5, 8, 6>, smem_small));
        return check_launch(dev, "compress_ring_kernel");
    }
    size_t n_tiles = (n + CMP_TILE - 1) / CMP_TILE;
    HJ_TRY(ensure_lookback_scratch(dev, n_tiles));
    HJ_TRY(count_epoch(dev));
    LookbackView lb = lookback_view(dev->lookback.base, dev->lookback.capacity_tiles, 0);
    int vec_ok = ((uintptr_t)mask & 15u) == 0;
    compress_kernel<<<(unsigned)n_tiles, CMP_THREADS, 0, dev->stream>>>(mask, n, size_buf, out_count,
                                                                       index_out, index_base, lb, vec_ok);
    return check_launch(dev, "compress_kernel");
}

}  // namespace hj

extern "C" void hj_debug_compress_trace(void* device_ptr) { hj::g_compress_trace = (unsigned long long*)device_ptr; }
