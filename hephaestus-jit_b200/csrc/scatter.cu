// scatter.cu — scatter-reduce (histogram), gather and fill kernels for sm_100a.
//
// ScatterReduce follows the meaning of the reference's kernel op
// (hephaestus-jit/src/backend/vulkan/codegen/glsl/mod.rs:400-446: atomicOp(buffer[idx], src)),
// Gather that of glsl/mod.rs:523-579.  In the reference both only exist inside JIT-generated
// one-thread-per-element shaders; BASELINE.json names them as hand-written kernels for the
// histogram / Monte-Carlo workloads, so execute_graph routes the matching IR shapes here.
//
// Histogram strategy (u32/i32 sum into n_dst bins):
//   * n_dst * 4 B <= 64 KiB : per-CTA privatised bins in shared memory (shared atomics),
//     flushed once with one global `red` per non-zero bin;
//   * otherwise             : keys are streamed with 128-bit loads and applied with
//     fire-and-forget `red.global` (L2 atomics; 2^16 bins x 4 B = 256 KiB stay L2-resident).
// Algorithmic bytes: 4 per key (+ sizeof(T) per value if a value buffer is given).
#include <algorithm>
#include <type_traits>

#include <cstdlib>

#include "common.cuh"
#include "hj_internal.h"
#include "peer.cuh"
#include "ring.cuh"

namespace hj {
namespace {

enum { R_MAX = HJ_REDUCE_MAX, R_MIN = HJ_REDUCE_MIN, R_SUM = HJ_REDUCE_SUM, R_OR = HJ_REDUCE_OR,
       R_AND = HJ_REDUCE_AND, R_XOR = HJ_REDUCE_XOR };

constexpr int SR_THREADS = 256;

template <typename T, int OP>
__device__ __forceinline__ void atomic_apply(T* addr, T v) {
    if constexpr (OP == R_SUM) atomicAdd(addr, v);
    else if constexpr (OP == R_MAX) atomicMax(addr, v);
    else if constexpr (OP == R_MIN) atomicMin(addr, v);
    else if constexpr (OP == R_OR) atomicOr(addr, v);
    else if constexpr (OP == R_AND) atomicAnd(addr, v);
    else atomicXor(addr, v);
}
// unsigned long long is the CUDA atomic type for 64-bit integers
template <int OP>
__device__ __forceinline__ void atomic_apply_u64(unsigned long long* addr, unsigned long long v) {
    atomic_apply<unsigned long long, OP>(addr, v);
}

// Generic path: global atomics, 4 keys per thread per step through one 128-bit load.
template <typename T, int OP>
__global__ void __launch_bounds__(SR_THREADS)
scatter_reduce_global(const uint32_t* __restrict__ idx, const T* __restrict__ src, T literal,
                      T* __restrict__ dst, size_t n, size_t n_dst, int vec_ok) {
    const size_t stride = (size_t)gridDim.x * SR_THREADS;
    size_t i = (size_t)blockIdx.x * SR_THREADS + threadIdx.x;
    if (vec_ok) {
        const size_t nvec = n / 4;
        const uint4* vidx = reinterpret_cast<const uint4*>(idx);
        for (size_t v = i; v < nvec; v += stride) {
            uint4 k = ld_stream_v4(vidx + v);
            T a = literal, b = literal, c = literal, d = literal;
            if (src) { a = src[4 * v]; b = src[4 * v + 1]; c = src[4 * v + 2]; d = src[4 * v + 3]; }
            if (k.x < n_dst) atomic_apply<T, OP>(dst + k.x, a);
            if (k.y < n_dst) atomic_apply<T, OP>(dst + k.y, b);
            if (k.z < n_dst) atomic_apply<T, OP>(dst + k.z, c);
            if (k.w < n_dst) atomic_apply<T, OP>(dst + k.w, d);
        }
        for (size_t e = nvec * 4 + i; e < n; e += stride) {
            uint32_t k = idx[e];
            if (k < n_dst) atomic_apply<T, OP>(dst + k, src ? src[e] : literal);
        }
    } else {
        for (size_t e = i; e < n; e += stride) {
            uint32_t k = idx[e];
            if (k < n_dst) atomic_apply<T, OP>(dst + k, src ? src[e] : literal);
        }
    }
}

// Privatised path for small bin counts: u32 sum with a literal value.
__global__ void __launch_bounds__(SR_THREADS)
histogram_smem(const uint32_t* __restrict__ idx, uint32_t literal, uint32_t* __restrict__ dst, size_t n,
               uint32_t n_dst, int vec_ok) {
    extern __shared__ uint32_t bins[];
    for (uint32_t b = threadIdx.x; b < n_dst; b += SR_THREADS) bins[b] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * SR_THREADS;
    size_t i = (size_t)blockIdx.x * SR_THREADS + threadIdx.x;
    if (vec_ok) {
        const size_t nvec = n / 4;
        const uint4* vidx = reinterpret_cast<const uint4*>(idx);
        for (size_t v = i; v < nvec; v += stride) {
            uint4 k = ld_stream_v4(vidx + v);
            if (k.x < n_dst) atomicAdd(bins + k.x, literal);
            if (k.y < n_dst) atomicAdd(bins + k.y, literal);
            if (k.z < n_dst) atomicAdd(bins + k.z, literal);
            if (k.w < n_dst) atomicAdd(bins + k.w, literal);
        }
        for (size_t e = nvec * 4 + i; e < n; e += stride) {
            uint32_t k = idx[e];
            if (k < n_dst) atomicAdd(bins + k, literal);
        }
    } else {
        for (size_t e = i; e < n; e += stride) {
            uint32_t k = idx[e];
            if (k < n_dst) atomicAdd(bins + k, literal);
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < n_dst; b += SR_THREADS) {
        uint32_t c = bins[b];
        if (c) atomicAdd(dst + b, c);
    }
}

// ---- privatised histogram with a TMA key ring (u32/i32 sum of a literal) ------------------------
// One persistent CTA per SM; keys stream through a shared-memory ring filled by TMA bulk copies
// (as in ring.cuh, minus the prefix logic: consumers wait for a stage, apply it, hand it back)
// and are applied to bins in shared memory with shared-memory atomics.  How the bins fit:
//   * PACKED16 (literal 1, <= 2^16 bins): two 16-bit counters per 32-bit word, so 2^16 bins are
//     128 KiB and every SM holds ALL bins.  Adds do not use the atomic's return value (ATOMS
//     without a return trip sustains 5.3 lanes/clk/SM, with one 4.0 — profiles/r01_histogram.txt),
//     so nobody sees a counter fill up; instead the consumer warps sweep all counters every
//     HR_SWEEP tiles and move their high bits to the bins in HBM (see HR_SWEEP below).
//   * windows (any literal, <= 8 x 32768 bins): the bin range is split into `parts` windows of
//     <= 32768 u32 bins; groups of `parts` CTAs walk the SAME key tiles, each applying only the
//     keys of its own window (the partners' copies of a tile come out of L2, so HBM still
//     delivers every key once, but every SM ingests `parts` times its share).
// Every CTA (group) ends with a private histogram; they are stored side by side in a scratch
// buffer with plain coalesced stores and a small fold kernel adds them into dst (one global
// atomic per bin and CTA would cost as much as the histogram itself).
// Rejected after measurement (profiles/r01_histogram.txt): a 2-CTA cluster applying the
// partner's keys over DSMEM — remote shared-memory atomics ran at ~0.3 per clock per SM.
constexpr int HR_TILE = 14336, HR_STAGES = 6, HR_WARPS = 28;  // one 512-byte row per warp and tile
constexpr uint32_t HR_MAX_BINS = 32768;

struct HistCtl {
    uint64_t full[HR_STAGES];
    uint64_t empty[HR_STAGES];
};

// PACKED16: every HR_SWEEP tiles the consumer warps sweep the counters between two barriers and move
// the top six bits of each to the bin in HBM.  After a sweep a counter is < 0x400, a CTA applies
// at most HR_SWEEP * HR_TILE / 4 = 64512 keys between two sweeps, 0x3ff + 64512 = 0xffff: no
// counter can wrap or carry whatever the key distribution.
constexpr int HR_SWEEP = 18;
constexpr uint32_t HR_HIGH = 0xfc00fc00u;
static_assert((~HR_HIGH & 0xffffu) + HR_SWEEP * (HR_TILE / 4) <= 0xffff, "a 16-bit counter must survive one sweep period");
// `rs` (u32 bins, one window): log2 of the number of COPIES of every bin.  Copy c of bin b is word
// (b << rs) + c and lane l adds to copy l & (2^rs - 1): with 32 copies (<= 1023 bins) a lane only
// ever touches bank l, so a warp's adds never collide — not in a bank, and not on an address either,
// which is what serialises a skewed histogram (all 32 lanes on one hot bin).  Fewer copies for more
// bins (16 up to 2047, ... 2 up to 16383) still thin the collisions out.
// ADAPTIVE sweeping (packed-16, round 2).  The sweeps cost 9 % and exist for one reason: a 16-bit counter
// must not wrap, which takes more than 65535 keys of ONE CTA's share in ONE bin — never for the uniform
// keys of the BASELINE histogram (27 per bin and CTA at 2^28 keys), routinely for skewed ones.  So the
// first HR_SWEEP tiles (64512 keys: nothing can wrap yet) run without a sweep, then the consumer warps
// look at the fullest counter once: if its count, projected over the CTA's remaining tiles, stays below
// half the range, the rest of the kernel runs WITHOUT sweeps, else with them as before.  The projection
// assumes stationary keys; when it is wrong the result is still exact: every add goes to exactly one
// 16-bit counter (a bin, or a lane's dummy counter for keys out of range), and a wrap — or a carry
// into the neighbouring half — only ever LOSES at least 65535 from the sum of all counters, so "sum of
// the CTA's counters == keys the CTA applied" proves that nothing wrapped; a CTA that cannot prove it
// clears its counters and walks its tiles again with the sweeps on (nothing has reached `dst` before).
// Uniform keys: 5450 -> 6340 GB/s at 2^28 keys; skewed keys take the sweeping path as before.
template <bool PACKED16>
__global__ void __launch_bounds__((HR_WARPS + 1) * 32, 1)
hist_ring_kernel(const uint32_t* __restrict__ keys, size_t n, uint32_t literal, void* __restrict__ out,
                 uint32_t* __restrict__ dst, uint32_t n_dst, uint32_t bins_per_part, uint32_t parts, uint32_t rs,
                 uint32_t adaptive) {
    extern __shared__ __align__(128) char smem[];
    __shared__ uint32_t s_max, s_sum, s_mode;
    char* stages = smem;
    HistCtl* ctl = reinterpret_cast<HistCtl*>(smem + (size_t)HR_STAGES * HR_TILE);
    uint32_t* bins = reinterpret_cast<uint32_t*>(smem + (size_t)HR_STAGES * HR_TILE + sizeof(HistCtl));
    const int warp = warp_id(), lane = lane_id();
    const uint32_t part = blockIdx.x % parts, group = blockIdx.x / parts, n_groups = gridDim.x / parts;
    const uint32_t lo = part * bins_per_part;
    const uint32_t nb = min(bins_per_part, n_dst - lo);
    const uint32_t n_words = PACKED16 ? (bins_per_part + 1) / 2 : (rs ? (bins_per_part + 1) << rs : bins_per_part);
    const size_t n_bytes = n * 4;
    const uint32_t n_tiles = (uint32_t)((n_bytes + HR_TILE - 1) / HR_TILE);
    const uint32_t my_tiles = group < n_tiles ? (n_tiles - group + n_groups - 1) / n_groups : 0u;
    const uint32_t n_zero = (PACKED16 ? ((n_words + 3u) & ~3u) : n_words) + 32u;  // + the 32 dummy words

    // 128-bit stores: the bins start 16-byte aligned behind the ring and its control block
    auto zero_bins = [&]() {
        uint4* z = reinterpret_cast<uint4*>(bins);
        for (uint32_t b = threadIdx.x; b < n_zero / 4; b += blockDim.x) z[b] = make_uint4(0, 0, 0, 0);
        for (uint32_t b = (n_zero & ~3u) + threadIdx.x; b < n_zero; b += blockDim.x) bins[b] = 0;
    };
    zero_bins();
    if (threadIdx.x == 0) {
        for (int s = 0; s < HR_STAGES; s++) {
            mbar_init(&ctl->full[s], 1);
            mbar_init(&ctl->empty[s], HR_WARPS);
        }
        s_max = 0;
        s_sum = 0;
        s_mode = 0;
        // programmatic dependent launch: the fold kernel may be set up while this one runs (it waits
        // for this grid's completion and memory flush before it reads the private histograms)
        pdl_launch_dependents();
    }
    pdl_wait();  // zeroing the bins and the barrier init overlap the tail of the kernel in front (PDL)
    __syncthreads();

    // ring state survives a second walk over the tiles (mbarrier phases simply continue)
    int ps = 0, cs = 0;
    uint32_t puse = 0, cpar = 0;
    // hmode (uniform over the consumer warps): 0 first period, 1 no sweeps, 2 sweeps
    for (int attempt = 0; attempt < 2; attempt++) {
        if (warp == HR_WARPS) {  // the producer takes the highest warp id (issue arbiter favours it)
            for (uint32_t t = group; t < n_tiles; t += n_groups) {
                if (puse > 0) mbar_wait(&ctl->empty[ps], (puse - 1) & 1);
                char* stage = stages + (size_t)ps * HR_TILE;
                const size_t off = (size_t)t * HR_TILE;
                const size_t left = n_bytes - off;
                const uint32_t bytes = left < (size_t)HR_TILE ? (uint32_t)left : (uint32_t)HR_TILE;
                const uint32_t bulk = bytes & ~15u;
                if (bytes < (uint32_t)HR_TILE) {
                    // ragged last tile: pad with a key no window accepts
                    for (uint32_t b = bulk + lane * 4; b < (uint32_t)HR_TILE; b += 128)
                        *reinterpret_cast<uint32_t*>(stage + b) = b < bytes ? keys[(off + b) / 4] : 0xffffffffu;
                    __syncwarp();
                }
                if (lane == 0) {
                    if (bulk) {
                        mbar_expect_tx(&ctl->full[ps], bulk);
                        tma_load_1d(stage, reinterpret_cast<const char*>(keys) + off, bulk, &ctl->full[ps]);
                    } else {
                        mbar_arrive(&ctl->full[ps]);
                    }
                }
                if (++ps == HR_STAGES) { ps = 0; puse++; }
            }
        } else {
            const int cw = warp;
            const uint32_t bins_s = smem_u32(bins);
            uint32_t since_sweep = 0;
            uint32_t hmode = (PACKED16 && adaptive && attempt == 0) ? 0u : 2u;
            for (uint32_t t = group; t < n_tiles; t += n_groups) {
                mbar_wait(&ctl->full[cs], cpar);
                const uint4 k = lds_v4(stages + (size_t)cs * HR_TILE + cw * 512 + lane * 16);
                const uint32_t kk[4] = {k.x, k.y, k.z, k.w};
                if (PACKED16) {
                    // PACKED16 runs with one window (lo == 0).  No branch per key: keys outside the bins
                    // (and the padding of a ragged tile) go to a dummy counter — one word per lane, so
                    // they do not serialise — behind the real ones, which is never flushed or stored.
                    const uint32_t dummy = ((nb + 1u) & ~1u) + 2u * lane;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        // keys in [nb, dummy) are dummy counters too (of lower lanes, or the unused upper
                        // half of the last word when nb is odd): one VIMNMX instead of compare + select
                        const uint32_t a = min(kk[i], dummy);
                        uint32_t addr, val;
                        asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr) : "r"(a & ~1u), "r"(bins_s));  // word address
                        asm("mad.lo.u32 %0, %1, 0xffff, 1;" : "=r"(val) : "r"(a & 1u));           // 1 or 0x10000
                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(val) : "memory");
                    }
                } else {
                    // same trick with u32 bins: keys outside this window (below lo they wrap to huge
                    // values) land on the lane's dummy word behind the window
                    if (rs) {  // replicated bins; row `nb` is the dummy row
                        const uint32_t mine_s = bins_s + 4u * (lane & ((1u << rs) - 1u)), sh = rs + 2u;
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint32_t a = min(kk[i] - lo, nb);
                            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(mine_s + (a << sh)), "r"(literal) : "memory");
                        }
                    } else {
                        const uint32_t dummy = nb + lane;
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint32_t a = min(kk[i] - lo, dummy);
                            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(bins_s + 4u * a), "r"(literal) : "memory");
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->empty[cs]);
                if (++cs == HR_STAGES) { cs = 0; cpar ^= 1; }
                if (PACKED16 && hmode != 1u && ++since_sweep == HR_SWEEP) {
                    // every consumer warp has applied the same HR_SWEEP tiles: decide / sweep between
                    // barriers of the consumer warps (the producer keeps streaming keys meanwhile)
                    since_sweep = 0;
                    asm volatile("bar.sync 1, %0;" ::"n"(HR_WARPS * 32) : "memory");
                    if (hmode == 0u) {
                        // the end of the first period: how full is the fullest counter (dummies included)?
                        uint32_t m = 0;
                        for (uint32_t w = threadIdx.x * 4; w < n_words + 32u; w += HR_WARPS * 32 * 4) {
                            const uint4 v = *reinterpret_cast<const uint4*>(bins + w);
                            m = __vmaxu2(m, __vmaxu2(__vmaxu2(v.x, v.y), __vmaxu2(v.z, v.w)));  // per 16-bit half
                        }
                        m = max(m & 0xffffu, m >> 16);
                        m = __reduce_max_sync(0xffffffffu, m);
                        if (lane == 0) atomicMax(&s_max, m);
                        asm volatile("bar.sync 1, %0;" ::"n"(HR_WARPS * 32) : "memory");
                        const uint32_t periods = (my_tiles + HR_SWEEP - 1) / HR_SWEEP;
                        hmode = (uint64_t)s_max * periods < 0x8000u ? 1u : 2u;
                    }
                    if (hmode == 2u) {
                        for (uint32_t w = threadIdx.x * 4; w < n_words; w += HR_WARPS * 32 * 4) {
                            uint4 v = *reinterpret_cast<uint4*>(bins + w);
                            if ((v.x | v.y | v.z | v.w) & HR_HIGH) {
                                const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                                for (int i = 0; i < 4; i++) {
                                    const uint32_t c = vv[i] & HR_HIGH;
                                    if (c && w + i < n_words) {  // (the dummy words behind the bins are never flushed)
                                        bins[w + i] = vv[i] - c;
                                        const uint32_t b0 = 2 * (w + i);
                                        if (c & 0xffffu) atomicAdd(dst + lo + b0, c & 0xffffu);
                                        if ((c >> 16) && b0 + 1 < nb) atomicAdd(dst + lo + b0 + 1, c >> 16);
                                    }
                                }
                            }
                        }
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(HR_WARPS * 32) : "memory");
                }
            }
            if (threadIdx.x == 0) s_mode = hmode;
        }
        __syncthreads();
        if (!(PACKED16 && adaptive) || attempt == 1 || s_mode != 1u) break;  // swept (exact by construction), or too short to wrap
        // the sweep-free walk: prove that no counter wrapped
        uint32_t sum = 0;
        for (uint32_t w = threadIdx.x; w < n_words + 32u; w += blockDim.x) sum += (bins[w] & 0xffffu) + (bins[w] >> 16);
        sum = __reduce_add_sync(0xffffffffu, sum);
        if (lane == 0) atomicAdd(&s_sum, sum);
        __syncthreads();
        if (s_sum == my_tiles * (uint32_t)(HR_TILE / 4)) break;
        __syncthreads();
        zero_bins();  // a counter wrapped: once more, with the sweeps on
        __syncthreads();
    }
    // private histogram of this group: coalesced plain stores, folded by hist_fold_kernel
    if (PACKED16) {
        const uint32_t n_w = (n_dst + 1) / 2;
        uint32_t* mine = reinterpret_cast<uint32_t*>(out) + (size_t)group * n_w;
        if ((((size_t)group * n_w) & 3u) == 0) {  // 128-bit stores when this group's slice of the scratch is 16-byte aligned
            uint4* m4 = reinterpret_cast<uint4*>(mine);
            const uint4* b4 = reinterpret_cast<const uint4*>(bins);
            for (uint32_t b = threadIdx.x; b < n_w / 4; b += blockDim.x) m4[b] = b4[b];
            for (uint32_t b = (n_w & ~3u) + threadIdx.x; b < n_w; b += blockDim.x) mine[b] = bins[b];
        } else {
            for (uint32_t b = threadIdx.x; b < n_w; b += blockDim.x) mine[b] = bins[b];
        }
    } else {
        uint32_t* mine = reinterpret_cast<uint32_t*>(out) + (size_t)group * n_dst + lo;
        if (rs) {  // sum the copies; thread b starts at copy b so that a warp's reads spread over the banks
            const uint32_t copies = 1u << rs;
            for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
                uint32_t acc = 0;
                for (uint32_t c = 0; c < copies; c++) acc += bins[(b << rs) + ((c + b) & (copies - 1u))];
                mine[b] = acc;
            }
        } else {
            for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) mine[b] = bins[b];
        }
    }
}


// (Round 2 measured and rejected a two-pass radix PARTITION of the keys by 32 Ki-bin window for histograms
// with many bins — rank keys inside (tile, window) groups with warp match + shared atomics, one block of a
// partition buffer per tile, then one window per CTA: 798 / 532 / 325 GB/s at 2^18 / 2^20 / 2^22 bins against
// 1048 GB/s for the window groups below and 775 GB/s for plain L2 atomics at any bin count;
// profiles/r02_histogram.txt.)
// dst[b] += sum over groups of their private counter (u32, or u16 pairs packed in u32 words).
// The groups are split over gridDim.y so that enough loads are in flight to stream the scratch
// (148 x 128 KiB) at HBM speed; each slice adds its partial sum with one global atomic per bin.
constexpr int HF_SLICES = 8;
template <bool PACKED16>
__global__ void __launch_bounds__(256)
hist_fold_kernel(const uint32_t* __restrict__ scratch, uint32_t n_groups, uint32_t* __restrict__ dst, uint32_t n_dst) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    const uint32_t per = (n_groups + gridDim.y - 1) / gridDim.y;
    const uint32_t g0 = blockIdx.y * per, g1 = min(n_groups, g0 + per);
    if (PACKED16) {
        const uint32_t n_words = (n_dst + 1) / 2;
        if (i >= n_words) return;
        uint32_t even = 0, odd = 0;
#pragma unroll 4
        for (uint32_t g = g0; g < g1; g++) {
            const uint32_t w = __ldg(scratch + (size_t)g * n_words + i);
            even += w & 0xffffu;
            odd += w >> 16;
        }
        if (even) atomicAdd(dst + 2 * i, even);
        if (odd && 2 * i + 1 < n_dst) atomicAdd(dst + 2 * i + 1, odd);
    } else {
        if (i >= n_dst) return;
        uint32_t acc = 0;
#pragma unroll 4
        for (uint32_t g = g0; g < g1; g++) acc += __ldg(scratch + (size_t)g * n_dst + i);
        if (acc) atomicAdd(dst + i, acc);
    }
}

// PACKED16 fold AND, on a sharded launch, the cross-GPU all-reduce of the bins in ONE kernel
// (launched with programmatic stream serialization behind hist_ring_kernel).  A CTA owns 64 packed
// words (128 bins) and splits the group list over 4 slices of 64 threads; slice 0 then holds the
// rank's final counts (round 2: 16 slices of 32 words instead of 4 of 64 — the first version was
// latency-bound at 12.7 us for 19 MB, ncu profiles/r02_ncu_summary.txt), adds what dst held, pushes
// each pair of bins as one self-validating uint4
// (bin, epoch, bin, epoch) into every peer's inbox over NVLink, polls its own inbox for the peers'
// pairs and stores the sum over all ranks (peer.cuh: array_pair_allreduce_add) — fold kernel,
// 8 x 65536 global atomics and the separate exchange kernel of round 1 become one launch.
constexpr int HFX_WORDS = 32, HFX_SLICES = 16;  // 512 threads: one 128-byte line per slice warp and group, ~9 loads in flight per thread
__global__ void __launch_bounds__(HFX_WORDS * HFX_SLICES)
hist_fold_exchange_kernel(const uint32_t* __restrict__ scratch, uint32_t n_groups, uint32_t* __restrict__ dst, uint32_t n_dst,
                          ArrayPeerView ax) {
    __shared__ uint32_t s_even[HFX_SLICES][HFX_WORDS], s_odd[HFX_SLICES][HFX_WORDS];
    __shared__ uint32_t s_epoch;
    const uint32_t wl = threadIdx.x % HFX_WORDS, slice = threadIdx.x / HFX_WORDS;
    const uint32_t w = blockIdx.x * HFX_WORDS + wl;
    const uint32_t n_words = (n_dst + 1) / 2;
    pdl_launch_dependents();
    pdl_wait();  // the private histograms of hist_ring_kernel are complete and visible from here on
    uint32_t even = 0, odd = 0;
    if (w < n_words) {
        const uint32_t per = (n_groups + HFX_SLICES - 1) / HFX_SLICES;
        const uint32_t g0 = slice * per, g1 = min(n_groups, g0 + per);
        constexpr int HFX_MAX = 10;  // groups per slice with <= 160 groups: all loads of a thread in flight at once
        if (g1 - g0 <= HFX_MAX) {
            uint32_t word[HFX_MAX];
#pragma unroll
            for (int k = 0; k < HFX_MAX; k++) word[k] = g0 + k < g1 ? __ldg(scratch + (size_t)(g0 + k) * n_words + w) : 0u;
#pragma unroll
            for (int k = 0; k < HFX_MAX; k++) {
                even += word[k] & 0xffffu;
                odd += word[k] >> 16;
            }
        } else {
#pragma unroll 8
            for (uint32_t g = g0; g < g1; g++) {
                const uint32_t word = __ldg(scratch + (size_t)g * n_words + w);
                even += word & 0xffffu;
                odd += word >> 16;
            }
        }
    }
    s_even[slice][wl] = even;
    s_odd[slice][wl] = odd;
    if (threadIdx.x == 0 && ax.world > 1) s_epoch = xepoch_begin(ax.xepoch);
    __syncthreads();
    const uint32_t epoch = s_epoch;
    if (slice == 0 && w < n_words) {
#pragma unroll
        for (int k = 1; k < HFX_SLICES; k++) {
            even += s_even[k][wl];
            odd += s_odd[k][wl];
        }
        const bool has_odd = 2 * w + 1 < n_dst;
        uint32_t a = dst[2 * w] + even, b = has_odd ? dst[2 * w + 1] + odd : 0u;
        if (ax.world > 1) array_pair_allreduce_add(ax, epoch, w, a, b, &a, &b);
        dst[2 * w] = a;
        if (has_odd) dst[2 * w + 1] = b;
    }
    // the CTA that finishes last commits the exchange epoch (every CTA has read it by then)
    if (ax.world > 1) array_exchange_commit(ax, epoch);
}

template <typename T>
__global__ void __launch_bounds__(256)
gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ idx, T* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) dst[i] = __ldg(src + idx[i]);
}

// 4-byte elements, 16-byte aligned idx / dst: indices arrive as 128-bit streaming loads, NV vectors
// (4 * NV independent gathers) per thread are in flight, results leave as 128-bit streaming stores.
// The table goes through the read-only path; what bounds the kernel is the L2 sector rate
// (L2-resident table) or the DRAM random-access rate (one 32-byte sector — 64 bytes with the default
// L2 fetch granularity — fetched per 4 bytes used).  NOALLOC: bypass L1 for the table (a table far
// larger than L1 + L2 never hits there; the lines only evict the index stream's).
// How the table is read (LOADK): 0 ld.global.nc (__ldg), 1 ld.global.nc.L1::no_allocate, 2 plain ld.global,
// 3 ld.global.cg (L2 only), 4 ld.global.cv, 5 ld.relaxed.gpu, 6 ld.global.nc.L2::cache_hint evict_first
template <int LOADK>
__device__ __forceinline__ uint32_t ld_table(const uint32_t* p, uint64_t pol) {
    uint32_t v;
    if constexpr (LOADK == 1) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if constexpr (LOADK == 2) asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if constexpr (LOADK == 3) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if constexpr (LOADK == 4) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if constexpr (LOADK == 5) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if constexpr (LOADK == 6) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    else v = __ldg(p);
    return v;
}
template <int NV, int LOADK>
__global__ void __launch_bounds__(256)
gather4_vec_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t* __restrict__ dst,
                   size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    uint64_t pol = 0;
    if constexpr (LOADK == 6) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const size_t nvec = n / 4;
    const uint4* vidx = reinterpret_cast<const uint4*>(idx);
    uint4* vdst = reinterpret_cast<uint4*>(dst);
    const size_t stride = (size_t)gridDim.x * 256;
    size_t v = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (; v + (NV - 1) * stride < nvec; v += NV * stride) {
        uint4 a[NV], x[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) a[k] = ld_stream_v4(vidx + v + k * stride);
#pragma unroll
        for (int k = 0; k < NV; k++) {
            x[k].x = ld_table<LOADK>(src + a[k].x, pol); x[k].y = ld_table<LOADK>(src + a[k].y, pol);
            x[k].z = ld_table<LOADK>(src + a[k].z, pol); x[k].w = ld_table<LOADK>(src + a[k].w, pol);
        }
#pragma unroll
        for (int k = 0; k < NV; k++) st_stream_v4(vdst + v + k * stride, x[k]);
    }
    for (; v < nvec; v += stride) {
        const uint4 a = ld_stream_v4(vidx + v);
        uint4 x;
        x.x = __ldg(src + a.x); x.y = __ldg(src + a.y); x.z = __ldg(src + a.z); x.w = __ldg(src + a.w);
        st_stream_v4(vdst + v, x);
    }
    for (size_t e = nvec * 4 + (size_t)blockIdx.x * 256 + threadIdx.x; e < n; e += stride) dst[e] = __ldg(src + idx[e]);
}

template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T* __restrict__ dst, size_t n, T v) {
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) dst[i] = v;
}

template <typename T, int OP>
hj_status run_sr(hj_device* dev, size_t n, const uint32_t* idx, const void* src, uint64_t literal,
                 void* dst, size_t n_dst) {
    T lit;
    memcpy(&lit, &literal, sizeof(T));
    int vec_ok = ((uintptr_t)idx & 15u) == 0;
    size_t want = (n / 4 + SR_THREADS - 1) / SR_THREADS;
    size_t cap = (size_t)dev->sm_count * 8;
    int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    scatter_reduce_global<T, OP><<<grid, SR_THREADS, 0, dev->stream>>>(idx, (const T*)src, lit, (T*)dst, n,
                                                                      n_dst, vec_ok);
    return check_launch(dev, "scatter_reduce_global");
}

template <typename T>
hj_status run_sr_int(hj_device* dev, hj_reduce_op op, size_t n, const uint32_t* idx, const void* src,
                     uint64_t literal, void* dst, size_t n_dst) {
    switch (op) {
    case HJ_REDUCE_SUM: return run_sr<T, R_SUM>(dev, n, idx, src, literal, dst, n_dst);
    case HJ_REDUCE_MAX: return run_sr<T, R_MAX>(dev, n, idx, src, literal, dst, n_dst);
    case HJ_REDUCE_MIN: return run_sr<T, R_MIN>(dev, n, idx, src, literal, dst, n_dst);
    case HJ_REDUCE_OR: return run_sr<T, R_OR>(dev, n, idx, src, literal, dst, n_dst);
    case HJ_REDUCE_AND: return run_sr<T, R_AND>(dev, n, idx, src, literal, dst, n_dst);
    case HJ_REDUCE_XOR: return run_sr<T, R_XOR>(dev, n, idx, src, literal, dst, n_dst);
    default: return HJ_ERR_UNSUPPORTED;
    }
}

}  // namespace

hj_status launch_scatter_reduce(hj_device* dev, hj_reduce_op op, hj_type_kind ty, size_t n,
                                const uint32_t* idx, const void* src, uint64_t literal, void* dst,
                                size_t n_dst, const ArrayPeerView* ax, bool* exchanged) {
    if (exchanged) *exchanged = false;
    if (n == 0) return HJ_OK;
    hj_status s = HJ_ERR_UNSUPPORTED;
    if (op == HJ_REDUCE_PROD)  // todo!() in the reference (glsl/mod.rs:422)
        return fail(HJ_ERR_UNSUPPORTED, "scatter_reduce(Prod) is not implemented by the reference");
    // privatised shared-memory histogram (u32/i32 sum of a literal)
    static const int hist_cfg = getenv("HJ_HIST_CFG") ? atoi(getenv("HJ_HIST_CFG")) : 1;  // 0: global atomics
    if (op == HJ_REDUCE_SUM && (ty == HJ_U32 || ty == HJ_I32) && !src && n >= (1u << 16) && hist_cfg != 0) {
        const uint32_t parts = (uint32_t)((n_dst + HR_MAX_BINS - 1) / HR_MAX_BINS);
        if (((uintptr_t)idx & 15u) == 0 && parts <= 8 && n >= (1u << 20)) {
            const size_t ring = (size_t)HR_STAGES * HR_TILE + sizeof(HistCtl);
            const bool packed16 = parts > 1 && n_dst <= 65536 && (uint32_t)literal == 1u && hist_cfg != 2;
            if (packed16) {
                // every SM holds all bins as 16-bit counters and walks only its own key tiles
                const uint32_t n_groups = (uint32_t)dev->sm_count;
                const uint32_t n_words = (uint32_t)((n_dst + 1) / 2);
                HJ_TRY(ensure_hist_scratch(dev, (size_t)n_groups * n_words * 4));
                const size_t smem = ring + ((((size_t)n_words + 3) & ~(size_t)3) + 32) * 4;  // + 32 dummy words
                static const bool old_fold = getenv("HJ_HIST_OLD_FOLD") != nullptr;
                static const bool always_sweep = getenv("HJ_HIST_SWEEP") != nullptr;
                const bool fused_fold = !old_fold && ((uintptr_t)dst & 7u) == 0;
                const bool with_peers = fused_fold && ax && ax->world > 1 && n_words <= ax->slot_vecs;
                ArrayPeerView view = with_peers ? *ax : ArrayPeerView();
                if (!with_peers) { view.world = 1; view.rank = 0; }
                auto kern = hist_ring_kernel<true>;
                HJ_TRY(ensure_dynamic_smem(dev, (const void*)kern, smem));
                HJ_CUDA(launch_pdl(kern, dim3(n_groups), dim3((HR_WARPS + 1) * 32), smem, dev->stream, idx, n, 1u, dev->hist_scratch,
                                   (uint32_t*)dst, (uint32_t)n_dst, (uint32_t)n_dst, 1u, 0u, always_sweep ? 0u : 1u));
                HJ_TRY(check_launch(dev, "hist_ring_kernel"));
                if (fused_fold) {
                    // fold (+ cross-GPU exchange when `ax` is given) in one kernel behind a programmatic
                    // dependency: its launch overlaps the tail of the ring kernel
                    HJ_CUDA(launch_pdl(hist_fold_exchange_kernel, dim3((n_words + HFX_WORDS - 1) / HFX_WORDS),
                                       dim3(HFX_WORDS * HFX_SLICES), 0, dev->stream, (const uint32_t*)dev->hist_scratch, n_groups,
                                       (uint32_t*)dst, (uint32_t)n_dst, view));
                    if (exchanged) *exchanged = with_peers;
                    return check_launch(dev, "hist_fold_exchange_kernel");
                }
                hist_fold_kernel<true><<<dim3((n_words + 255) / 256, HF_SLICES), 256, 0, dev->stream>>>(
                    (const uint32_t*)dev->hist_scratch, n_groups, (uint32_t*)dst, (uint32_t)n_dst);
                return check_launch(dev, "hist_fold_kernel");
            }
            // u32 bins in windows: groups of `parts` CTAs share the key tiles
            const uint32_t bins_per_part = (uint32_t)((n_dst + parts - 1) / parts);
            const uint32_t n_groups = std::max(1u, (uint32_t)dev->sm_count / parts);
            HJ_TRY(ensure_hist_scratch(dev, (size_t)n_groups * n_dst * 4));
            // one window with room to spare: several copies of every bin (see hist_ring_kernel)
            uint32_t rs = 0;
            static const bool no_repl = getenv("HJ_HIST_NO_REPL") != nullptr;
            if (parts == 1 && !no_repl)
                while (rs < 5 && ((size_t)(bins_per_part + 1) << (rs + 1)) <= HR_MAX_BINS + 1024) rs++;  // 1024 bins x 32 copies + the dummy row
            const size_t smem = ring + (rs ? ((size_t)(bins_per_part + 1) << rs) + 32 : (size_t)bins_per_part + 32) * 4;  // + dummy words
            auto kern = hist_ring_kernel<false>;
            HJ_TRY(ensure_dynamic_smem(dev, (const void*)kern, smem));
            HJ_CUDA(launch_pdl(kern, dim3(n_groups * parts), dim3((HR_WARPS + 1) * 32), smem, dev->stream, idx, n, (uint32_t)literal,
                               dev->hist_scratch, (uint32_t*)dst, (uint32_t)n_dst, bins_per_part, parts, rs, 0u));
            HJ_TRY(check_launch(dev, "hist_ring_kernel"));
            hist_fold_kernel<false><<<dim3((unsigned)((n_dst + 255) / 256), HF_SLICES), 256, 0, dev->stream>>>(
                (const uint32_t*)dev->hist_scratch, n_groups, (uint32_t*)dst, (uint32_t)n_dst);
            return check_launch(dev, "hist_fold_kernel");
        }
        if (n_dst * 4 <= 64 * 1024) {
            int vec_ok = ((uintptr_t)idx & 15u) == 0;
            size_t want = (n / 4 + SR_THREADS * 8 - 1) / (SR_THREADS * 8);
            size_t cap = (size_t)dev->sm_count * 2;
            int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
            size_t smem = n_dst * 4;
            if (smem > 48 * 1024)
                HJ_CUDA(cudaFuncSetAttribute(histogram_smem, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
            histogram_smem<<<grid, SR_THREADS, smem, dev->stream>>>(idx, (uint32_t)literal, (uint32_t*)dst, n,
                                                                   (uint32_t)n_dst, vec_ok);
            return check_launch(dev, "histogram_smem");
        }
    }
    switch (ty) {
    case HJ_U32: s = run_sr_int<uint32_t>(dev, op, n, idx, src, literal, dst, n_dst); break;
    case HJ_I32: s = run_sr_int<int32_t>(dev, op, n, idx, src, literal, dst, n_dst); break;
    case HJ_U64: s = run_sr_int<unsigned long long>(dev, op, n, idx, src, literal, dst, n_dst); break;
    case HJ_F32:
        if (op == HJ_REDUCE_SUM) s = run_sr<float, R_SUM>(dev, n, idx, src, literal, dst, n_dst);
        break;
    default: break;
    }
    if (s == HJ_ERR_UNSUPPORTED)
        return fail(HJ_ERR_UNSUPPORTED, "scatter_reduce(%s, %s) unsupported", reduce_op_name(op),
                    type_name(ty));
    return s;
}

hj_status launch_gather(hj_device* dev, size_t elem_bytes, size_t n, const void* src, const uint32_t* idx,
                        void* dst) {
    if (n == 0) return HJ_OK;
    size_t want = (n + 255) / 256, cap = (size_t)dev->sm_count * 16;
    int grid = (int)(want > cap ? cap : want);
    switch (elem_bytes) {
    case 1: gather_kernel<uint8_t><<<grid, 256, 0, dev->stream>>>((const uint8_t*)src, idx, (uint8_t*)dst, n); break;
    case 2: gather_kernel<uint16_t><<<grid, 256, 0, dev->stream>>>((const uint16_t*)src, idx, (uint16_t*)dst, n); break;
    case 4:
        if ((((uintptr_t)idx | (uintptr_t)dst) & 15u) == 0 && n >= 4096) {
            // HJ_GATHER_CFG = 10 * (vectors per thread) + LOADK (see ld_table); default 2 vectors, ld.global.nc
            static const int cfg = getenv("HJ_GATHER_CFG") ? atoi(getenv("HJ_GATHER_CFG")) : 20;
            static const int ctas = getenv("HJ_GATHER_CTAS") ? atoi(getenv("HJ_GATHER_CTAS")) : 16;
            const int nv = cfg / 10;
            const size_t g4 = std::min<size_t>((n / (4 * (size_t)nv) + 255) / 256, (size_t)dev->sm_count * (size_t)ctas);
            const uint32_t *s32 = (const uint32_t*)src;
            uint32_t* d32 = (uint32_t*)dst;
            cudaError_t e;
#define HJ_G(NV, LK) e = launch_pdl(gather4_vec_kernel<NV, LK>, dim3((unsigned)g4), dim3(256), 0, dev->stream, s32, idx, d32, n)
            switch (cfg) {
            case 21: HJ_G(2, 1); break;
            case 22: HJ_G(2, 2); break;
            case 23: HJ_G(2, 3); break;
            case 24: HJ_G(2, 4); break;
            case 25: HJ_G(2, 5); break;
            case 26: HJ_G(2, 6); break;
            case 40: HJ_G(4, 0); break;
            case 42: HJ_G(4, 2); break;
            case 43: HJ_G(4, 3); break;
            case 82: HJ_G(8, 2); break;
            default: HJ_G(2, 0); break;
            }
#undef HJ_G
            HJ_CUDA(e);
        } else
            gather_kernel<uint32_t><<<grid, 256, 0, dev->stream>>>((const uint32_t*)src, idx, (uint32_t*)dst, n);
        break;
    case 8: gather_kernel<uint2><<<grid, 256, 0, dev->stream>>>((const uint2*)src, idx, (uint2*)dst, n); break;
    case 16: gather_kernel<uint4><<<grid, 256, 0, dev->stream>>>((const uint4*)src, idx, (uint4*)dst, n); break;
    default: return fail(HJ_ERR_INVALID, "gather: element size %zu not in {1,2,4,8,16}", elem_bytes);
    }
    return check_launch(dev, "gather_kernel");
}

hj_status launch_fill(hj_device* dev, void* dst, size_t n, size_t elem_bytes, uint64_t pattern) {
    if (n == 0) return HJ_OK;
    size_t want = (n + 255) / 256, cap = (size_t)dev->sm_count * 16;
    int grid = (int)(want > cap ? cap : want);
    switch (elem_bytes) {
    case 1: fill_kernel<uint8_t><<<grid, 256, 0, dev->stream>>>((uint8_t*)dst, n, (uint8_t)pattern); break;
    case 2: fill_kernel<uint16_t><<<grid, 256, 0, dev->stream>>>((uint16_t*)dst, n, (uint16_t)pattern); break;
    case 4: fill_kernel<uint32_t><<<grid, 256, 0, dev->stream>>>((uint32_t*)dst, n, (uint32_t)pattern); break;
    case 8: fill_kernel<unsigned long long><<<grid, 256, 0, dev->stream>>>((unsigned long long*)dst, n, pattern); break;
    default: return fail(HJ_ERR_INVALID, "fill: element size %zu not in {1,2,4,8}", elem_bytes);
    }
    return check_launch(dev, "fill_kernel");
}

}  // namespace hj
