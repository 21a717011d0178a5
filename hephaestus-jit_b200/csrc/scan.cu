// scan.cu — single-launch inclusive/exclusive prefix sum for sm_100a.
//
// Replaces builtin::prefix_sum::prefix_sum_large
// (hephaestus-jit/src/backend/vulkan/builtin/prefix_sum.rs:31-162 +
// kernels/prefix_sum_large.glsl + prefix_sum_large_init.glsl).
//
// Two kernels live here:
//   * scan_ring_kernel — THE DEFAULT for 16-byte aligned buffers of >= 64 KiB: the persistent
//     warp-specialised ring of ring.cuh (one CTA per SM, tiles by atomic ticket, a TMA bulk copy per
//     32 KiB tile into a 6-stage shared-memory ring, phase 1 = reduce tile i+3, phase 2 = scan tile i
//     from shared memory, round-based prefixes — nobody spins on the critical path).  DRAM traffic
//     is the algorithmic 2 * sizeof(T) bytes per element; 6.9 TB/s at 2^30 u32.
//     On a sharded launch its `finish` also runs the cross-GPU exchange of the shard totals over
//     peer memory and leaves this rank's exclusive offset in `seed_out` (the DEFERRED seed, comm.cu).
//   * scan_kernel — the fallback for small or misaligned inputs: decoupled look-back between 64 KiB
//     super-tiles, two sweeps per super-tile through L2 (sweep 1 sums each warp's 8 KiB segment,
//     look-back resolves the tile prefix, sweep 2 re-reads the rows from L2, scans and stores).
//     The textbook register-resident single sweep (the reference's design, 2048-item partitions)
//     stalls at 48 % of the HBM roofline on B200 and this two-sweep variant at 70 %
//     (profiles/r01_ncu_full_summary.txt) — which is why the ring replaced it as the default.
// Differences from the reference in both: no separate init dispatch (epoch-tagged status words),
// tail masked in the kernel (D7), true exclusive and inclusive variants (D10), correct carries for
// f32/u64/f64 (D3), and `seed`: a device-resident offset added to every output at no extra pass.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "hj_internal.h"
#include "lookback.cuh"
#include "peer.cuh"
#include "ring.cuh"

namespace hj {
unsigned long long* g_scan_trace = nullptr;  // set through hj_debug_scan_trace()
namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_ROWS = 16;   // 512-byte rows per warp  -> 8 KiB per warp, 64 KiB per CTA
constexpr int SCAN_BATCH = 4;   // rows loaded together in sweep 2
// A CTA waiting for its prefix holds no data, only thread slots: run at full occupancy
// (8 x 256 threads, <= 32 registers) so that enough CTAs are always streaming.
constexpr int SCAN_CTAS_PER_SM = 8;

// L2 eviction policies (createpolicy + .L2::cache_hint; the bare .L2::evict_* qualifiers are
// only accepted on 256-bit accesses on sm_100a).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ld_hint_v4(const void* p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_hint_v4(void* p, uint4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
}

template <typename T, typename P, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS, SCAN_CTAS_PER_SM)
scan_kernel(const T* __restrict__ src, T* __restrict__ dst, size_t n, const T* __restrict__ seed,
            LookbackView lb, int vec_ok, unsigned long long* __restrict__ trace) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int ROW = 32 * VEC;              // elements per coalesced warp row
    constexpr int SEG = SCAN_ROWS * ROW;       // elements per warp
    constexpr int TILE = SCAN_WARPS * SEG;     // elements per CTA
    __shared__ uint32_t s_tile, s_epoch;
    __shared__ P s_warp[SCAN_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long t_start = 0, t_summed = 0, t_prefix = 0;
    if (trace && tid == 0) t_start = globaltimer_ns();
    if (tid == 0) {
        const uint32_t e = epoch_begin(lb);
        uint32_t t = atomicAdd(lb.ticket, 1u);
        if (t == gridDim.x - 1) {  // last ticket: re-arm for the next launch
            *lb.ticket = 0;
            epoch_advance(lb, e);
        }
        s_tile = t;
        s_epoch = e;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    lb.epoch = s_epoch;
    const size_t base = (size_t)tile * TILE;
    const bool full = vec_ok && base + TILE <= n;
    const size_t lane_base = base + (size_t)warp * SEG + lane * VEC;  // this lane's vector in row 0

    const uint64_t keep = l2_policy_evict_last();    // sweep-1 lines are re-read in sweep 2
    const uint64_t drop = l2_policy_evict_first();   // sweep-2 reads and the output are final

    // ---- sweep 1: total of this warp's segment
    P total = (P)0;
    if (full) {
        const uint4* vsrc = reinterpret_cast<const uint4*>(src + lane_base);
        uint4 raw[SCAN_ROWS];
#pragma unroll
        for (int r = 0; r < SCAN_ROWS; r++) raw[r] = ld_hint_v4(vsrc + r * 32, keep);
#pragma unroll
        for (int r = 0; r < SCAN_ROWS; r++) {
            const T* e = reinterpret_cast<const T*>(&raw[r]);
#pragma unroll
            for (int j = 0; j < VEC; j++) total = (P)(total + (P)e[j]);
        }
    } else {
        for (int r = 0; r < SCAN_ROWS; r++)
            for (int j = 0; j < VEC; j++) {
                size_t e = lane_base + (size_t)r * ROW + j;
                if (e < n) total = (P)(total + (P)src[e]);
            }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) total = (P)(total + shfl_xor(total, m));
    if (lane == 0) s_warp[warp] = total;
    __syncthreads();
    if (trace && tid == 0) t_summed = globaltimer_ns();

    // ---- warp 0: scan the warp totals, resolve the tile prefix by look-back
    if (warp == 0) {
        P v = lane < SCAN_WARPS ? s_warp[lane] : (P)0;
        P inc = warp_inclusive_sum(v);
        P aggregate = shfl_idx(inc, 31);
        P up = shfl_up(inc, 1);
        P exclusive;
        if (tile == 0) {
            exclusive = seed ? (P)seed[0] : (P)0;
            if (lane == 0) tile_publish<P>(lb, 0, TILE_INCLUSIVE, (P)(exclusive + aggregate));
        } else {
            if (lane == 0) tile_publish<P>(lb, tile, TILE_AGGREGATE, aggregate);
            exclusive = tile_lookback<P>(lb, tile);
            if (lane == 0) tile_publish<P>(lb, tile, TILE_INCLUSIVE, (P)(exclusive + aggregate));
        }
        // offset of each warp's segment = tile prefix + totals of the warps before it
        if (lane < SCAN_WARPS) s_warp[lane] = (P)(exclusive + (lane == 0 ? (P)0 : up));
    }
    __syncthreads();
    if (trace && tid == 0) t_prefix = globaltimer_ns();

    // ---- sweep 2: re-read (L2), scan rows with a running carry, store
    P carry = s_warp[warp];
    if (full) {
        const uint4* vsrc = reinterpret_cast<const uint4*>(src + lane_base);
        uint4* vdst = reinterpret_cast<uint4*>(dst + lane_base);
#pragma unroll 1
        for (int r0 = 0; r0 < SCAN_ROWS; r0 += SCAN_BATCH) {
            uint4 raw[SCAN_BATCH];
#pragma unroll
            for (int b = 0; b < SCAN_BATCH; b++) raw[b] = ld_hint_v4(vsrc + (r0 + b) * 32, drop);
#pragma unroll
            for (int b = 0; b < SCAN_BATCH; b++) {
                const T* e = reinterpret_cast<const T*>(&raw[b]);
                P x[VEC];
                P s = (P)0;
#pragma unroll
                for (int j = 0; j < VEC; j++) { x[j] = (P)e[j]; s = (P)(s + x[j]); }
                P inc = warp_inclusive_sum(s);
                P up = shfl_up(inc, 1);
                P run = (P)(carry + (lane == 0 ? (P)0 : up));
                carry = (P)(carry + shfl_idx(inc, 31));
                T out[VEC];
#pragma unroll
                for (int j = 0; j < VEC; j++) {
                    if (INCLUSIVE) { run = (P)(run + x[j]); out[j] = (T)run; }
                    else { out[j] = (T)run; run = (P)(run + x[j]); }
                }
                st_hint_v4(vdst + (r0 + b) * 32, *reinterpret_cast<const uint4*>(out), drop);
            }
        }
    } else {
        for (int r = 0; r < SCAN_ROWS; r++) {
            P x[VEC];
            P s = (P)0;
            for (int j = 0; j < VEC; j++) {
                size_t e = lane_base + (size_t)r * ROW + j;
                x[j] = e < n ? (P)src[e] : (P)0;
                s = (P)(s + x[j]);
            }
            P inc = warp_inclusive_sum(s);
            P up = shfl_up(inc, 1);
            P run = (P)(carry + (lane == 0 ? (P)0 : up));
            carry = (P)(carry + shfl_idx(inc, 31));
            for (int j = 0; j < VEC; j++) {
                size_t e = lane_base + (size_t)r * ROW + j;
                T o;
                if (INCLUSIVE) { run = (P)(run + x[j]); o = (T)run; }
                else { o = (T)run; run = (P)(run + x[j]); }
                if (e < n) dst[e] = o;
            }
        }
    }
    if (trace && tid == 0) {  // development aid: per-tile timeline (tools/scan_timeline.py)
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        trace[tile * 5 + 0] = t_start;
        trace[tile * 5 + 1] = t_summed;
        trace[tile * 5 + 2] = t_prefix;
        trace[tile * 5 + 3] = globaltimer_ns();
        trace[tile * 5 + 4] = smid;
    }
}

// ---- ring pipeline variant (ring.cuh): the default for 16-byte aligned buffers -----------------
template <typename T, typename P_, bool INCLUSIVE, int SLICE>
struct ScanOp {
    using P = P_;
    static constexpr int VEC = 16 / sizeof(T);
    static constexpr int ROWS = SLICE / 512;
    struct Args {
        T* dst;
        // sharded scan with a DEFERRED seed (comm.cu): `exchange` makes the CTA that owns the last
        // tile publish the shard total to every peer and write this rank's exclusive offset
        T* seed_out;
        PeerView pv;
        uint32_t exchange;
    };
    // phase 1: total of this warp's slice (the stage tail of a ragged tile is zero-filled)
    static __device__ __forceinline__ P total(const char* slice, char*, int, int, int lane) {
        P acc = (P)0;
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            const uint4 raw = lds_v4(slice + r * 512 + lane * 16);
            const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
            for (int j = 0; j < VEC; j++) acc = (P)(acc + (P)e[j]);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc = (P)(acc + shfl_xor(acc, m));
        return acc;
    }
    // phase 2: scan the slice row by row with a running carry and stream the result out
    static __device__ __forceinline__ void emit(const char* slice, char*, int, size_t byte_off, uint32_t valid, P carry,
                                                P, int lane, int, const Args& a) {
        char* out_base = reinterpret_cast<char*>(a.dst) + byte_off;
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            const uint32_t o = r * 512 + lane * 16;
            if ((uint32_t)(r * 512) >= valid) break;  // warp-uniform: nothing left in this slice
            const uint4 raw = lds_v4(slice + o);
            const T* e = reinterpret_cast<const T*>(&raw);
            P x[VEC];
            P s = (P)0;
#pragma unroll
            for (int j = 0; j < VEC; j++) { x[j] = (P)e[j]; s = (P)(s + x[j]); }
            const P inc = warp_inclusive_sum(s);
            P run = (P)(carry + (P)(inc - s));
            carry = (P)(carry + shfl_idx(inc, 31));
            T out[VEC];
#pragma unroll
            for (int j = 0; j < VEC; j++) {
                if (INCLUSIVE) { run = (P)(run + x[j]); out[j] = (T)run; }
                else { out[j] = (T)run; run = (P)(run + x[j]); }
            }
            if (o + 16 <= valid) {
                st_stream_v4(out_base + o, *reinterpret_cast<const uint4*>(out));
            } else {
#pragma unroll
                for (int j = 0; j < VEC; j++)
                    if (o + j * sizeof(T) < valid) reinterpret_cast<T*>(out_base + o)[j] = out[j];
            }
        }
    }
    // Sharded scan, one kernel: the shard total (the last tile's inclusive prefix) goes to every
    // peer's mailbox over NVLink, the totals of the ranks before this one are summed in rank order
    // and left in seed_out[0] — the offset every consumer of `dst` adds.
    static __device__ __forceinline__ void finish(P total, const Args& a, int lane) {
        if (a.exchange == 0) return;
        unsigned long long bits = 0;
        memcpy(&bits, &total, sizeof(P));
        const unsigned long long got = peer_allgather_warp(a.pv, bits, lane);
        P mine;
        memcpy(&mine, &got, sizeof(P));
        P before = (P)0;
        for (int q = 0; q < a.pv.rank; q++) before = (P)(before + shfl_idx(mine, q));
        if (lane == 0) a.seed_out[0] = (T)before;
    }
};

template <typename T, typename P, bool INCLUSIVE, int TILE, int STAGES, int CWARPS, int AHEAD>
__global__ void __launch_bounds__((CWARPS + 3) * 32, 1)
scan_ring_kernel(const T* __restrict__ src, T* __restrict__ dst, size_t n, const T* __restrict__ seed,
                 LookbackView lb, uint32_t n_tiles, uint32_t G, T* __restrict__ seed_out, PeerView pv, uint32_t exchange) {
    extern __shared__ __align__(128) char smem[];
    using Op = ScanOp<T, P, INCLUSIVE, TILE / CWARPS>;
    typename Op::Args args{dst, seed_out, pv, exchange};
    P seed_v = (P)0;
    if (seed) {
        pdl_wait();  // the seed is the output of the kernel in front (sharded scan: the totals pass)
        seed_v = (P)seed[0];
    }
    ring_pipeline<Op, TILE, STAGES, CWARPS, AHEAD, STAGES, false, 10>(reinterpret_cast<const char*>(src), n * sizeof(T), n_tiles,
                                                   seed_v, lb, G, args, smem);
}

template <typename T, typename P, int TILE, int STAGES, int CWARPS, int AHEAD>
hj_status run_ring(hj_device* dev, size_t n, bool inclusive, const void* src, void* dst, const void* seed,
                   void* seed_out, const PeerView* peers) {
    const size_t n_tiles = (n * sizeof(T) + TILE - 1) / TILE;
    HJ_REQUIRE(n_tiles < (1ull << 31), "prefix_sum: too many tiles");
    HJ_TRY(ensure_lookback_scratch(dev, n_tiles));
    HJ_TRY(count_epoch(dev));
    LookbackView lb = lookback_view(dev->lookback.base, dev->lookback.capacity_tiles, 0);
    const size_t smem = ring_smem_bytes<P, TILE, STAGES, CWARPS>();
    const unsigned grid = (unsigned)std::min<size_t>(n_tiles, (size_t)dev->sm_count);
    const uint32_t G = (grid + 31u) & ~31u;  // tiles per round: about one per CTA
    auto launch = [&](auto kernel) -> hj_status {
        HJ_TRY(ensure_dynamic_smem(dev, (const void*)kernel, smem));
        HJ_CUDA(launch_pdl(kernel, dim3(grid), dim3((CWARPS + 3) * 32), smem, dev->stream, (const T*)src, (T*)dst, n,
                           (const T*)seed, lb, (uint32_t)n_tiles, G, (T*)seed_out, peers ? *peers : PeerView(),
                           peers ? 1u : 0u));
        return check_launch(dev, "scan_ring_kernel");
    };
    return inclusive ? launch(scan_ring_kernel<T, P, true, TILE, STAGES, CWARPS, AHEAD>)
                     : launch(scan_ring_kernel<T, P, false, TILE, STAGES, CWARPS, AHEAD>);
}

template <typename T, typename P>
hj_status run(hj_device* dev, size_t n, bool inclusive, const void* src, void* dst, const void* seed,
              void* seed_out = nullptr, const PeerView* peers = nullptr) {
    // 16-byte aligned buffers (every buffer this library allocates) take the ring pipeline;
    // anything else, and tiny inputs, the look-back kernel below.
    const bool aligned = (((uintptr_t)src | (uintptr_t)dst) & 15u) == 0;
    static const int cfg = getenv("HJ_SCAN_CFG") ? atoi(getenv("HJ_SCAN_CFG")) : 1;  // 0: look-back kernel
    if (aligned && cfg != 0 && n * sizeof(T) >= (64u << 10)) {
        // measured on B200 (profiles/r01_scan_ring_sweep.txt): 32 KiB x 6 stages, phase 1 three
        // tiles ahead for <= 4-byte prefixes; 8-byte prefixes (twice the shuffle work per byte,
        // two status words per tile) do better with fewer, larger tiles
        if (sizeof(P) == 8) return run_ring<T, P, 49152, 4, 16, 2>(dev, n, inclusive, src, dst, seed, seed_out, peers);
        return run_ring<T, P, 32768, 6, 16, 3>(dev, n, inclusive, src, dst, seed, seed_out, peers);
    }
    if (peers) return fail(HJ_ERR_UNSUPPORTED, "prefix_sum: the fused exchange needs the ring kernel");
    constexpr int VEC = 16 / sizeof(T);
    constexpr size_t TILE = (size_t)SCAN_WARPS * SCAN_ROWS * 32 * VEC;
    size_t n_tiles = (n + TILE - 1) / TILE;
    HJ_REQUIRE(n_tiles < (1ull << 31), "prefix_sum: too many tiles");
    HJ_TRY(ensure_lookback_scratch(dev, n_tiles));
    HJ_TRY(count_epoch(dev));
    LookbackView lb = lookback_view(dev->lookback.base, dev->lookback.capacity_tiles, 0);
    int vec_ok = (((uintptr_t)src | (uintptr_t)dst) & 15u) == 0;
    if (inclusive)
        scan_kernel<T, P, true><<<(unsigned)n_tiles, SCAN_THREADS, 0, dev->stream>>>(
            (const T*)src, (T*)dst, n, (const T*)seed, lb, vec_ok, g_scan_trace);
    else
        scan_kernel<T, P, false><<<(unsigned)n_tiles, SCAN_THREADS, 0, dev->stream>>>(
            (const T*)src, (T*)dst, n, (const T*)seed, lb, vec_ok, g_scan_trace);
    return check_launch(dev, "scan_kernel");
}

}  // namespace

bool prefix_sum_can_fuse_exchange(hj_type_kind ty, size_t n, const void* src, const void* dst) {
    const size_t es = type_size(ty);
    static const int cfg = getenv("HJ_SCAN_CFG") ? atoi(getenv("HJ_SCAN_CFG")) : 1;
    return (((uintptr_t)src | (uintptr_t)dst) & 15u) == 0 && cfg != 0 && n * es >= (64u << 10);
}

hj_status launch_prefix_sum(hj_device* dev, hj_type_kind ty, size_t n, bool inclusive, const void* src,
                            void* dst, const void* seed, void* seed_out, const PeerView* peers) {
    // Integer sums wrap, so signed types run on the unsigned kernel of the same width.
    switch (ty) {
    case HJ_I8: case HJ_U8: return run<uint8_t, uint32_t>(dev, n, inclusive, src, dst, seed, seed_out, peers);
    case HJ_I16: case HJ_U16: return run<uint16_t, uint32_t>(dev, n, inclusive, src, dst, seed, seed_out, peers);
    case HJ_I32: case HJ_U32: return run<uint32_t, uint32_t>(dev, n, inclusive, src, dst, seed, seed_out, peers);
    case HJ_I64: case HJ_U64: return run<uint64_t, uint64_t>(dev, n, inclusive, src, dst, seed, seed_out, peers);
    case HJ_F32: return run<float, float>(dev, n, inclusive, src, dst, seed, seed_out, peers);
    case HJ_F64: return run<double, double>(dev, n, inclusive, src, dst, seed, seed_out, peers);
    default:
        return fail(HJ_ERR_UNSUPPORTED, "prefix_sum: unsupported element type %s", type_name(ty));
    }
}

}  // namespace hj

// Development aid (not part of include/hj.h): record {start, summed, prefix known, end, smid}
// per tile of subsequent scans into a device buffer of 5 * n_tiles u64 (NULL disables).
extern "C" void hj_debug_scan_trace(void* device_ptr) { hj::g_scan_trace = (unsigned long long*)device_ptr; }
