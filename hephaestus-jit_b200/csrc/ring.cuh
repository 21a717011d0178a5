// ring.cuh — persistent "reduce-ahead" tile pipeline shared by the scan and compress kernels.
//
// The reference's scan/compress (kernels/prefix_sum_large.glsl:165-359,
// compress_large.glsl:76-229) are one-workgroup-per-2048-items decoupled look-back kernels: a
// partition loads its items, publishes its aggregate, then spins walking back over its
// predecessors' status words while it HOLDS its items in registers.  On B200 that wait (several
// loaded-fabric L2 round trips) sits on the critical path of every tile (profiles/r01_*: 48 % of
// the HBM roofline register-resident, 70 % with an L2-resident second sweep that re-fetches 40 %
// of the input from DRAM).  This skeleton takes the wait off the critical path instead:
//
//   * one persistent CTA per SM; tiles are handed out by an atomic ticket (so a tile only ever
//     depends on tiles whose CTA is already running — no co-residency assumption);
//   * a producer lane streams each tile into a shared-memory ring with ONE TMA bulk copy
//     (cp.async.bulk + mbarrier complete_tx): no registers, no LSU queue slots are held by data
//     in flight, and the number of bytes in flight per SM is bounded by the ring;
//   * consumer warps run two phases per tile, software-pipelined AHEAD tiles apart:
//       phase 1 (tile i+AHEAD): reduce the tile from shared memory, publish the tile aggregate;
//       phase 2 (tile i)      : re-read the tile from shared memory, apply the exclusive prefix,
//                                write the result to HBM;
//   * a prefix warp turns published aggregates into tile prefixes.  Because aggregates are
//     published AHEAD tiles (several microseconds) before they are needed, it never has to
//     wait in steady state, and no "inclusive" status is ever published: tiles are grouped in
//     rounds of G consecutive tickets, prefix(t) = (sum of all complete rounds before t's
//     round) + (sum of the aggregates of the tiles before t in its round); each CTA keeps the
//     running sum of complete rounds in a register.
//
// HBM traffic is exactly the algorithmic bytes (every input byte is fetched once, by TMA; every
// output byte written once); the status-word polling is L2-resident and ~2 % of the traffic.
#pragma once
#include "lookback.cuh"

namespace hj {

constexpr uint32_t RING_END = 0xffffffffu;  // s_tile value: no more tiles

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(const void* p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "r"(smem_u32(p)));
    return r;
}

// Sums of the published aggregates of tiles [a, mid) and [mid, b) in ONE sweep — executed by one
// full warp, results in every lane.  Up to 32*M status words are fetched per round trip; words
// not yet published are re-polled (with back-off) until they are.
template <typename P, int M>
__device__ __forceinline__ void ring_sum_ranges(const LookbackView& lb, uint32_t a, uint32_t mid, uint32_t b,
                                                P* sum_lo, P* sum_hi) {
    const int lane = lane_id();
    P lo = (P)0, hi = (P)0;
    for (uint32_t base = a; base < b; base += 32 * M) {
        P v[M];
        uint32_t st[M];
#pragma unroll
        for (int m = 0; m < M; m++) {
            const uint32_t u = base + m * 32 + lane;
            v[m] = (P)0;
            st[m] = TILE_AGGREGATE;
            if (u < b) st[m] = tile_read<P>(lb, u, &v[m]);
        }
        // anything not yet published: re-poll all of them together (independent loads), backing off
        unsigned ns = 32;
        while (true) {
            bool pending = false;
#pragma unroll
            for (int m = 0; m < M; m++) pending |= (st[m] == TILE_INVALID);
            if (!__any_sync(0xffffffffu, pending)) break;
            __nanosleep(ns);
            if (ns < 256) ns *= 2;
#pragma unroll
            for (int m = 0; m < M; m++)
                if (st[m] == TILE_INVALID) st[m] = tile_read<P>(lb, base + m * 32 + lane, &v[m]);
        }
#pragma unroll
        for (int m = 0; m < M; m++) {
            const uint32_t u = base + m * 32 + lane;
            if (u < mid) lo = (P)(lo + v[m]);
            else hi = (P)(hi + v[m]);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        lo = (P)(lo + shfl_xor(lo, s));
        hi = (P)(hi + shfl_xor(hi, s));
    }
    *sum_lo = lo;
    *sum_hi = hi;
}

// Shared-memory control block of the ring (placed after the data stages).  Two rings: the DATA
// stages the TMA fills (full / empty), and the TILE slots that carry a tile's bookkeeping from
// phase 1 to phase 2 (agg / pub / pref and the per-warp sums and offsets).  With EARLY release
// (phase 2 does not need the tile's bytes any more, e.g. compress keeps only bit masks) a data
// stage goes back to the producer right after phase 1, so every stage but one is in flight and
// the number of tile slots — not the number of stages — bounds how far phase 1 runs ahead.
template <typename P, int STAGES, int TSLOTS, int CWARPS>
struct RingCtl {
    uint64_t full[STAGES];    // producer -> consumers: tile landed (tx-count barrier)
    uint64_t empty[STAGES];   // consumers -> producer: stage may be overwritten
    uint64_t agg[TSLOTS];     // consumers -> publisher warp: warp totals written
    uint64_t pub[TSLOTS];     // publisher warp -> prefix warp: aggregate published, woff written
    uint64_t pref[TSLOTS];    // prefix warp -> consumers: tile prefix written
    uint32_t stile[STAGES];   // ticket of the tile in each data stage, RING_END after the last one
    uint32_t ttile[TSLOTS][CWARPS];  // ticket of the tile in each tile slot, one private copy per consumer warp
                                     // (a fast warp may re-enter a slot while a slow one still reads its own)
    P wsum[TSLOTS][CWARPS];   // per-warp totals of the tile (phase 1)
    P woff[TSLOTS][CWARPS];   // offset of each warp's slice inside the tile (publisher warp)
    P tagg[TSLOTS];           // tile aggregate
    P tpre[TSLOTS];           // exclusive prefix of the tile (prefix warp)
    uint32_t epoch;           // this launch's status-word epoch (lookback.cuh: epoch_begin)
};

// The kernel body.  `Op` supplies the element-level work:
//   using P                                   prefix type (u32 / u64 / float / double)
//   static P    total(slice, aux, slot, cw, lane)
//                                              phase 1: this warp's total of its SLICE bytes; may park
//                                              per-tile data in `aux` (shared memory after the ring)
//   static void emit(slice, aux, slot, byte_off, valid_bytes, carry, slice_total, lane, cw, args)
//                                              phase 2: write the outputs of the slice that starts
//                                              `byte_off` bytes into the input (slice is only valid
//                                              without EARLY release)
//   static void finish(total, args, lane)     called once, by the prefix warp (all 32 lanes) of the CTA
//                                              that owns the last tile; `total` includes the seed
// TILE bytes per stage, STAGES data stages, TSLOTS tile slots, CWARPS consumer warps; phase 1 runs
// AHEAD tiles ahead of phase 2.
// TRACE (development aid, tools/ring_timeline.py): globaltimer stamps per tile in `trace` (10 words per tile).
template <class Op, int TILE, int STAGES, int CWARPS, int AHEAD, int TSLOTS = STAGES, bool EARLY = false, int SWEEP_M = 8,
          bool TRACE = false>
__device__ __forceinline__ void ring_pipeline(const char* __restrict__ src, size_t n_bytes, uint32_t n_tiles,
                                              typename Op::P seed, LookbackView lb, uint32_t G,
                                              const typename Op::Args& args, char* smem,
                                              unsigned long long* trace = nullptr) {
    using P = typename Op::P;
    using Ctl = RingCtl<P, STAGES, TSLOTS, CWARPS>;
    static_assert(AHEAD >= 1 && AHEAD < TSLOTS, "a tile slot must outlive the AHEAD tiles between its two phases");
    static_assert(EARLY || AHEAD < STAGES, "without early release the data must stay in the ring until phase 2");
    constexpr int SLICE = TILE / CWARPS;  // bytes of a tile owned by one consumer warp
    static_assert(SLICE % 512 == 0, "a warp slice is a whole number of 512-byte rows");
    char* stages = smem;
    Ctl* ctl = reinterpret_cast<Ctl*>(smem + (size_t)STAGES * TILE);
    char* aux = smem + (size_t)STAGES * TILE + ((sizeof(Ctl) + 127) & ~(size_t)127);
    const int warp = warp_id(), lane = lane_id();
    // stamp k of tile t: 0 drawn, 1 phase 1 starts, 2 phase 1 done, 3 aggregate published,
    // 4 sweep starts, 5 prefix handed over, 6 phase 2 starts, 7 phase 2 done (consumer warp 0), 8 CTA
    auto stamp = [&](uint32_t t, int k) {
        if constexpr (TRACE) {
            if (trace && t != RING_END) trace[(size_t)t * 10 + k] = k == 8 ? (unsigned long long)blockIdx.x : globaltimer_ns();
        }
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&ctl->full[s], 1);
            mbar_init(&ctl->empty[s], CWARPS);
        }
        for (int q = 0; q < TSLOTS; q++) {
            mbar_init(&ctl->agg[q], CWARPS);
            mbar_init(&ctl->pub[q], 1);
            mbar_init(&ctl->pref[q], 1);
        }
        // everything above is local to the CTA: it overlaps the tail of the kernel in front (PDL)
        pdl_launch_dependents();
        pdl_wait();
        ctl->epoch = epoch_begin(lb);
    }
    pdl_wait();  // every thread: no global access before the kernel in front has completed
    __syncthreads();
    lb.epoch = ctl->epoch;

    if (warp == 0) {
        // ---------------- producer: ticket -> TMA bulk copy of the tile into the next stage.
        // The ticket atomic is an L2 round trip (~0.5 us loaded); drawn one at a time it caps the
        // CTA at one tile per round trip (measured: 0.66 us fixed cost per tile).  TICKETS draws
        // are therefore kept in flight: the value consumed now was requested TICKETS tiles ago.
        constexpr int TICKETS = 1;
        const uint32_t last_draw = n_tiles + gridDim.x * TICKETS - 1;  // every CTA draws its tiles + TICKETS
        uint32_t tq[TICKETS];
#pragma unroll
        for (int i = 0; i < TICKETS; i++) tq[i] = lane == 0 ? atomicAdd(lb.ticket, 1u) : 0u;
        for (uint32_t it = 0;; it++) {
            const int s = it % STAGES;
            const uint32_t use = it / STAGES;
            uint32_t t = tq[0];
#pragma unroll
            for (int i = 0; i + 1 < TICKETS; i++) tq[i] = tq[i + 1];
            if (lane == 0 && t == last_draw) {  // the very last draw re-arms the counter, advances the epoch
                *lb.ticket = 0;
                epoch_advance(lb, lb.epoch);
            }
            if (use > 0) mbar_wait(&ctl->empty[s], (use - 1) & 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= n_tiles) {
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i + 1 < TICKETS; i++)
                        if (tq[i] == last_draw) {
                            *lb.ticket = 0;
                            epoch_advance(lb, lb.epoch);
                        }
                    ctl->stile[s] = RING_END;
                    mbar_arrive(&ctl->full[s]);
                }
                break;
            }
            tq[TICKETS - 1] = lane == 0 ? atomicAdd(lb.ticket, 1u) : 0u;
            char* dst = stages + (size_t)s * TILE;
            const size_t off = (size_t)t * TILE;
            const size_t left = n_bytes - off;
            const uint32_t bytes = left < (size_t)TILE ? (uint32_t)left : (uint32_t)TILE;
            const uint32_t bulk = bytes & ~15u;
            if (bytes < (uint32_t)TILE) {
                // ragged last tile: the tail of the stage reads as zero (identity of the sum)
                for (uint32_t b = bulk + lane * 4; b < (uint32_t)TILE; b += 128)
                    *reinterpret_cast<uint32_t*>(dst + b) = 0u;
                __syncwarp();
                for (uint32_t b = bulk + lane; b < bytes; b += 32) dst[b] = src[off + b];
                __syncwarp();
            }
            if (lane == 0) {
                stamp(t, 0);
                stamp(t, 8);
                ctl->stile[s] = t;
                if (bulk) {
                    mbar_expect_tx(&ctl->full[s], bulk);
                    tma_load_1d(dst, src + off, bulk, &ctl->full[s]);
                } else {
                    mbar_arrive(&ctl->full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- publisher warp: warp totals -> tile aggregate, published at once.
        // It never waits on another CTA, so aggregates appear AHEAD tiles before they are needed.
        int q = 0;
        uint32_t par = 0;
        for (;;) {
            mbar_wait(&ctl->agg[q], par);
            const uint32_t t = ctl->ttile[q][0];
            if (t == RING_END) {
                if (lane == 0) mbar_arrive(&ctl->pub[q]);
                break;
            }
            P w = lane < CWARPS ? ctl->wsum[q][lane] : (P)0;
            P inc = warp_inclusive_sum(w);
            const P aggregate = shfl_idx(inc, 31);
            if (lane == 0) {
                tile_publish<P>(lb, t, TILE_AGGREGATE, aggregate);
                ctl->tagg[q] = aggregate;
                stamp(t, 3);
            }
            if (lane < CWARPS) ctl->woff[q][lane] = (P)(inc - w);
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->pub[q]);
            if (++q == TSLOTS) { q = 0; par ^= 1; }
        }
    } else if (warp == 2) {
        // ---------------- prefix warp: published aggregates -> exclusive prefix of the tile.
        // Tiles are grouped in rounds of G consecutive tickets: prefix(t) = (sum of all complete
        // rounds before t's round, kept in a register) + (aggregates of the tiles before t in its
        // round).  [next_round*G, k*G) — rounds completed since this CTA's previous tile — and
        // [k*G, t) are contiguous, so one sweep (one L2 round trip for up to 256 status words)
        // serves both.  Variants measured and dropped (profiles/r01_ring_sweeps.txt): summing only
        // the tickets between this CTA's consecutive tiles (f32 scan 10 % slower), two sweeps in
        // flight (every prefix arrives a tile period later: scan 12 % slower).
        P rounds_total = seed;
        uint32_t next_round = 0;
        int q = 0, sd = 0;     // tile slot; data stage of the tile (only tracked without EARLY release)
        uint32_t par = 0, spar = 0;
        for (;;) {
            // Without EARLY release the data stage still names its tile when the prefix warp gets
            // to it, so the sweep can start as soon as the tile has landed and overlap phase 1 and
            // the publisher; with EARLY release the ticket is only safe to read after `pub`.
            uint32_t t;
            if (!EARLY) {
                mbar_wait(&ctl->full[sd], spar);
                t = ctl->stile[sd];
                if (++sd == STAGES) { sd = 0; spar ^= 1; }
            } else {
                mbar_wait(&ctl->pub[q], par);
                t = ctl->ttile[q][0];
            }
            if (t == RING_END) break;
            const uint32_t k = t / G;
            if (lane == 0) stamp(t, 4);
            P full_rounds = (P)0, partial = (P)0;
            ring_sum_ranges<P, SWEEP_M>(lb, next_round * G, k * G, t, &full_rounds, &partial);
            rounds_total = (P)(rounds_total + full_rounds);
            next_round = k;
            const P exclusive = (P)(rounds_total + partial);
            if (!EARLY) mbar_wait(&ctl->pub[q], par);
            if (lane == 0) {
                ctl->tpre[q] = exclusive;
                stamp(t, 5);
                mbar_arrive(&ctl->pref[q]);
            }
            // the whole warp: a sharded launch runs its cross-GPU exchange here (one lane per peer),
            // after the last tile's consumers have been released
            if (t == n_tiles - 1) Op::finish((P)(exclusive + ctl->tagg[q]), args, lane);
            if (++q == TSLOTS) { q = 0; par ^= 1; }
        }
    } else {
        // ---------------- consumers: phase 1 of tile it, phase 2 of tile it - AHEAD
        const int cw = warp - 3;
        uint32_t n_iter = 0xffffffffu;  // number of real tiles of this CTA, known at RING_END
        // (stage / slot, parity) advance incrementally: no division by STAGES
        int s1 = 0, s2 = 0, q1 = 0, q2 = 0;
        uint32_t spar1 = 0, qpar2 = 0;
        for (uint32_t it = 0;; it++) {
            if (n_iter == 0xffffffffu) {
                mbar_wait(&ctl->full[s1], spar1);
                const uint32_t t = ctl->stile[s1];
                P total = (P)0;
                if (t == RING_END) {
                    n_iter = it;
                } else {
                    if (cw == 0 && lane == 0) stamp(t, 1);
                    total = Op::total(stages + (size_t)s1 * TILE + (size_t)cw * SLICE, aux, q1, cw, lane);
                    if (EARLY) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&ctl->empty[s1]);
                    }
                }
                // every warp arrives (also at RING_END, so the control warps wake up and stop)
                if (lane == 0) {
                    if (cw == 0) stamp(t, 2);
                    ctl->ttile[q1][cw] = t;
                    ctl->wsum[q1][cw] = total;
                    mbar_arrive(&ctl->agg[q1]);
                }
                if (++s1 == STAGES) { s1 = 0; spar1 ^= 1; }
                if (++q1 == TSLOTS) q1 = 0;
            }
            if (it >= (uint32_t)AHEAD) {
                if (it - AHEAD >= n_iter) break;
                mbar_wait(&ctl->pref[q2], qpar2);
                const uint32_t t = ctl->ttile[q2][cw];
                if (cw == 0 && lane == 0) stamp(t, 6);
                const size_t byte_off = (size_t)t * TILE + (size_t)cw * SLICE;
                const size_t valid = byte_off < n_bytes ? n_bytes - byte_off : 0;
                Op::emit(stages + (size_t)s2 * TILE + (size_t)cw * SLICE, aux, q2, byte_off,
                         valid < (size_t)SLICE ? (uint32_t)valid : (uint32_t)SLICE,
                         (P)(ctl->tpre[q2] + ctl->woff[q2][cw]), ctl->wsum[q2][cw], lane, cw, args);
                if (TRACE && cw == 0) {
                    __syncwarp();
                    if (lane == 0) stamp(t, 7);
                }
                if (!EARLY) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ctl->empty[s2]);
                }
                if (++s2 == STAGES) s2 = 0;
                if (++q2 == TSLOTS) { q2 = 0; qpar2 ^= 1; }
            }
        }
    }
}

template <typename P, int TILE, int STAGES, int CWARPS, int TSLOTS = STAGES>
constexpr size_t ring_smem_bytes(size_t aux_bytes = 0) {
    return (size_t)STAGES * TILE + ((sizeof(RingCtl<P, STAGES, TSLOTS, CWARPS>) + 127) & ~(size_t)127) + aux_bytes;
}

}  // namespace hj
