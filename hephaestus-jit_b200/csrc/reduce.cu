// reduce.cu — single-pass full-array reduction for sm_100a.
//
// Replaces builtin::reduce::reduce (hephaestus-jit/src/backend/vulkan/builtin/reduce.rs:22-314
// + kernels/reduce.glsl): the reference copies the input into a scratch buffer and runs
// ceil(log32 n) radix-32 tree passes of 32-thread workgroups (≈3x the algorithmic bytes).
// Here: ONE launch, a grid sized to the SM count, every thread streams 128-bit
// `ld.global.nc.L1::no_allocate` loads (UNROLL of them in flight), folds them into
// register accumulators, the block folds with warp shuffles, and the last block to finish
// (ticket counter) folds the per-block partials and writes dst[0].
// Algorithmic bytes: sizeof(T) per element, read exactly once; HBM-bound.
//
// (op, type) support and identities follow the reference's table (reduce.rs:84-166);
// pairs it leaves as todo!() return HJ_ERR_UNSUPPORTED.  Integer results are exact
// (wrapping).  Float sums are a fixed (deterministic for a given n and device) but
// different association than the reference's radix-32 tree: tolerance in DESIGN.md.
#include <type_traits>

#include "common.cuh"
#include "hj_internal.h"
#include "peer.cuh"

namespace hj {
namespace {

enum { R_MAX = HJ_REDUCE_MAX, R_MIN = HJ_REDUCE_MIN, R_SUM = HJ_REDUCE_SUM, R_PROD = HJ_REDUCE_PROD,
       R_OR = HJ_REDUCE_OR, R_AND = HJ_REDUCE_AND, R_XOR = HJ_REDUCE_XOR };

constexpr int RED_THREADS = 256;
constexpr int RED_UNROLL = 4;
constexpr int RED_CTAS_PER_SM = 8;

// ---- scalar operator on the (possibly widened) scalar type S --------------------------
__device__ __forceinline__ float fp_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double fp_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float fp_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double fp_min(double a, double b) { return fmin(a, b); }

template <typename S, int OP>
__device__ __forceinline__ S sop(S a, S b) {
    constexpr bool fp = std::is_floating_point<S>::value;
    if constexpr (OP == R_SUM) return (S)(a + b);
    else if constexpr (OP == R_PROD) return (S)(a * b);
    else if constexpr (OP == R_MAX) {
        if constexpr (fp) return fp_max(a, b);
        else return a < b ? b : a;
    } else if constexpr (OP == R_MIN) {
        if constexpr (fp) return fp_min(a, b);
        else return b < a ? b : a;
    } else if constexpr (fp) return a;  // bitwise ops on floats are never dispatched
    else if constexpr (OP == R_OR) return (S)(a | b);
    else if constexpr (OP == R_AND) return (S)(a & b);
    else return (S)(a ^ b);
}

template <typename T> struct Limits;
template <> struct Limits<int8_t> { static constexpr int lo = -128, hi = 127; };
template <> struct Limits<uint8_t> { static constexpr unsigned lo = 0, hi = 255; };
template <> struct Limits<int16_t> { static constexpr int lo = -32768, hi = 32767; };
template <> struct Limits<uint16_t> { static constexpr unsigned lo = 0, hi = 65535; };
template <> struct Limits<int32_t> { static constexpr int32_t lo = INT32_MIN, hi = INT32_MAX; };
template <> struct Limits<uint32_t> { static constexpr uint32_t lo = 0, hi = UINT32_MAX; };
template <> struct Limits<int64_t> { static constexpr int64_t lo = INT64_MIN, hi = INT64_MAX; };
template <> struct Limits<uint64_t> { static constexpr uint64_t lo = 0, hi = UINT64_MAX; };

// identity of OP in the scalar type S for element type T (reduce.rs:84-166)
template <typename T, typename S, int OP>
__device__ __forceinline__ S identity() {
    constexpr bool fp = std::is_floating_point<T>::value;
    if constexpr (OP == R_SUM || OP == R_OR || OP == R_XOR) return (S)0;
    else if constexpr (OP == R_PROD) return (S)1;
    else if constexpr (OP == R_AND) {
        if constexpr (fp) return (S)0;
        else return (S)~(S)0;
    } else if constexpr (OP == R_MAX) {
        if constexpr (fp) return (S)(-INFINITY);
        else return (S)Limits<T>::lo;
    } else {
        if constexpr (fp) return (S)INFINITY;
        else return (S)Limits<T>::hi;
    }
}

// ---- how a 16-byte vector is folded into per-thread accumulators ----------------------
// 4/8-byte types: one accumulator per lane, in the element type.
// 1/2-byte types: the accumulators are packed 32-bit words combined with the SIMD-in-word
// video intrinsics (lane-wise wrapping add / min / max), unpacked once at the end.
template <typename T> struct Scalar { using type = T; };
template <> struct Scalar<int8_t> { using type = int32_t; };
template <> struct Scalar<uint8_t> { using type = uint32_t; };
template <> struct Scalar<int16_t> { using type = int32_t; };
template <> struct Scalar<uint16_t> { using type = uint32_t; };

__device__ __forceinline__ uint32_t vmul4(uint32_t a, uint32_t b) {
    return ((a & 0xffu) * (b & 0xffu) & 0xffu) | ((((a >> 8) & 0xffu) * ((b >> 8) & 0xffu) & 0xffu) << 8) |
           ((((a >> 16) & 0xffu) * ((b >> 16) & 0xffu) & 0xffu) << 16) | (((a >> 24) * (b >> 24) & 0xffu) << 24);
}
__device__ __forceinline__ uint32_t vmul2(uint32_t a, uint32_t b) {
    return ((a & 0xffffu) * (b & 0xffffu) & 0xffffu) | (((a >> 16) * (b >> 16)) << 16);
}

template <typename T, int OP>
__device__ __forceinline__ uint32_t packed_op(uint32_t a, uint32_t b) {
    constexpr bool is8 = sizeof(T) == 1;
    constexpr bool sgn = (T)-1 < (T)0;
    if constexpr (OP == R_SUM) return is8 ? __vadd4(a, b) : __vadd2(a, b);
    else if constexpr (OP == R_PROD) return is8 ? vmul4(a, b) : vmul2(a, b);
    else if constexpr (OP == R_MAX) return is8 ? (sgn ? __vmaxs4(a, b) : __vmaxu4(a, b)) : (sgn ? __vmaxs2(a, b) : __vmaxu2(a, b));
    else if constexpr (OP == R_MIN) return is8 ? (sgn ? __vmins4(a, b) : __vminu4(a, b)) : (sgn ? __vmins2(a, b) : __vminu2(a, b));
    else if constexpr (OP == R_OR) return a | b;
    else if constexpr (OP == R_AND) return a & b;
    else return a ^ b;
}
template <typename T, int OP>
__device__ __forceinline__ uint32_t packed_identity() {
    using S = typename Scalar<T>::type;
    uint32_t lane = (uint32_t)identity<T, S, OP>();
    if constexpr (sizeof(T) == 1) { lane &= 0xffu; return lane * 0x01010101u; }
    else { lane &= 0xffffu; return lane * 0x00010001u; }
}
template <typename T, int OP>
__device__ __forceinline__ typename Scalar<T>::type packed_fold(uint32_t w) {
    using S = typename Scalar<T>::type;
    if constexpr (sizeof(T) == 1) {
        S a = (S)(T)(w & 0xffu), b = (S)(T)((w >> 8) & 0xffu), c = (S)(T)((w >> 16) & 0xffu), d = (S)(T)(w >> 24);
        return sop<S, OP>(sop<S, OP>(a, b), sop<S, OP>(c, d));
    } else {
        S a = (S)(T)(w & 0xffffu), b = (S)(T)(w >> 16);
        return sop<S, OP>(a, b);
    }
}

template <typename T, int OP>
struct Acc {
    using S = typename Scalar<T>::type;
    static constexpr bool packed = sizeof(T) < 4;
    static constexpr int NW = sizeof(T) == 8 ? 2 : 4;
    using W = typename std::conditional<packed, uint32_t, T>::type;
    W w[NW];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < NW; i++) {
            if constexpr (packed) w[i] = packed_identity<T, OP>();
            else w[i] = identity<T, T, OP>();
        }
    }
    __device__ __forceinline__ void add(const uint4& v) {
        if constexpr (packed) {
            w[0] = packed_op<T, OP>(w[0], v.x); w[1] = packed_op<T, OP>(w[1], v.y);
            w[2] = packed_op<T, OP>(w[2], v.z); w[3] = packed_op<T, OP>(w[3], v.w);
        } else if constexpr (sizeof(T) == 4) {
            const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int i = 0; i < 4; i++) w[i] = sop<T, OP>(w[i], e[i]);
        } else {
            const T* e = reinterpret_cast<const T*>(&v);
            w[0] = sop<T, OP>(w[0], e[0]); w[1] = sop<T, OP>(w[1], e[1]);
        }
    }
    __device__ __forceinline__ S fold() const {
        if constexpr (packed) {
            uint32_t a = packed_op<T, OP>(packed_op<T, OP>(w[0], w[1]), packed_op<T, OP>(w[2], w[3]));
            return packed_fold<T, OP>(a);
        } else if constexpr (NW == 4) {
            return sop<T, OP>(sop<T, OP>(w[0], w[1]), sop<T, OP>(w[2], w[3]));
        } else {
            return sop<T, OP>(w[0], w[1]);
        }
    }
};

template <typename S, int OP>
__device__ __forceinline__ S warp_reduce(S v) {
    if constexpr (sizeof(S) == 4 && std::is_integral<S>::value && OP != R_PROD) {
        // 32-bit integer: single-instruction warp reduction (REDUX) on sm_80+
        if constexpr (OP == R_SUM) return (S)__reduce_add_sync(0xffffffffu, v);
        else if constexpr (OP == R_MAX) return (S)__reduce_max_sync(0xffffffffu, v);
        else if constexpr (OP == R_MIN) return (S)__reduce_min_sync(0xffffffffu, v);
        else if constexpr (OP == R_OR) return (S)__reduce_or_sync(0xffffffffu, (unsigned)v);
        else if constexpr (OP == R_AND) return (S)__reduce_and_sync(0xffffffffu, (unsigned)v);
        else return (S)__reduce_xor_sync(0xffffffffu, (unsigned)v);
    } else {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) v = sop<S, OP>(v, shfl_xor(v, m));
        return v;
    }
}

// Result valid in thread 0.
template <typename T, typename S, int OP, int THREADS>
__device__ __forceinline__ S block_reduce(S v, S* smem) {
    v = warp_reduce<S, OP>(v);
    if (lane_id() == 0) smem[warp_id()] = v;
    __syncthreads();
    if (warp_id() == 0) {
        v = lane_id() < THREADS / 32 ? smem[lane_id()] : identity<T, S, OP>();
        v = warp_reduce<S, OP>(v);
    }
    return v;
}

// Called by every thread of the CTA that holds the final local value (valid in thread 0).
// mode == 0: write it.  Otherwise (sharded reduce) this CTA also runs the exchange: the local
// partial goes to every peer's mailbox over NVLink, theirs are collected from the own mailbox, and
// thread 0 folds them IN RANK ORDER — compute and collective in one kernel, deterministic for floats.
// mode == 1: the fold over ALL ranks (sharded reduce); mode == 2: the fold over the ranks BEFORE this
// one (the seed of a sharded scan / compaction).
template <typename T, typename S, int OP>
__device__ __forceinline__ void finish(S v, T* __restrict__ dst, const PeerView& pv, uint32_t mode) {
    if (mode == 0) {
        if (threadIdx.x == 0) dst[0] = (T)v;
        return;
    }
    __shared__ unsigned long long s_mine;
    __shared__ unsigned long long s_all[HJ_MAX_PEERS];
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) {
        T t = (T)v;
        unsigned long long bits = 0;
        memcpy(&bits, &t, sizeof(T));
        s_mine = bits;
        s_epoch = xepoch_begin(pv.xepoch);
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    if ((int)threadIdx.x < pv.world) s_all[threadIdx.x] = peer_exchange(pv, epoch, s_mine, (int)threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        *pv.xepoch = epoch;  // the exchange is complete on this rank
        const int upto = mode == 2 ? pv.rank : pv.world;
        S acc = identity<T, S, OP>();
        for (int q = 0; q < upto; q++) {
            T t;
            memcpy(&t, &s_all[q], sizeof(T));
            acc = q == 0 ? (S)t : sop<S, OP>(acc, (S)t);
        }
        dst[0] = (T)acc;
    }
}

template <typename T, int OP, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS)
reduce_kernel(const T* __restrict__ src, size_t n, size_t head, typename Scalar<T>::type* partials,
              unsigned* ticket, T* __restrict__ dst, PeerView pv, uint32_t mode) {
    using S = typename Scalar<T>::type;
    constexpr int VEC = 16 / sizeof(T);
    __shared__ S smem[THREADS / 32];
    __shared__ bool is_last;
    pdl_launch_dependents();  // PDL: the next kernel's prologue may overlap this kernel's tail ...
    pdl_wait();               // ... and this one reads nothing before the kernel in front has completed

    // [0, head) scalar prologue so that the vector body is 16-byte aligned, then nvec
    // vectors, then a scalar tail.
    const uint4* vsrc = reinterpret_cast<const uint4*>(src + head);
    const size_t nvec = (n - head) / VEC;
    const size_t tail_start = head + nvec * VEC;

    Acc<T, OP> acc;
    acc.init();
    const size_t stride = (size_t)gridDim.x * THREADS;
    size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < nvec; i += UNROLL * stride) {
        uint4 r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) r[u] = ld_stream_v4(vsrc + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc.add(r[u]);
    }
    for (; i < nvec; i += stride) acc.add(ld_stream_v4(vsrc + i));
    S s = acc.fold();

    if (blockIdx.x == 0) {  // ragged ends: < 2*VEC elements in total
        if (threadIdx.x < head) s = sop<S, OP>(s, (S)src[threadIdx.x]);
        size_t t = tail_start + threadIdx.x;
        if (t < n) s = sop<S, OP>(s, (S)src[t]);
    }

    s = block_reduce<T, S, OP, THREADS>(s, smem);
    if (gridDim.x == 1) {
        finish<T, S, OP>(s, dst, pv, mode);
        return;
    }
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s;
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    S v = identity<T, S, OP>();
    for (unsigned b = threadIdx.x; b < gridDim.x; b += THREADS) v = sop<S, OP>(v, partials[b]);
    __syncthreads();  // smem reuse
    v = block_reduce<T, S, OP, THREADS>(v, smem);
    if (threadIdx.x == 0) *ticket = 0;  // ready for the next launch on this stream
    finish<T, S, OP>(v, dst, pv, mode);
}

template <typename T, int OP>
hj_status run(hj_device* dev, size_t n, const void* src, void* dst, const PeerView* peers, uint32_t mode) {
    using S = typename Scalar<T>::type;
    constexpr int VEC = 16 / sizeof(T);
    size_t mis = (size_t)((uintptr_t)src & 15u);
    size_t head = mis ? (16 - mis) / sizeof(T) : 0;
    if (head > n) head = n;
    size_t nvec = (n - head) / VEC;
    size_t want = (nvec + (size_t)RED_THREADS * RED_UNROLL - 1) / ((size_t)RED_THREADS * RED_UNROLL);
    size_t cap = (size_t)dev->sm_count * RED_CTAS_PER_SM;
    int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    HJ_TRY(ensure_reduce_scratch(dev, 64 + cap * sizeof(uint64_t)));
    unsigned* ticket = reinterpret_cast<unsigned*>(dev->reduce_scratch);
    S* partials = reinterpret_cast<S*>(reinterpret_cast<char*>(dev->reduce_scratch) + 64);
    HJ_CUDA(launch_pdl(reduce_kernel<T, OP, RED_THREADS, RED_UNROLL>, dim3(grid), dim3(RED_THREADS), 0, dev->stream,
                       reinterpret_cast<const T*>(src), n, head, partials, ticket, reinterpret_cast<T*>(dst),
                       peers ? *peers : PeerView(), peers ? mode : 0u));
    return check_launch(dev, "reduce_kernel");
}

template <typename T>
hj_status run_arith(hj_device* dev, hj_reduce_op op, size_t n, const void* src, void* dst, const PeerView* pv, uint32_t ep) {
    switch (op) {
    case HJ_REDUCE_MAX: return run<T, R_MAX>(dev, n, src, dst, pv, ep);
    case HJ_REDUCE_MIN: return run<T, R_MIN>(dev, n, src, dst, pv, ep);
    case HJ_REDUCE_SUM: return run<T, R_SUM>(dev, n, src, dst, pv, ep);
    case HJ_REDUCE_PROD: return run<T, R_PROD>(dev, n, src, dst, pv, ep);
    default: return HJ_ERR_UNSUPPORTED;
    }
}
template <typename T>
hj_status run_bitwise(hj_device* dev, hj_reduce_op op, size_t n, const void* src, void* dst, const PeerView* pv, uint32_t ep) {
    switch (op) {
    case HJ_REDUCE_OR: return run<T, R_OR>(dev, n, src, dst, pv, ep);
    case HJ_REDUCE_AND: return run<T, R_AND>(dev, n, src, dst, pv, ep);
    case HJ_REDUCE_XOR: return run<T, R_XOR>(dev, n, src, dst, pv, ep);
    default: return HJ_ERR_UNSUPPORTED;
    }
}
template <typename T>
hj_status run_uint(hj_device* dev, hj_reduce_op op, size_t n, const void* src, void* dst, const PeerView* pv, uint32_t ep) {
    hj_status s = run_arith<T>(dev, op, n, src, dst, pv, ep);
    return s == HJ_ERR_UNSUPPORTED ? run_bitwise<T>(dev, op, n, src, dst, pv, ep) : s;
}

}  // namespace

hj_status launch_reduce(hj_device* dev, hj_reduce_op op, hj_type_kind ty, size_t n, const void* src,
                        void* dst, const PeerView* peers, uint32_t mode) {
    hj_status s = HJ_ERR_UNSUPPORTED;
    switch (ty) {
    case HJ_BOOL: s = run_bitwise<uint8_t>(dev, op, n, src, dst, peers, mode); break;  // reduce.rs:141,150,159
    case HJ_I8: s = run_arith<int8_t>(dev, op, n, src, dst, peers, mode); break;
    case HJ_U8: s = run_uint<uint8_t>(dev, op, n, src, dst, peers, mode); break;
    case HJ_I16: s = run_arith<int16_t>(dev, op, n, src, dst, peers, mode); break;
    case HJ_U16: s = run_uint<uint16_t>(dev, op, n, src, dst, peers, mode); break;
    case HJ_I32: s = run_arith<int32_t>(dev, op, n, src, dst, peers, mode); break;
    case HJ_U32: s = run_uint<uint32_t>(dev, op, n, src, dst, peers, mode); break;
    case HJ_I64: s = run_arith<int64_t>(dev, op, n, src, dst, peers, mode); break;
    case HJ_U64: s = run_uint<uint64_t>(dev, op, n, src, dst, peers, mode); break;
    case HJ_F32: s = run_arith<float>(dev, op, n, src, dst, peers, mode); break;
    case HJ_F64: s = run_arith<double>(dev, op, n, src, dst, peers, mode); break;
    default: break;  // F16: todo!() in the reference
    }
    if (s == HJ_ERR_UNSUPPORTED)
        return fail(HJ_ERR_UNSUPPORTED, "reduce(%s, %s) is not implemented by the reference (todo!())",
                    reduce_op_name(op), type_name(ty));
    return s;
}

}  // namespace hj
