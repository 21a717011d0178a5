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
#include <type_traits>

#include "common.cuh"
#include "hj_internal.h"

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

template <typename T>
__global__ void __launch_bounds__(256)
gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ idx, T* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) dst[i] = __ldg(src + idx[i]);
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
                                size_t n_dst) {
    if (n == 0) return HJ_OK;
    hj_status s = HJ_ERR_UNSUPPORTED;
    if (op == HJ_REDUCE_PROD)  // todo!() in the reference (glsl/mod.rs:422)
        return fail(HJ_ERR_UNSUPPORTED, "scatter_reduce(Prod) is not implemented by the reference");
    // privatised shared-memory histogram for small bin counts
    if (op == HJ_REDUCE_SUM && (ty == HJ_U32 || ty == HJ_I32) && !src && n_dst * 4 <= 64 * 1024 &&
        n >= (1u << 16)) {
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
    case 4: gather_kernel<uint32_t><<<grid, 256, 0, dev->stream>>>((const uint32_t*)src, idx, (uint32_t*)dst, n); break;
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
