// common.cuh — device helpers shared by the hand-written sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hj {

constexpr int WARP = 32;

// 128-bit streaming load: read-only path, do not allocate in L1 (each byte is touched once).
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// 128-bit streaming store (evict-first in L2: the output is not re-read by this kernel).
__device__ __forceinline__ void st_stream_v4(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// 64-bit relaxed device-scope load/store for the look-back status words.
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64_any(const void* p) {
    return ld_relaxed_u64(reinterpret_cast<const unsigned long long*>(p));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

template <typename T>
__device__ __forceinline__ T shfl_up(T v, int delta) {
    if constexpr (sizeof(T) == 8) {
        unsigned long long u = *reinterpret_cast<unsigned long long*>(&v);
        u = __shfl_up_sync(0xffffffffu, u, delta);
        return *reinterpret_cast<T*>(&u);
    } else {
        return __shfl_up_sync(0xffffffffu, v, delta);
    }
}
template <typename T>
__device__ __forceinline__ T shfl_xor(T v, int mask) {
    if constexpr (sizeof(T) == 8) {
        unsigned long long u = *reinterpret_cast<unsigned long long*>(&v);
        u = __shfl_xor_sync(0xffffffffu, u, mask);
        return *reinterpret_cast<T*>(&u);
    } else {
        return __shfl_xor_sync(0xffffffffu, v, mask);
    }
}
template <typename T>
__device__ __forceinline__ T shfl_idx(T v, int src) {
    if constexpr (sizeof(T) == 8) {
        unsigned long long u = *reinterpret_cast<unsigned long long*>(&v);
        u = __shfl_sync(0xffffffffu, u, src);
        return *reinterpret_cast<T*>(&u);
    } else {
        return __shfl_sync(0xffffffffu, v, src);
    }
}

// Inclusive warp scan (sum) with shuffles.
template <typename T>
__device__ __forceinline__ T warp_inclusive_sum(T v) {
    const int lane = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = shfl_up(v, d);
        if (lane >= d) v = (T)(v + o);
    }
    return v;
}

static inline int div_up(size_t a, size_t b) { return (int)((a + b - 1) / b); }

}  // namespace hj
