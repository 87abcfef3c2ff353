// Common device helpers for the marllb_b200 kernels (sm_100a).
#pragma once
#include <type_traits>
#include <cuda_runtime.h>
#include <stdint.h>

#define MLB_FULL 0xffffffffu
#define MLB_OBS_COLS 11
#define MLB_INF __int_as_float(0x7f800000)

namespace mlb {

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// total order on floats as unsigned ints (for REDUX-based argmin)
__device__ __forceinline__ uint32_t f32_orderable(float f) {
    uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

__device__ __forceinline__ float f32_from_orderable(uint32_t u) {
    return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MLB_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MLB_FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(MLB_FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(MLB_FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(MLB_FULL, v, o));
    return v;
}

// inclusive warp scan (Kogge-Stone)
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(MLB_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// 2^x, x <= 0: single MUFU.EX2 (2 ulp), denormal results flush to zero
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Ampere-style asynchronous global -> shared copies (LDGSTS); .cg bypasses L1
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// streaming (evict-first) global accesses for data touched once per step
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
    return __ldcs(reinterpret_cast<const float2*>(p));
}
__device__ __forceinline__ float ldg_stream1(const float* p) { return __ldcs(p); }

// compile-time loop: f(std::integral_constant<int, 0>{}), ..., f(std::integral_constant<int, N-1>{}).
// Register arrays indexed inside `#pragma unroll` loops whose bodies hold data-dependent loops stayed in local
// memory (the unroll request was dropped); with constant indices they are scalars.
template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (N > 0) {
        static_for<N - 1>(f);
        f(std::integral_constant<int, N - 1>{});
    }
}

}  // namespace mlb
