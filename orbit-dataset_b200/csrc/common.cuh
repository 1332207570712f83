// Shared device/host helpers for liborbit_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/orbit_b200.h"

#define ORBIT_RETURN_IF_LAUNCH_FAILED()                    \
    do {                                                   \
        cudaError_t err__ = cudaGetLastError();            \
        if (err__ != cudaSuccess) return (int)err__;       \
    } while (0)

#define ORBIT_CUDA(call)                                   \
    do {                                                   \
        cudaError_t err__ = (call);                        \
        if (err__ != cudaSuccess) return (int)err__;       \
    } while (0)

namespace orbit {

constexpr int kWarp = 32;

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Parity-grade activations: full-precision expf and IEEE division (the reference computes
// x*sigmoid(x) in fp32 on the CPU). The memory-bound kernels that use them have the ALU headroom.
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x / (1.0f + expf(-x)); }

// x * sigmoid(x) on the SFU: ex2.approx.ftz + rcp.approx.ftz (2 MUFU + 3 FP32 ops, ~3e-7 relative error). The
// CUDA intrinsics __expf/__fdividef wrap the same two instructions in denormal-range fix-ups (~6 more issue slots
// per element) that x*sigmoid(x) does not need: e underflows to 0 (result x) or overflows to inf (result -0).
__device__ __forceinline__ float silu_sfu(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming (read-once) 128-bit load that does not allocate in L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void fma4(float4& a, const float4& x, const float4& w) {
    a.x = fmaf(x.x, w.x, a.x); a.y = fmaf(x.y, w.y, a.y);
    a.z = fmaf(x.z, w.z, a.z); a.w = fmaf(x.w, w.w, a.w);
}
__device__ __forceinline__ void add4(float4& a, const float4& x) { a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace orbit
