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

// x * sigmoid(x) on a packed fp32 PAIR without the SFU. Two MUFU ops per element (ex2 + rcp, 8 issue-port clocks each per
// warp) make SiLU the limiter of the epilogues that finish a whole tile in a few warps (ncu / in-kernel trace, round 2:
// a 128 x 96 SiLU tile keeps each SM sub-partition's MUFU port busy for 1,536 clocks). This version runs on the FMA and
// integer pipes only, so call sites can send a fraction of their pairs here and balance the two:
//   2^t, t = -x log2(e) clamped to [-125, 125]:  n = round(t) by the 1.5 * 2^23 magic add, f = t - n in [-0.5, 0.5],
//        degree-5 polynomial (max relative error 2.0e-7 evaluated in fp32: the same as ex2.approx), exponent by integer add;
//   1 / d, d = 1 + 2^t:  y0 = bit trick (5 % off), three Newton steps y <- y (2 - d y) -> 1.2e-7 (1 ulp, as rcp.approx).
__device__ __forceinline__ unsigned long long silu2_fma(unsigned long long x) {
    typedef unsigned long long u64;
    auto pk = [](float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; };
    auto fma2 = [](u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; };
    auto add2 = [](u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; };
    auto mul2 = [](u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; };
    float t0, t1;
    {
        const u64 t = mul2(x, pk(-1.4426950408889634f, -1.4426950408889634f));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
    }
    t0 = fminf(fmaxf(t0, -125.0f), 125.0f);
    t1 = fminf(fmaxf(t1, -125.0f), 125.0f);
    const u64 t = pk(t0, t1);
    const u64 r = add2(t, pk(12582912.0f, 12582912.0f));                  // integer part in the low mantissa bits
    const u64 nf = add2(r, pk(-12582912.0f, -12582912.0f));
    const u64 f = fma2(nf, pk(-1.0f, -1.0f), t);
    u64 p = fma2(pk(1.326697064e-03f, 1.326697064e-03f), f, pk(9.675459936e-03f, 9.675459936e-03f));
    p = fma2(p, f, pk(5.550742522e-02f, 5.550742522e-02f));
    p = fma2(p, f, pk(2.402212173e-01f, 2.402212173e-01f));
    p = fma2(p, f, pk(6.931469440e-01f, 6.931469440e-01f));
    p = fma2(p, f, pk(1.000000119e+00f, 1.000000119e+00f));
    unsigned r0, r1, p0, p1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(r0), "=r"(r1) : "l"(r));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(p));
    p0 += r0 << 23; p1 += r1 << 23;                                      // * 2^n (the magic constant's bits shift out)
    u64 e;
    asm("mov.b64 %0, {%1, %2};" : "=l"(e) : "r"(p0), "r"(p1));
    const u64 nd = fma2(e, pk(-1.0f, -1.0f), pk(-1.0f, -1.0f));           // -(1 + 2^t)
    unsigned d0, d1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(d0), "=r"(d1) : "l"(nd));
    d0 = 0xFEF311C7u - d0; d1 = 0xFEF311C7u - d1;                        // 0x7EF311C7 - bits(d), bits(d) = bits(nd) ^ 0x80000000
    u64 y;
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "r"(d0), "r"(d1));
    const u64 two = pk(2.0f, 2.0f);
    y = mul2(y, fma2(nd, y, two));
    y = mul2(y, fma2(nd, y, two));
    y = mul2(y, fma2(nd, y, two));
    return mul2(x, y);
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
