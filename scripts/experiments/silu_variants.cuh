// NOT COMPILED INTO THE LIBRARY: record of a round-2 experiment ("tried and removed").
// SFU-free / half-SFU x * sigmoid(x) on packed fp32 pairs, to relieve the MUFU port (2 ops per element: ex2 + rcp) that the
// in-kernel trace and ncu showed busy in the SiLU epilogues. Both are accurate (2e-7 relative, like ex2.approx + rcp.approx) but
// cost 19 / 5 more issue slots per pair, and the kernels in question turned out to be issue- / latency-bound, not MUFU-bound:
//   tcgen05 GEMM epilogue, 25 % of the pairs on the FMA path: K = 16 -> N = 96 layer 1,428 -> 1,238 us, every N > 96 layer 2-5 % slower;
//   fused expand + depthwise kernel: 1,037 -> 1,207 us (half) / 1,375 us (all) per 512 frames; half-SFU variant 1,154 / 1,232 us.
#pragma once
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

// x * sigmoid(x) on a pair with ONE SFU op per element: ex2.approx stays on the MUFU port, the reciprocal is the bit-trick seed
// + three Newton steps of silu2_fma on the FMA pipe (halves the MUFU time for ~5 more issue slots per pair).
__device__ __forceinline__ unsigned long long silu2_half_sfu(unsigned long long x) {
    typedef unsigned long long u64;
    auto pk = [](float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; };
    auto fma2 = [](u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; };
    auto mul2 = [](u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; };
    float t0, t1, e0, e1;
    {
        const u64 t = mul2(x, pk(-1.4426950408889634f, -1.4426950408889634f));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
    }
    t0 = fminf(t0, 125.0f); t1 = fminf(t1, 125.0f);                      // keeps 1 + 2^t inside the seed's range
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
    const u64 nd = fma2(pk(e0, e1), pk(-1.0f, -1.0f), pk(-1.0f, -1.0f));   // -(1 + 2^t)
    unsigned d0, d1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(d0), "=r"(d1) : "l"(nd));
    d0 = 0xFEF311C7u - d0; d1 = 0xFEF311C7u - d1;
    u64 y;
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "r"(d0), "r"(d1));
    const u64 two = pk(2.0f, 2.0f);
    y = mul2(y, fma2(nd, y, two));
    y = mul2(y, fma2(nd, y, two));
    y = mul2(y, fma2(nd, y, two));
    return mul2(x, y);
}

