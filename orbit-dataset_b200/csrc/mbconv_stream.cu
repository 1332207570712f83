// Fused MBConv front half, register-resident variant for the 3x3 STRIDE-2 block with 16 input channels (EfficientNet-B0
// block 1.0: 112x112x16 -> expand to 96 -> depthwise 3x3 s2 -> 56x56x96): expand 1x1 + bn1 + SiLU -> depthwise + bn2 (FiLM
// site, model/film.py:43-44) + SiLU + SE squeeze sums, with the expanded tensor living in REGISTERS, not shared memory.
//
// Reference op sites: timm InvertedResidual conv_pw/bn1 -> conv_dw/bn2 inside the extractor invoked at
// model/few_shot_recognisers.py:114-117,143-146.
//
// mbx_kernel (convnet.cu) stages input rows and the expanded rows through shared memory and is instruction-bound (ncu: 55 %
// issue, XU 42 %, 33 thread-instructions per expanded value). This kernel applies what made the row-streaming GEMM fast:
//   * a warp owns 16 output channels (two 8-column mma tiles) of a band of output rows and walks the expanded rows;
//   * A fragments come straight from global memory with 128-bit loads: the 16 rows of an mma tile are 16 CONSECUTIVE pixels
//     of an input row, mapped row g <-> pixel 16 j + 2 g (even), row g + 8 <-> pixel 16 j + 2 g + 1 (odd), and the 16 input
//     channels are ONE k-step (lane t holds channels 4t .. 4t+3 in its four k slots; the weights use the same permutation);
//   * FP16x3 split as everywhere (hi.hi into a fresh accumulator, hi.lo + lo.hi scaled by 2^-11), bn1 + SiLU on the C
//     fragment: lane (g, t) then holds, for channels 2t, 2t+1 of each tile, the expanded pixels 2 ox and 2 ox + 1 of ITS output
//     column ox = 8 j + g -- exactly two of the three taps of a stride-2 3x3 window; the third (2 ox + 2) is the even pixel of
//     lane g + 1 (one shuffle; for g = 7 the next tile's g = 0);
//   * the depthwise accumulators of two output rows (the row finishing and the row starting at an even expanded row) stay
//     in registers; bn2 + SiLU + store + SE sums when a row completes.
// TF "SAME" padding of a 3x3 stride-2 conv on an even size: nothing above / left, one zero row / column below / right; the
// padding applies to the EXPANDED tensor (zeros, not silu(bn1(0))).
#include "convnet.cuh"

namespace orbit {
namespace mbs {

typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) { f2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t f2_silu(f2_t x) {
    float t0, t1, e0, e1, r0, r1;
    f2_unpack(f2_mul(x, f2_pack(-1.4426950408889634f, -1.4426950408889634f)), t0, t1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
    f2_unpack(f2_add(f2_pack(e0, e1), f2_pack(1.0f, 1.0f)), t0, t1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(t0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(t1));
    return f2_mul(x, f2_pack(r0, r1));
}
__device__ __forceinline__ void split_f16x2(f2_t x, uint32_t& hi, uint32_t& lo) {
    float x0, x1, h0, h1, r0, r1;
    f2_unpack(x, x0, x1);
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
    const f2_t r = f2_fma(f2_pack(h0, h1), f2_pack(-2048.0f, -2048.0f), f2_mul(x, f2_pack(2048.0f, 2048.0f)));
    f2_unpack(r, r0, r1);
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, const float (&c)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
__device__ __forceinline__ f2_t shfl_f2(f2_t v, int src_lane) {
    return __shfl_sync(0xffffffffu, v, src_lane);
}

constexpr int kCin = 16;
constexpr int kTilesMax = 7;          // 16-pixel tiles per input row: W <= 112

// grid (row bands, frames); block = (C / 16) warps. W = 16 * TILES, H even; Wo = W / 2, Ho = H / 2.
template <int TILES>
__global__ void __launch_bounds__(192, 2)
mbs_kernel(const float* __restrict__ x, const float* __restrict__ we, const float* __restrict__ scale1, const float* __restrict__ shift1,
           const float* __restrict__ wt, const float* __restrict__ scale2, const float* __restrict__ shift2, float* __restrict__ y,
           float* __restrict__ partial, int H, int C, int rows_per_band) {
    constexpr int W = 16 * TILES, Wo = 8 * TILES;
    const int Ho = H >> 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.y, band = blockIdx.x, bands = gridDim.x;
    const int oy0 = band * rows_per_band, oy1 = min(Ho, oy0 + rows_per_band);
    const int n0 = warp * 16;                                  // this warp's 16 channels: mma tiles n0 .. n0+7 and n0+8 .. n0+15
    // ---- per-lane constants: weight fragments (channel n0 + 8 jn + g, input channels 4t .. 4t+3), bn1 / bn2 of channels 2t, 2t+1
    uint32_t bh[2][2], bl[2][2];
    f2_t s1[2], h1[2], s2[2], h2[2];
#pragma unroll
    for (int jn = 0; jn < 2; ++jn) {
        const float4 wv = ldg4(we + (int64_t)(n0 + 8 * jn + g) * kCin + 4 * t);
        split_f16x2(f2_pack(wv.x, wv.y), bh[jn][0], bl[jn][0]);
        split_f16x2(f2_pack(wv.z, wv.w), bh[jn][1], bl[jn][1]);
        const int c = n0 + 8 * jn + 2 * t;
        s1[jn] = __ldg(reinterpret_cast<const f2_t*>(scale1 + c)); h1[jn] = __ldg(reinterpret_cast<const f2_t*>(shift1 + c));
        s2[jn] = __ldg(reinterpret_cast<const f2_t*>(scale2 + c)); h2[jn] = __ldg(reinterpret_cast<const f2_t*>(shift2 + c));
    }
    const float* xb = x + (int64_t)b * H * W * kCin;
    float* yb = y + (int64_t)b * Ho * Wo * C;
    f2_t acc_cur[TILES][2], acc_nxt[TILES][2];                 // output row being finished / the next one (channel pairs)
#pragma unroll
    for (int j = 0; j < TILES; ++j)
#pragma unroll
        for (int jn = 0; jn < 2; ++jn) { acc_cur[j][jn] = 0ull; acc_nxt[j][jn] = 0ull; }
    f2_t sum[2] = {0ull, 0ull};
    const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
    const f2_t inv = f2_pack(4.8828125e-4f, 4.8828125e-4f);   // 2^-11

    // one expanded row e (= input row e): expand its TILES tiles, feed the depthwise accumulators.
    //   ky_cur >= 0: this row is tap row ky_cur of the output row held in acc_cur;  ky_nxt >= 0: ... of acc_nxt
    auto expanded_row = [&](int e, int ky_cur, int ky_nxt) {
        if (e >= H) return;                                    // the zero row below the image
        const float* xr = xb + (int64_t)e * W * kCin + 4 * t;
        f2_t wc[2][3], wn[2][3];                               // depthwise taps of this row for the two channel pairs
#pragma unroll
        for (int jn = 0; jn < 2; ++jn)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int c = n0 + 8 * jn + 2 * t;
                wc[jn][kx] = ky_cur >= 0 ? __ldg(reinterpret_cast<const f2_t*>(wt + (int64_t)(ky_cur * 3 + kx) * C + c)) : 0ull;
                wn[jn][kx] = ky_nxt >= 0 ? __ldg(reinterpret_cast<const f2_t*>(wt + (int64_t)(ky_nxt * 3 + kx) * C + c)) : 0ull;
            }
        f2_t pe[2] = {0ull, 0ull}, po[2] = {0ull, 0ull};       // previous tile's even / odd expanded pixels
#pragma unroll
        for (int j = 0; j <= TILES; ++j) {
            f2_t ce[2] = {0ull, 0ull}, co[2] = {0ull, 0ull};   // this tile's (j == TILES: the zero column right of the image)
            if (j < TILES) {
                const float4 va = ldg4(xr + (int64_t)(16 * j + 2 * g) * kCin);          // pixel 16 j + 2 g      (mma row g)
                const float4 vb = ldg4(xr + (int64_t)(16 * j + 2 * g + 1) * kCin);      // pixel 16 j + 2 g + 1  (mma row g + 8)
                uint32_t ah[4], al[4];
                split_f16x2(f2_pack(va.x, va.y), ah[0], al[0]); split_f16x2(f2_pack(vb.x, vb.y), ah[1], al[1]);
                split_f16x2(f2_pack(va.z, va.w), ah[2], al[2]); split_f16x2(f2_pack(vb.z, vb.w), ah[3], al[3]);
#pragma unroll
                for (int jn = 0; jn < 2; ++jn) {
                    float m4[4], cor[4];
                    mma_f16(m4, ah, bh[jn][0], bh[jn][1], zero4);
                    mma_f16(cor, al, bh[jn][0], bh[jn][1], zero4);
                    mma_f16(cor, ah, bl[jn][0], bl[jn][1], cor);
                    ce[jn] = f2_silu(f2_fma(f2_fma(f2_pack(cor[0], cor[1]), inv, f2_pack(m4[0], m4[1])), s1[jn], h1[jn]));
                    co[jn] = f2_silu(f2_fma(f2_fma(f2_pack(cor[2], cor[3]), inv, f2_pack(m4[2], m4[3])), s1[jn], h1[jn]));
                }
            }
            if (j > 0) {                                       // finish tile j - 1: its third tap column is lane g + 1's even pixel
#pragma unroll
                for (int jn = 0; jn < 2; ++jn) {
                    const f2_t from_next_lane = shfl_f2(pe[jn], lane + 4);
                    const f2_t from_next_tile = shfl_f2(ce[jn], t);
                    const f2_t third = g < 7 ? from_next_lane : from_next_tile;
                    if (ky_cur >= 0)
                        acc_cur[j - 1][jn] = f2_fma(third, wc[jn][2], f2_fma(po[jn], wc[jn][1], f2_fma(pe[jn], wc[jn][0], acc_cur[j - 1][jn])));
                    if (ky_nxt >= 0)
                        acc_nxt[j - 1][jn] = f2_fma(third, wn[jn][2], f2_fma(po[jn], wn[jn][1], f2_fma(pe[jn], wn[jn][0], acc_nxt[j - 1][jn])));
                }
            }
#pragma unroll
            for (int jn = 0; jn < 2; ++jn) { pe[jn] = ce[jn]; po[jn] = co[jn]; }
        }
    };
    // bn2 + SiLU + store + SE sums of the finished output row, then the next row's accumulators become the current ones
    auto finish_row = [&](int oy) {
        float* yr = yb + (int64_t)oy * Wo * C + n0 + 2 * t;
#pragma unroll
        for (int j = 0; j < TILES; ++j)
#pragma unroll
            for (int jn = 0; jn < 2; ++jn) {
                const f2_t o = f2_silu(f2_fma(acc_cur[j][jn], s2[jn], h2[jn]));
                sum[jn] = f2_add(sum[jn], o);
                *reinterpret_cast<f2_t*>(yr + (int64_t)(8 * j + g) * C + 8 * jn) = o;
                acc_cur[j][jn] = acc_nxt[j][jn];
                acc_nxt[j][jn] = 0ull;
            }
    };
    // output row oy = expanded rows 2 oy (ky 0), 2 oy + 1 (ky 1), 2 oy + 2 (ky 2; also ky 0 of row oy + 1)
    if (oy0 < oy1) {
        expanded_row(2 * oy0, 0, -1);
        for (int oy = oy0; oy < oy1; ++oy) {
            expanded_row(2 * oy + 1, 1, -1);
            expanded_row(2 * oy + 2, 2, oy + 1 < oy1 ? 0 : -1);
            finish_row(oy);
        }
    }
    if (partial) {          // per (frame, band, channel) sums of the activated outputs: reduce over the 8 pixel lanes g
#pragma unroll
        for (int jn = 0; jn < 2; ++jn) {
            float a, c2;
            f2_unpack(sum[jn], a, c2);
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c2 += __shfl_xor_sync(0xffffffffu, c2, o); }
            if (g == 0) *reinterpret_cast<float2*>(partial + ((int64_t)b * bands + band) * C + n0 + 8 * jn + 2 * t) = make_float2(a, c2);
        }
    }
}

}  // namespace mbs

static int g_mbs = 1;      // dev A/B switch (orbit_set_global_option "mbconv_stream")
void set_mbconv_stream(int on) { g_mbs = on; }
int get_mbconv_stream() { return g_mbs; }

static int mbs_bands(int Ho) { return Ho >= 56 ? 4 : (Ho >= 28 ? 2 : 1); }

bool mbs_supported(int Cin, int C, int H, int W, int k, int stride) {
    return g_mbs && Cin == mbs::kCin && k == 3 && stride == 2 && C % 16 == 0 && C <= 96 && H % 2 == 0 && W % 16 == 0 && W / 16 <= mbs::kTilesMax;
}
int mbs_partial_groups(int Ho) { return mbs_bands(Ho); }

int launch_mbconv_stream(const float* xin, const float* we, const float* scale1, const float* shift1, const float* wt,
                         const float* scale, const float* shift, float* y, float* partial, int B, int H, int W, int Cin, int C,
                         cudaStream_t st) {
    if (!mbs_supported(Cin, C, H, W, 3, 2)) return ORBIT_ERR_UNSUPPORTED;
    if (B <= 0) return ORBIT_OK;
    const int Ho = H / 2, bands = mbs_bands(Ho), rows_per_band = ceil_div(Ho, bands);
    dim3 grid(bands, B), block(32 * (C / 16));
#define ORBIT_MBS(T) if (W == 16 * T) { mbs::mbs_kernel<T><<<grid, block, 0, st>>>(xin, we, scale1, shift1, wt, scale, shift, y, partial, H, C, rows_per_band); \
                                        ORBIT_RETURN_IF_LAUNCH_FAILED(); return ORBIT_OK; }
    ORBIT_MBS(7) ORBIT_MBS(6) ORBIT_MBS(5) ORBIT_MBS(4) ORBIT_MBS(3) ORBIT_MBS(2) ORBIT_MBS(1)
#undef ORBIT_MBS
    return ORBIT_ERR_UNSUPPORTED;
}

}  // namespace orbit
