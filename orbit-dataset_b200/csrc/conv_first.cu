// First convolution of the ResNet-18 extension (7x7 stride 2, 3 -> 64) and of the set encoder (3x3 stride 1, 3 -> 64) as a
// DIRECT tensor-core convolution on the NCHW fp32 frames: no im2col matrix.
//
// Reference op sites: SimplePrePoolNet.layer1 (model/set_encoders.py:91-105: Conv2d(3, 64, 3, padding=1) + BatchNorm2d + ReLU)
// over every support frame (few_shot_recognisers.py:361-386); torchvision resnet18.conv1 + bn1 + relu for the resnet18
// extension of BASELINE.json configs 1 and 3.
//
// Why: with 3 input channels the explicit im2col matrix is 9 x (3x3) to 12 x (7x7 stride 2) the size of the frame, written and
// read back through HBM: torch.profiler on an S3 episode showed `im2col_nchw_rows_kernel` at 4.5 of 17.3 ms (and it would be
// ~20 ms of an S2-scale CNAPs episode, whose set encoder sees 1,600 support frames).
// How: the row-streaming GEMM scheme (gemm_stream.cu) with the A fragment GATHERED from the frame. K is laid out as
// (channel, ky) groups of GP kx slots (GP = 4 for 3x3, 8 for 7x7; the spare slot has a zero weight): the four k slots of lane t
// in a 16-wide k-step are then four CONSECUTIVE input columns of one (channel, row) -- L1-resident scalar loads with one row
// test per group. A tile = 16 consecutive output pixels of an output row; FP16x3 with per-k-step round-to-nearest promotion;
// folded BatchNorm scale / shift (+ ReLU) in registers; NHWC fp32 output.
#include "convnet.cuh"

namespace orbit {
namespace cf {

typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
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

constexpr int kCout = 64, kNT = kCout / 8;

// KS x KS convolution, stride S, symmetric padding PAD, 3 input channels (NCHW), 64 output channels (NHWC).
template <int KS, int S, int PAD>
__global__ void __launch_bounds__(256, 2)
conv_first_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                  const float* __restrict__ shift, float* __restrict__ y, int B, int H, int W, int Ho, int Wo, int act) {
    constexpr int GP = KS <= 4 ? 4 : 8;                 // kx slots per (channel, ky) group
    constexpr int NG = 3 * KS;                          // groups
    constexpr int KSTEPS = (NG * GP + 15) / 16;
    constexpr int GPS = 16 / GP;                        // groups per k-step: 4 or 2
    extern __shared__ __align__(16) uint8_t s_raw[];
    uint4* s_w = reinterpret_cast<uint4*>(s_raw);                                   // [KSTEPS][kNT][32] {hi b0, hi b1, lo b0, lo b1}
    float2* s_ss = reinterpret_cast<float2*>(s_raw + (size_t)KSTEPS * kNT * 32 * 16);   // scale pairs [32], shift pairs [32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    // lane t's four k slots of a k-step: group (per k-step) and first kx
    const int grp_in_step = GP == 4 ? t : (t >> 1), kx0 = GP == 4 ? 0 : 4 * (t & 1);
    // ---- weight fragments: n = 8 jn + (l >> 2), the four k slots of lane l & 3 ----
    for (int i = threadIdx.x; i < KSTEPS * kNT * 32; i += blockDim.x) {
        const int l = i & 31, jn = (i >> 5) % kNT, s = i / (32 * kNT);
        const int n = 8 * jn + (l >> 2), tt = l & 3;
        const int grp = s * GPS + (GP == 4 ? tt : (tt >> 1)), k0 = GP == 4 ? 0 : 4 * (tt & 1);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (grp < NG) {
            const int c = grp / KS, ky = grp % KS;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (k0 + e < KS) v[e] = w[(((int64_t)n * 3 + c) * KS + ky) * KS + k0 + e];
        }
        uint4 f;
        split_f16x2(f2_pack(v[0], v[1]), f.x, f.z);
        split_f16x2(f2_pack(v[2], v[3]), f.y, f.w);
        s_w[i] = f;
    }
    for (int i = threadIdx.x; i < kCout / 2; i += blockDim.x) {
        s_ss[i] = make_float2(scale[2 * i], scale[2 * i + 1]);
        s_ss[kCout / 2 + i] = make_float2(shift[2 * i], shift[2 * i + 1]);
    }
    __syncthreads();

    const int tiles_per_row = (Wo + 15) >> 4;
    const int64_t tiles = (int64_t)B * Ho * tiles_per_row;
    const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t tile = blockIdx.x * 8 + warp; tile < tiles; tile += (int64_t)gridDim.x * 8) {
        const int j = (int)(tile % tiles_per_row), oy = (int)((tile / tiles_per_row) % Ho), b = (int)(tile / ((int64_t)tiles_per_row * Ho));
        const int ox_lo = 16 * j + g, ox_hi = ox_lo + 8;                    // mma rows g and g + 8
        const float* xb = x + (int64_t)b * 3 * H * W;
        float acc[kNT][4], cor[kNT][4];
#pragma unroll
        for (int jn = 0; jn < kNT; ++jn)
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[jn][i] = 0.f; cor[jn][i] = 0.f; }
#pragma unroll
        for (int s = 0; s < KSTEPS; ++s) {
            const int grp = s * GPS + grp_in_step;
            float va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
            if (grp < NG) {
                const int c = grp / KS, ky = grp - c * KS;
                const int iy = oy * S - PAD + ky;
                if (iy >= 0 && iy < H) {
                    const float* rp = xb + ((int64_t)c * H + iy) * W;
                    const int cl = ox_lo * S - PAD + kx0, ch = ox_hi * S - PAD + kx0;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (kx0 + e < KS) {                                 // (spare slots: zero weight, leave the zero)
                            if ((unsigned)(cl + e) < (unsigned)W) va[e] = __ldg(rp + cl + e);
                            if ((unsigned)(ch + e) < (unsigned)W) vb[e] = __ldg(rp + ch + e);
                        }
                    }
                }
            }
            uint32_t ah[4], al[4];
            split_f16x2(f2_pack(va[0], va[1]), ah[0], al[0]); split_f16x2(f2_pack(vb[0], vb[1]), ah[1], al[1]);
            split_f16x2(f2_pack(va[2], va[3]), ah[2], al[2]); split_f16x2(f2_pack(vb[2], vb[3]), ah[3], al[3]);
#pragma unroll
            for (int jn = 0; jn < kNT; ++jn) {
                const uint4 wf = s_w[(s * kNT + jn) * 32 + lane];
                float m4[4];
                mma_f16(m4, ah, wf.x, wf.y, zero4);                          // hi.hi of ONE k-step
                mma_f16(cor[jn], al, wf.x, wf.y, cor[jn]);                   // lo.hi
                mma_f16(cor[jn], ah, wf.z, wf.w, cor[jn]);                   // hi.lo
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[jn][i] += m4[i];             // round-to-nearest promotion
            }
        }
        const bool ok_lo = ox_lo < Wo, ok_hi = ox_hi < Wo;
        float* o_lo = y + (((int64_t)b * Ho + oy) * Wo + min(ox_lo, Wo - 1)) * kCout + 2 * t;
        float* o_hi = y + (((int64_t)b * Ho + oy) * Wo + min(ox_hi, Wo - 1)) * kCout + 2 * t;
        const f2_t inv = f2_pack(4.8828125e-4f, 4.8828125e-4f);               // 2^-11
#pragma unroll
        for (int jn = 0; jn < kNT; ++jn) {
            const float2 sc = s_ss[4 * jn + t], sh = s_ss[kCout / 2 + 4 * jn + t];
            const f2_t sc2 = f2_pack(sc.x, sc.y), sh2 = f2_pack(sh.x, sh.y);
            f2_t v_lo = f2_fma(f2_fma(f2_pack(cor[jn][0], cor[jn][1]), inv, f2_pack(acc[jn][0], acc[jn][1])), sc2, sh2);
            f2_t v_hi = f2_fma(f2_fma(f2_pack(cor[jn][2], cor[jn][3]), inv, f2_pack(acc[jn][2], acc[jn][3])), sc2, sh2);
            if (act == ACT_RELU) {
                float a0, a1, b0, b1;
                f2_unpack(v_lo, a0, a1); f2_unpack(v_hi, b0, b1);
                v_lo = f2_pack(fmaxf(a0, 0.f), fmaxf(a1, 0.f)); v_hi = f2_pack(fmaxf(b0, 0.f), fmaxf(b1, 0.f));
            }
            if (ok_lo) *reinterpret_cast<f2_t*>(o_lo + 8 * jn) = v_lo;
            if (ok_hi) *reinterpret_cast<f2_t*>(o_hi + 8 * jn) = v_hi;
        }
    }
}

template <int KS, int S, int PAD>
int launch_instance(const float* x, const float* w, const float* scale, const float* shift, float* y, int B, int H, int W, int Ho, int Wo,
                    int act, cudaStream_t st) {
    constexpr int GP = KS <= 4 ? 4 : 8, KSTEPS = (3 * KS * GP + 15) / 16;
    const size_t smem = (size_t)KSTEPS * kNT * 32 * 16 + (size_t)kCout * 2 * sizeof(float);
    static_assert((size_t)KSTEPS * kNT * 32 * 16 + kCout * 8 <= 48 * 1024, "weight fragments must fit the default 48 KB");
    const int64_t tiles = (int64_t)B * Ho * ((Wo + 15) / 16);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(148 * 2, (tiles + 7) / 8));
    conv_first_kernel<KS, S, PAD><<<grid, 256, smem, st>>>(x, w, scale, shift, y, B, H, W, Ho, Wo, act);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

}  // namespace cf

static int g_conv_first = 1;      // dev A/B switch (orbit_set_global_option "conv_first")
void set_conv_first(int on) { g_conv_first = on; }
int get_conv_first() { return g_conv_first; }

// x [B,3,H,W] NCHW -> y [B,Ho,Wo,64] NHWC. ORBIT_ERR_UNSUPPORTED when no instance covers the geometry (callers fall back to im2col).
int launch_conv_first(const float* x, const float* w, const float* scale, const float* shift, float* y, int B, int H, int W, int Cin,
                      int Cout, int k, int stride, int pad, int Ho, int Wo, int act, cudaStream_t st) {
    if (!g_conv_first || Cin != 3 || Cout != cf::kCout || (act != ACT_RELU && act != ACT_NONE)) return ORBIT_ERR_UNSUPPORTED;
    if (B <= 0) return ORBIT_OK;
    if (k == 3 && stride == 1 && pad == 1) return cf::launch_instance<3, 1, 1>(x, w, scale, shift, y, B, H, W, Ho, Wo, act, st);
    if (k == 7 && stride == 2 && pad == 3) return cf::launch_instance<7, 2, 3>(x, w, scale, shift, y, B, H, W, Ho, Wo, act, st);
    return ORBIT_ERR_UNSUPPORTED;
}

}  // namespace orbit
