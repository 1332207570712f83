// Row-streaming pointwise-conv GEMM for the HBM-shaped 1x1 layers with SMALL K and N (the 112x112 / 56x56 / 28x28 stages of
// EfficientNet-B0: K, N <= 240, millions of rows). out[M,N] = act((A[M,K] (*gate)) W[N,K]^T * scale + shift) (+ residual).
//
// Reference op: the timm conv_pw / conv_pwl 1x1 convolutions + BatchNorm (+ FiLM) (+ SiLU) (+ skip) inside the extractor
// invoked at model/few_shot_recognisers.py:114-117,143-146.
//
// Why a second GEMM kernel. The in-kernel role trace of pw_tcgen05_kernel on these layers (round 2)
// shows 1,900-2,200 clocks per 128-row tile where HBM needs 1,060: a tile is ONE or two k-blocks, so the per-tile hand-offs
// TMA -> transform -> tcgen05.mma -> TMEM -> epilogue -> staging -> TMA store (each a few hundred clocks of latency, three
// ring stages deep) are paid per 25-60 KB of traffic, and nothing amortises them. Here there is no pipeline to hand work
// through: every warp streams 16-row groups on its own --
//   A rows      straight from global memory into mma fragments: within a 16-wide k-step lane t owns the logical k slots
//               {2t, 2t+1, 2t+8, 2t+9}; they are MAPPED to the actual columns {4t .. 4t+3}, so a fragment is one 128-bit load
//               per row (the weight fragments use the same permutation, so the products pair up correctly);
//   FP16x3      the same split as the tcgen05 kernel (hi = fp16(x), lo = fp16((x - hi) 2^11)): hi.hi per k-step into a FRESH
//               accumulator that is added to the running sum in fp32 registers with round-to-nearest (tensor-core accumulation
//               truncates), hi.lo + lo.hi accumulate in the tensor core and are scaled by 2^-11 at the end;
//   weights     pre-split fp16 fragments of the whole [N,K] matrix live in shared memory ([k-step][n-tile][lane] x 16 bytes,
//               one conflict-free LDS.128 per (k-step, n-tile));
//   epilogue    in registers: scale / shift (+ SiLU) (+ residual), then 64-bit stores that cover whole 32-byte sectors.
// Latency is hidden by occupancy (2-3 blocks of 8 warps per SM, each warp with 2-9 KB of loads in flight), as in the
// depthwise kernels -- not by an asynchronous ring.
#include <mma.h>

#include "gemm_tcgen05.cuh"

namespace orbit {
namespace st {

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
// (x0, x1) -> packed fp16 hi pair (x0 in the low half) and the scaled residual pair lo = fp16((x - hi) 2^11)
__device__ __forceinline__ void split_f16x2(f2_t x, uint32_t& hi, uint32_t& lo) {
    float x0, x1, h0, h1, r0, r1;
    f2_unpack(x, x0, x1);
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
    const f2_t r = f2_fma(f2_pack(h0, h1), f2_pack(-2048.0f, -2048.0f), f2_mul(x, f2_pack(2048.0f, 2048.0f)));
    f2_unpack(r, r0, r1);
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
// D = A (16x16 fp16, row) * B (16x8 fp16, col) + C, fp32 accumulate
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, const float (&c)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

struct Params {
    const float* A;
    const __half* w_hi;      // [N][Kp] fp16 (launch_weight_split)
    const __half* w_lo;      // [N][Kp] fp16, scaled by 2^11
    const float* scale;
    const float* shift;
    const float* gate;       // [frames][K] or null
    const float* residual;   // [M][N] or null
    float* out;
    int M, Kp;
    uint32_t rpf_mul, rpf_shift;     // row / rows_per_frame = umulhi(row, mul) >> shift (mul == 0: the row itself)
    float debias;            // added to every promoted hi.hi partial, in ulps of that partial (0: none)
};

// K, N: the layer's true sizes (K % 4 == 0, N % 8 == 0). NPASS n-tiles (8 columns each) are accumulated per pass over A.
template <int K, int N, bool GATED, int ACT, bool RES, int NPASS>
__global__ void __launch_bounds__(256, 2)
pw_stream_kernel(const Params p) {
    constexpr int KS = (K + 15) / 16;            // k-steps of 16
    constexpr int NT = N / 8;                    // n-tiles of 8 columns
    constexpr int PASSES = (NT + NPASS - 1) / NPASS;
    extern __shared__ __align__(16) uint8_t s_raw[];
    uint4* s_w = reinterpret_cast<uint4*>(s_raw);                    // [KS][NT][32] {hi b0, hi b1, lo b0, lo b1}
    float2* s_ss = reinterpret_cast<float2*>(s_raw + (size_t)KS * NT * 32 * 16);   // [N/2] scale pair, then [N/2] shift pair
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    // ---- weight fragments: lane (g, t) of n-tile j, k-step s holds W[8j + g][16s + 4t .. 16s + 4t + 3] ----
    for (int i = threadIdx.x; i < KS * NT * 32; i += blockDim.x) {
        const int l = i & 31, j = (i >> 5) % NT, s = i / (32 * NT);
        const int n = 8 * j + (l >> 2), k = 16 * s + 4 * (l & 3);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (k < K) {                                                 // K % 4 == 0: the four columns are valid together
            const uint2 h = *reinterpret_cast<const uint2*>(p.w_hi + (size_t)n * p.Kp + k);
            const uint2 lo = *reinterpret_cast<const uint2*>(p.w_lo + (size_t)n * p.Kp + k);
            v = make_uint4(h.x, h.y, lo.x, lo.y);
        }
        s_w[i] = v;
    }
    for (int i = threadIdx.x; i < N / 2; i += blockDim.x) {
        s_ss[i] = make_float2(p.scale[2 * i], p.scale[2 * i + 1]);
        s_ss[N / 2 + i] = make_float2(p.shift[2 * i], p.shift[2 * i + 1]);
    }
    __syncthreads();

    const int groups = (p.M + 15) >> 4;
    const bool kvalid_last = (16 * (KS - 1) + 4 * t) < K;            // this lane's columns of the last k-step exist
    for (int grp = blockIdx.x * 8 + warp; grp < groups; grp += gridDim.x * 8) {
        const int r_lo = grp * 16 + g, r_hi = r_lo + 8;
        const bool ok_lo = r_lo < p.M, ok_hi = r_hi < p.M;
        const float* a_lo = p.A + (size_t)min(r_lo, p.M - 1) * K + 4 * t;
        const float* a_hi = p.A + (size_t)min(r_hi, p.M - 1) * K + 4 * t;
        // ---- A fragments of the whole row (all k-steps), gated, split into fp16 hi / lo ----
        float4 va[KS], vb[KS];
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const bool kv = s + 1 < KS || kvalid_last;
            va[s] = (ok_lo && kv) ? ldg4_stream(a_lo + 16 * s) : make_float4(0.f, 0.f, 0.f, 0.f);
            vb[s] = (ok_hi && kv) ? ldg4_stream(a_hi + 16 * s) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float* g_lo = nullptr;
        const float* g_hi = nullptr;
        if (GATED) {
            const uint32_t rl = (uint32_t)min(r_lo, p.M - 1), rh = (uint32_t)min(r_hi, p.M - 1);
            const uint32_t f_lo = p.rpf_mul ? (__umulhi(rl, p.rpf_mul) >> p.rpf_shift) : rl;
            const uint32_t f_hi = p.rpf_mul ? (__umulhi(rh, p.rpf_mul) >> p.rpf_shift) : rh;
            g_lo = p.gate + (size_t)f_lo * K + 4 * t;
            g_hi = p.gate + (size_t)f_hi * K + 4 * t;
        }
        // gate (fp32, round to nearest) and split one k-step's fragments:
        // a0 = (row g, slots 2t, 2t+1), a1 = (row g+8, same), a2 = (row g, slots 2t+8, 2t+9), a3 = (row g+8, same)
        auto split_step = [&](int s, uint32_t (&h)[4], uint32_t (&l)[4]) {
            f2_t x0 = f2_pack(va[s].x, va[s].y), x1 = f2_pack(va[s].z, va[s].w);
            f2_t y0 = f2_pack(vb[s].x, vb[s].y), y1 = f2_pack(vb[s].z, vb[s].w);
            if (GATED) {
                const bool kv = s + 1 < KS || kvalid_last;
                const float4 ga = kv ? ldg4(g_lo + 16 * s) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 gb = kv ? ldg4(g_hi + 16 * s) : make_float4(0.f, 0.f, 0.f, 0.f);
                x0 = f2_mul(x0, f2_pack(ga.x, ga.y)); x1 = f2_mul(x1, f2_pack(ga.z, ga.w));
                y0 = f2_mul(y0, f2_pack(gb.x, gb.y)); y1 = f2_mul(y1, f2_pack(gb.z, gb.w));
            }
            split_f16x2(x0, h[0], l[0]); split_f16x2(y0, h[1], l[1]);
            split_f16x2(x1, h[2], l[2]); split_f16x2(y1, h[3], l[3]);
        };
        // several passes over the columns re-use the split fragments (small K); a single pass splits each k-step right
        // before its MMAs, so that only the raw fp32 row stays live (large K: registers)
        constexpr int HS = PASSES > 1 ? KS : 1;
        uint32_t ah[HS][4], al[HS][4];
        if (PASSES > 1) {
#pragma unroll
            for (int s = 0; s < KS; ++s) split_step(s, ah[s], al[s]);
        }
        float* o_lo = p.out + (size_t)min(r_lo, p.M - 1) * N + 2 * t;
        float* o_hi = p.out + (size_t)min(r_hi, p.M - 1) * N + 2 * t;
        const float* q_lo = RES ? p.residual + (size_t)min(r_lo, p.M - 1) * N + 2 * t : nullptr;
        const float* q_hi = RES ? p.residual + (size_t)min(r_hi, p.M - 1) * N + 2 * t : nullptr;
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
            const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
            float acc[NPASS][4], cor[NPASS][4];
            f2_t res_lo[NPASS], res_hi[NPASS];
#pragma unroll
            for (int jj = 0; jj < NPASS; ++jj) {
                const int j = ps * NPASS + jj;
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc[jj][i] = 0.f; cor[jj][i] = 0.f; }
                if (RES && j < NT) {            // issued early: consumed after the MMAs
                    res_lo[jj] = ok_lo ? __ldg(reinterpret_cast<const f2_t*>(q_lo + 8 * j)) : 0ull;
                    res_hi[jj] = ok_hi ? __ldg(reinterpret_cast<const f2_t*>(q_hi + 8 * j)) : 0ull;
                }
            }
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                constexpr bool once = PASSES == 1;
                if (once) split_step(s, ah[0], al[0]);
                const uint32_t (&fh)[4] = ah[once ? 0 : s];
                const uint32_t (&fl)[4] = al[once ? 0 : s];
#pragma unroll
                for (int jj = 0; jj < NPASS; ++jj) {
                    const int j = ps * NPASS + jj;
                    if (j < NT) {
                        const uint4 w = s_w[(s * NT + j) * 32 + lane];
                        float m4[4];
                        mma_f16(m4, fh, w.x, w.y, zero4);                    // hi.hi of ONE k-step
                        mma_f16(cor[jj], fl, w.x, w.y, cor[jj]);              // lo.hi
                        mma_f16(cor[jj], fh, w.z, w.w, cor[jj]);              // hi.lo
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[jj][i] += m4[i];      // round-to-nearest promotion
                    }
                }
            }
#pragma unroll
            for (int jj = 0; jj < NPASS; ++jj) {
                const int j = ps * NPASS + jj;
                if (j < NT) {
                    const f2_t inv = f2_pack(4.8828125e-4f, 4.8828125e-4f);   // 2^-11
                    f2_t v_lo = f2_fma(f2_pack(cor[jj][0], cor[jj][1]), inv, f2_pack(acc[jj][0], acc[jj][1]));
                    f2_t v_hi = f2_fma(f2_pack(cor[jj][2], cor[jj][3]), inv, f2_pack(acc[jj][2], acc[jj][3]));
                    const float2 sc = s_ss[4 * j + t], sh = s_ss[N / 2 + 4 * j + t];
                    const f2_t sc2 = f2_pack(sc.x, sc.y), sh2 = f2_pack(sh.x, sh.y);
                    v_lo = f2_fma(v_lo, sc2, sh2);
                    v_hi = f2_fma(v_hi, sc2, sh2);
                    if (ACT == 1) { v_lo = f2_silu(v_lo); v_hi = f2_silu(v_hi); }
                    if (RES) { v_lo = f2_add(v_lo, res_lo[jj]); v_hi = f2_add(v_hi, res_hi[jj]); }
                    if (ok_lo) *reinterpret_cast<f2_t*>(o_lo + 8 * j) = v_lo;
                    if (ok_hi) *reinterpret_cast<f2_t*>(o_hi + 8 * j) = v_hi;
                }
            }
        }
    }
}

template <int K, int N, bool GATED, int ACT, bool RES, int NPASS>
int launch_instance(const Params& p, cudaStream_t stream) {
    constexpr int KS = (K + 15) / 16, NT = N / 8;
    const size_t smem = (size_t)KS * NT * 32 * 16 + (size_t)N * 2 * sizeof(float);
    auto fn = pw_stream_kernel<K, N, GATED, ACT, RES, NPASS>;
    static_assert((size_t)KS * NT * 32 * 16 + (size_t)N * 2 * sizeof(float) <= 48 * 1024, "weight fragments must fit the default 48 KB");
    static int blocks_per_sm = 0;        // per instantiation (identical on every B200 of the node)
    if (blocks_per_sm == 0) {
        int b = 0;
        ORBIT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, fn, 256, smem));
        if (b < 1) return ORBIT_ERR_UNSUPPORTED;
        blocks_per_sm = b;
    }
    const int groups = (p.M + 15) / 16;
    const int grid = std::max(1, std::min(148 * blocks_per_sm, (groups + 7) / 8));
    fn<<<grid, 256, smem, stream>>>(p);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

}  // namespace st

static int g_stream_gemm = 1;      // dev A/B switch (orbit_set_global_option "tc_stream")
void set_stream_gemm(int on) { g_stream_gemm = on; }
int get_stream_gemm() { return g_stream_gemm; }

// Returns ORBIT_ERR_UNSUPPORTED when no instance covers the shape: the caller then runs the tcgen05 kernel.
int launch_pointwise_stream(const float* A, const float* w_split, const float* scale, const float* shift, const float* gate,
                            const float* residual, float* out, int M, int N, int K, int rows_per_frame, int act, cudaStream_t stream) {
    if (!g_stream_gemm) return ORBIT_ERR_UNSUPPORTED;
    if (M <= 0) return ORBIT_OK;
    st::Params p;
    p.A = A; p.scale = scale; p.shift = shift; p.gate = gate; p.residual = residual; p.out = out;
    p.M = M; p.Kp = (K + 7) / 8 * 8;
    p.w_hi = reinterpret_cast<const __half*>(w_split);
    p.w_lo = p.w_hi + (size_t)N * p.Kp;
    p.debias = 0.f;
    {
        const uint32_t d = (uint32_t)std::max(rows_per_frame, 1);
        uint32_t l = 0;
        while ((1ull << l) < d) ++l;
        p.rpf_mul = d == 1 ? 0u : (uint32_t)(((1ull << (31 + l)) / d) + 1);
        p.rpf_shift = d == 1 ? 0u : (31 + l - 32);
    }
    const bool g = gate != nullptr, r = residual != nullptr;
#define ORBIT_ST(KK, NN, GG, AA, RR, NP) \
    if (K == KK && N == NN && g == GG && act == AA && r == RR) return st::launch_instance<KK, NN, GG, AA, RR, NP>(p, stream);
    // EfficientNet-B0 layers at 112x112 / 56x56 / 28x28 (K, N, gated, act, residual, n-tiles per pass). Measured on B200
    // (us per 1,024 frames, tcgen05 -> streaming): 613 -> 415, 407 -> 259, 855 -> 414, 1,461 -> 1,072, 255 -> 225. The K = 144
    // projections (550 -> 568, 171 -> 177: the per-row gate loads and 81 MMAs per 16 rows) stay with the tcgen05 kernel.
    ORBIT_ST(32, 16, true, 0, false, 2)        // block 0 project
    ORBIT_ST(96, 24, true, 0, false, 3)        // block 1.0 project
    ORBIT_ST(24, 144, false, 1, false, 6)      // block 1.1 expand
    ORBIT_ST(16, 96, false, 1, false, 6)       // block 1.0 expand (when not fused)
    ORBIT_ST(40, 240, false, 1, false, 6)      // block 2.1 / 3.0 expand
#undef ORBIT_ST
    return ORBIT_ERR_UNSUPPORTED;
}

}  // namespace orbit
