// CUDA-core kernels of the convolutional backbones (EfficientNet-B0 family; fp32, NHWC).
//
// Reference op sites (microsoft/ORBIT-Dataset @ 97ccae1): the timm model invoked at
// model/few_shot_recognisers.py:114-117,143-146 (timm==0.6.12 tf_efficientnet_b0, external), with
// FiLM = substituted BatchNorm affine parameters (model/film.py:38-74).  Per layer:
//   conv_stem+bn1+SiLU        -> stem_kernel            (direct conv, NCHW in / NHWC out)
//   conv_dw+bn(+FiLM)+SiLU    -> dw2_kernel             (HBM-bound stencil, fused SE squeeze partials)
//   SqueezeExcite FCs         -> se_gate_kernel
//   conv_pw/conv_pwl/conv_head-> pw_ffma_kernel (this file, fp32 FFMA tiles) or the tcgen05 kernel
//                                in gemm_tcgen05.cu; BN/FiLM scale-shift, SiLU, SE gate and residual fused
//   global_pool               -> spatial_mean_kernel
#include <type_traits>

#include "convnet.cuh"

namespace orbit {

// ------------------------------------------------------------------------------------------------
// BatchNorm fold: scale = gamma * rsqrt(var + eps), shift = beta - mean * scale, with FiLM gamma'/beta'
// substituted where the layer is a FiLM site.
// ------------------------------------------------------------------------------------------------
constexpr int kFoldBatch = 32;
struct FoldTable {
    FoldEntry e[kFoldBatch];
};

__global__ void __launch_bounds__(256)
bn_fold_kernel(FoldTable tab, const float* __restrict__ params, const float* __restrict__ film, float* __restrict__ derived) {
    const FoldEntry& e = tab.e[blockIdx.y];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < e.channels; c += gridDim.x * blockDim.x) {
        const float g = (film && e.film_gamma >= 0) ? film[e.film_gamma + c] : params[e.gamma + c];
        const float b = (film && e.film_beta >= 0) ? film[e.film_beta + c] : params[e.beta + c];
        const float inv = 1.0f / sqrtf(params[e.var + c] + e.eps);
        const float sc = g * inv;
        const float cb = e.conv_bias >= 0 ? params[e.conv_bias + c] : 0.f;
        derived[e.out + c] = sc;
        derived[e.out + e.channels + c] = b + (cb - params[e.mean + c]) * sc;
    }
}

int launch_bn_fold(const FoldEntry* entries, int n, const float* params, const float* film, float* derived, cudaStream_t st) {
    for (int s = 0; s < n; s += kFoldBatch) {
        FoldTable tab;
        const int m = std::min(kFoldBatch, n - s);
        int maxc = 1;
        for (int i = 0; i < m; ++i) { tab.e[i] = entries[s + i]; maxc = std::max(maxc, entries[s + i].channels); }
        dim3 grid(ceil_div(maxc, 256), m);
        bn_fold_kernel<<<grid, 256, 0, st>>>(tab, params, film, derived);
        ORBIT_RETURN_IF_LAUNCH_FAILED();
    }
    return ORBIT_OK;
}

__global__ void add_vec_kernel(float* __restrict__ a, const float* __restrict__ b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}
int launch_add_vec(float* a, const float* b, int n, cudaStream_t st) {
    add_vec_kernel<<<ceil_div(n, 256), 256, 0, st>>>(a, b, n);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

__global__ void fill_identity_kernel(float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[i] = 1.0f; out[n + i] = 0.0f; }
}
int launch_fill_identity(float* out, int n, cudaStream_t st) {
    fill_identity_kernel<<<ceil_div(n, 256), 256, 0, st>>>(out, n);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// One block per 32 channels; 8 row lanes x 32 channels; two passes (mean, then centred second moment).
__global__ void __launch_bounds__(256)
channel_stats_kernel(const float* __restrict__ x, int64_t M, int C, float* __restrict__ mean, float* __restrict__ var) {
    __shared__ double s_acc[8][33];
    __shared__ double s_mean[32];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const bool live = c < C;
    double acc = 0.0;
    if (live) for (int64_t r = rl; r < M; r += 8) acc += (double)x[r * C + c];
    s_acc[rl][cl] = acc;
    __syncthreads();
    if (rl == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += s_acc[i][cl];
        s_mean[cl] = t / (double)M;
    }
    __syncthreads();
    const double mu = s_mean[cl];
    acc = 0.0;
    if (live) for (int64_t r = rl; r < M; r += 8) { const double d = (double)x[r * C + c] - mu; acc += d * d; }
    __syncthreads();
    s_acc[rl][cl] = acc;
    __syncthreads();
    if (rl == 0 && live) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += s_acc[i][cl];
        mean[c] = (float)mu;
        var[c] = (float)(t / (double)(M > 1 ? M - 1 : 1));
    }
}
int launch_channel_stats(const float* x, int64_t M, int C, float* mean, float* var, cudaStream_t st) {
    channel_stats_kernel<<<ceil_div(C, 32), 256, 0, st>>>(x, M, C, mean, var);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

__global__ void dw_relayout_kernel(const float* __restrict__ w, int C, int kk, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C * kk) { const int t = i / C, c = i % C; out[i] = w[c * kk + t]; }
}
int launch_dw_relayout(const float* w, int C, int kk, float* out, cudaStream_t st) {
    dw_relayout_kernel<<<ceil_div(C * kk, 256), 256, 0, st>>>(w, C, kk, out);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_SILU) return siluf_(v);
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));   // exact (erf) GELU, as timm's nn.GELU
    return v;
}

// x * sigmoid(x) with the SFU approximations (ex2.approx, rcp.approx; ~3e-7 relative error). ncu showed the
// depthwise kernels to be ISSUE-bound, not memory-bound, with the accurate expf + IEEE division taking ~30 of
// the ~85 instructions per output.
__device__ __forceinline__ float act_fast(float v, int act) {
    if (act == ACT_SILU) return silu_sfu(v);
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    return v;
}


// ------------------------------------------------------------------------------------------------
// Stem: 3x3 stride-2 conv on the fp32 NCHW frames, one output pixel (all COUT channels) per thread.
// The NCHW->NHWC change of layout is fused here so the frames are read exactly once.
// ------------------------------------------------------------------------------------------------
// Two horizontally adjacent output pixels per thread (they share one of their three input columns and every
// weight fetch), channel pairs on the packed fma.rn.f32x2 pipe: 27 x (8 LDS.128 + 32 FFMA2) per pixel pair instead
// of 2 x 27 x (8 LDS.128 + 32 FFMA) in round 1.
typedef unsigned long long stem_f2_t;
__device__ __forceinline__ stem_f2_t stem_dup(float v) { stem_f2_t r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v)); return r; }
__device__ __forceinline__ stem_f2_t stem_fma2(stem_f2_t a, stem_f2_t b, stem_f2_t c) {
    stem_f2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int COUT>
__global__ void __launch_bounds__(128, 4)
stem_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
            const float* __restrict__ shift, float* __restrict__ y, int B, int H, int W, int Ho, int Wo, int pad_t,
            int pad_l, int act) {
    constexpr int NP = COUT / 2;
    __shared__ __align__(16) float s_w[27 * COUT];  // [tap][co]
    __shared__ __align__(16) float s_sc[COUT], s_sh[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) { const int tap = i / COUT, co = i % COUT; s_w[i] = w[co * 27 + tap]; }
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) { s_sc[i] = scale[i]; s_sh[i] = shift[i]; }
    __syncthreads();
    const int Wp = (Wo + 1) / 2;                      // pixel pairs per output row
    const int64_t total_pairs = (int64_t)B * Ho * Wp;
    // a block walks over several 128-pair groups (grid-stride): the weight staging above is paid once per block, not once per 256 pixels
    for (int64_t grp = blockIdx.x; grp * blockDim.x < total_pairs; grp += gridDim.x) {
    const int64_t idx = grp * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= total_pairs) break;                    // (whole warps leave together except in the last group: no block barrier below)
    const int ox = 2 * (int)(idx % Wp), oy = (int)((idx / Wp) % Ho), b = (int)(idx / ((int64_t)Wp * Ho));
    const bool second = ox + 1 < Wo;
    stem_f2_t acc[2][NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) { acc[0][i] = 0ull; acc[1][i] = 0ull; }
    const int ix0 = ox * 2 - pad_l;                   // input columns ix0 .. ix0+4 feed the two outputs
    // all 45 input values first (one exposed HBM latency instead of nine), then the 27 x 32 packed FMAs
    float vin[3][3][5];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const float* xp = x + ((int64_t)b * 3 + ci) * H * W;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * 2 - pad_t + ky;
            const bool rok = iy >= 0 && iy < H;
            const float* rp = xp + (int64_t)iy * W + ix0;
#pragma unroll
            for (int j = 0; j < 5; ++j) vin[ci][ky][j] = (rok && ix0 + j >= 0 && ix0 + j < W) ? __ldg(rp + j) : 0.f;
        }
    }
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const stem_f2_t v0 = stem_dup(vin[ci][ky][kx]), v1 = stem_dup(vin[ci][ky][kx + 2]);
                const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(s_w + (ci * 9 + ky * 3 + kx) * COUT);
#pragma unroll
                for (int q = 0; q < COUT / 4; ++q) {
                    const ulonglong2 w2 = wp[q];
                    acc[0][2 * q] = stem_fma2(v0, w2.x, acc[0][2 * q]); acc[0][2 * q + 1] = stem_fma2(v0, w2.y, acc[0][2 * q + 1]);
                    acc[1][2 * q] = stem_fma2(v1, w2.x, acc[1][2 * q]); acc[1][2 * q + 1] = stem_fma2(v1, w2.y, acc[1][2 * q + 1]);
                }
            }
        }
    }
    // Epilogue. A warp's 32 pixel pairs are 64 consecutive NHWC pixels = 8 KB of contiguous output when the row width is
    // even (always for the 224 / 84 / 96 px pyramids): stage them in shared memory (XOR-swizzled, conflict-free both ways)
    // and write 512 contiguous bytes per store instruction instead of 32 scattered 16-byte pieces.
    __shared__ __align__(16) float4 s_out[4][64 * 8];
    const int64_t pix = ((int64_t)b * Ho + oy) * Wo + ox;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool staged = (Wo & 1) == 0 && (grp + 1) * (int64_t)blockDim.x <= total_pairs;   // full group, no odd tail
#pragma unroll
    for (int px = 0; px < 2; ++px) {
        if (px == 1 && !second) break;
        float4* yp = reinterpret_cast<float4*>(y + (pix + px) * COUT);
        const int p = 2 * lane + px;
#pragma unroll
        for (int q = 0; q < COUT / 4; ++q) {
            float a0, a1, a2, a3;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc[px][2 * q]));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(acc[px][2 * q + 1]));
            float4 o;
            o.x = act_fast(fmaf(a0, s_sc[4 * q + 0], s_sh[4 * q + 0]), act);
            o.y = act_fast(fmaf(a1, s_sc[4 * q + 1], s_sh[4 * q + 1]), act);
            o.z = act_fast(fmaf(a2, s_sc[4 * q + 2], s_sh[4 * q + 2]), act);
            o.w = act_fast(fmaf(a3, s_sc[4 * q + 3], s_sh[4 * q + 3]), act);
            if (staged) s_out[warp][p * 8 + (q ^ ((p >> 1) & 7))] = o;
            else yp[q] = o;
        }
    }
    if (staged) {
        __syncwarp();
        float4* yw = reinterpret_cast<float4*>(y + (pix - 2 * lane) * COUT);   // the warp's first pixel (lane 0's pix)
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const int i = it * 32 + lane, p = i >> 3, q = i & 7;
            yw[i] = s_out[warp][p * 8 + (q ^ ((p >> 1) & 7))];
        }
        __syncwarp();                                 // the warp's staging tile is free for the next group
    }
    }
}

static int g_stem_groups = 8;      // dev A/B switch (orbit_set_global_option "stem_groups"): 128-pair groups per stem block
void set_stem_groups(int n) { g_stem_groups = n < 1 ? 1 : n; }
int get_stem_groups() { return g_stem_groups; }

int launch_stem(const float* x, const float* w, const float* scale, const float* shift, float* y, int B, int H, int W,
                int Ho, int Wo, int pad_t, int pad_l, int cout, int act, cudaStream_t st) {
    if (cout != 32) return ORBIT_ERR_UNSUPPORTED;
    const int64_t total = (int64_t)B * Ho * ((Wo + 1) / 2);
    ORBIT_CUDA(cudaFuncSetAttribute(stem_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));   // 4 blocks x 36 KB per SM
    const int64_t groups = ceil_div64(total, 128);
    const unsigned grid = (unsigned)std::min<int64_t>(groups, std::max<int64_t>(148 * 4, ceil_div64(groups, g_stem_groups)));   // ~g_stem_groups groups per block
    stem_kernel<32><<<grid, 128, 0, st>>>(x, w, scale, shift, y, B, H, W, Ho, Wo, pad_t, pad_l, act);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// ------------------------------------------------------------------------------------------------
// Depthwise KxK conv + folded BN/FiLM + activation, NHWC, HBM-bound.
// A thread owns 4 channels (one 128-bit lane) and a strip of TW=4 output columns, and WALKS DOWN the rows of
// its tile with rolling accumulators: every input row is loaded once (SPAN 128-bit loads), and scattered into
// the R = ceil(K/S) output rows it contributes to, which live in registers; an output row is written when its
// last input row has been consumed. Loads per output drop from K*SPAN/TW (row-at-a-time) to SPAN/TW.
// Everything about the ring (which slot a (row phase, ky) pair hits, which slot completes) is resolved at
// compile time by unrolling over the P = R*S row phases.
// grid (tiles * strip_blocks, channel chunks, frames); block = LX lanes x LY strips.
// Epilogue: the activated outputs are summed per (frame, block, channel) for the SE squeeze with a
// fixed-order shared-memory reduction (deterministic; no atomics).
// ------------------------------------------------------------------------------------------------
constexpr int kDwTW = 4;

__host__ __device__ constexpr int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
__host__ __device__ constexpr int floor_mod(int a, int b) { return a - floor_div(a, b) * b; }

struct DwPlan {
    int VEC, LX, LY, nchunks, tiles, rows_per_tile, strip_blocks, groups;
};

// K=3: 4 channels per thread (128-bit accesses); K=5: 2 channels per thread (64-bit) so that the 25 taps,
// the rolling accumulators and the streamed row all stay in registers (~110) without spilling.
static DwPlan dw_plan(int C, int Ho, int Wo, int k, int stride) {
    DwPlan p;
    p.VEC = k == 3 ? 4 : 2;
    const int Cv = C / p.VEC;
    p.nchunks = ceil_div(Cv, 32);
    p.LX = ceil_div(Cv, p.nchunks);
    const int strips = ceil_div(Wo, kDwTW);
    p.LY = std::max(1, std::min(256 / p.LX, strips));
    p.strip_blocks = ceil_div(strips, p.LY);
    const int R = ceil_div(k, stride);
    const int want_tiles = std::min(ceil_div(Ho, R), Ho >= 112 ? 4 : (Ho >= 56 ? 2 : 1));   // few, tall tiles: every tile re-reads K-1 halo rows
    p.rows_per_tile = ceil_div(ceil_div(Ho, want_tiles), R) * R;   // multiple of R: tiles start on a ring boundary
    p.tiles = ceil_div(Ho, p.rows_per_tile);
    p.groups = p.tiles * p.strip_blocks;
    return p;
}

int dw_partial_groups(int C, int Ho, int Wo, int k, int stride) { return dw_plan(C, Ho, Wo, k, stride).groups; }

template <int VEC> struct VecIO;
template <> struct VecIO<4> {
    static __device__ __forceinline__ void load(const float* p, float* v) { const float4 t = ldg4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void store(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct VecIO<2> {
    static __device__ __forceinline__ void load(const float* p, float* v) { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void store(float* p, const float* v) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};

// ------------------------------------------------------------------------------------------------
// The depthwise kernel. ncu of its first, scalar version on the K=5 layers: 98 thread-instructions per output of which only 25 were FFMAs -- 64-bit address
// arithmetic (25 %), ring-rotation MOVs (13 %) and divergence bookkeeping made it ISSUE-bound at 1.3-1.7 TB/s.
// Here: (i) channel PAIRS are processed with the Blackwell packed `fma.rn.f32x2` (two IEEE fp32 FMAs per issue
// slot, bit-identical to two fmaf), (ii) column validity and column offsets are hoisted out of the row loop,
// (iii) the k*k taps live in shared memory ([tap][lane], one conflict-free LDS per tap per row) instead of 50-100
// registers, which pays for (iv) a register prefetch of the NEXT input row while the current one is consumed.
// ------------------------------------------------------------------------------------------------
typedef unsigned long long f2_t;   // two packed fp32
__device__ __forceinline__ void f2_unpack(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t f2_pack(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
    f2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// x * sigmoid(x) on a channel pair: the four MUFU ops stay scalar, the three fp32 ops around them are packed
__device__ __forceinline__ f2_t f2_silu_pair(f2_t x) {
    f2_t t, r;
    float t0, t1, e0, e1, r0, r1;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(x), "l"(f2_pack(-1.4426950408889634f, -1.4426950408889634f)));
    f2_unpack(t, t0, t1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(f2_pack(e0, e1)), "l"(f2_pack(1.0f, 1.0f)));
    f2_unpack(t, t0, t1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(t0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(t1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(f2_pack(r0, r1)));
    return r;
}
template <int NP> struct PairIO;
template <> struct PairIO<1> {
    static __device__ __forceinline__ void load(const float* p, f2_t* v) { v[0] = __ldg(reinterpret_cast<const f2_t*>(p)); }
    static __device__ __forceinline__ void load_shared(const float* p, f2_t* v) { v[0] = *reinterpret_cast<const f2_t*>(p); }
    static __device__ __forceinline__ void store(float* p, const float* r) { *reinterpret_cast<float2*>(p) = make_float2(r[0], r[1]); }
};
template <> struct PairIO<2> {
    static __device__ __forceinline__ void load(const float* p, f2_t* v) {
        const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(p)); v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void load_shared(const float* p, f2_t* v) {
        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(p); v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void store(float* p, const float* r) { *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]); }
};

// CT = compile-time channel count (0: use the runtime C): column offsets j*C become load/store immediates.
// (Tried and removed, see DESIGN.md section 4: a row loop unrolled over the R phases of the accumulator ring -- 5x the code,
// slower on instruction fetch -- and a per-thread cp.async input ring for K=5 stride 1 -- 4-10 % slower.)
template <int K, int S, int VEC, int MINB, int CT>
__global__ void __launch_bounds__(256, MINB)
dw2_kernel(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ scale,
           const float* __restrict__ shift, float* __restrict__ y, float* __restrict__ partial, int H, int W, int C_rt,
           int Ho, int Wo, int pad_t, int pad_l, int act, int LX, int LY, int rows_per_tile, int strip_blocks) {
    constexpr int TW = kDwTW, R = (K + S - 1) / S, SPAN = (TW - 1) * S + K, HALF = (K - 1) / S, NP = VEC / 2;
    const int C = CT ? CT : C_rt;
    extern __shared__ __align__(16) float s_dyn[];
    constexpr int LXS = 32;                               // lane stride of the tap table (LX <= 32): tap offsets are immediates
    float* s_w = s_dyn;                                   // [K*K][LXS][VEC] taps of this block's channel chunk
    float* s_red = s_dyn + K * K * LXS * VEC;              // [LY][LX][VEC] partial-sum reduction
    const int tile = blockIdx.x / strip_blocks, sb = blockIdx.x % strip_blocks, chunk = blockIdx.y, b = blockIdx.z;
    const int groups = gridDim.x;
    const int lx = threadIdx.x % LX, ly = threadIdx.x / LX;
    const int Cv = C / VEC;
    const int cv = chunk * LX + lx;
    const int strip = sb * LY + ly;
    const bool live = cv < Cv && strip * TW < Wo;
    for (int i = threadIdx.x; i < K * K * LXS; i += blockDim.x) {
        const int t = i / LXS, l = i % LXS, c = chunk * LX + l;
        float wv[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) wv[e] = 0.f;
        if (l < LX && c < Cv) VecIO<VEC>::load(wt + (int64_t)t * C + c * VEC, wv);
        VecIO<VEC>::store(s_w + (size_t)i * VEC, wv);
    }
    __syncthreads();
    float sum[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) sum[e] = 0.f;
    const int row0 = tile * rows_per_tile, row1 = min(Ho, row0 + rows_per_tile);
    if (live && row0 < row1) {
        f2_t sc[NP], sh[NP];
        PairIO<NP>::load(scale + cv * VEC, sc);
        PairIO<NP>::load(shift + cv * VEC, sh);
        const int ox0 = strip * TW, ixb = ox0 * S - pad_l;
        unsigned cmask = 0;                               // bit j: input column ixb + j exists
#pragma unroll
        for (int j = 0; j < SPAN; ++j) cmask |= (ixb + j >= 0 && ixb + j < W) ? (1u << j) : 0u;
        const int64_t row_stride = (int64_t)W * C;
        const float* xcol = x + (int64_t)b * H * row_stride + (int64_t)ixb * C + cv * VEC;   // (row 0, column ixb): only valid columns are dereferenced
        float* yb = y + ((int64_t)b * Ho * Wo + ox0) * C + cv * VEC;
        const float* wlane = s_w + lx * VEC;
        // virtual row vy = input row + pad_t; rows that exist and that this tile needs: [vy_lo, vy_hi]
        const int vy_lo = pad_t, vy_hi = min(H - 1 + pad_t, (row1 - 1) * S + K - 1);
        xcol -= pad_t * row_stride;
        auto load_row = [&](int vy, f2_t (&dst)[SPAN][NP]) {
            if (vy >= vy_lo && vy <= vy_hi) {
                const float* rp = xcol + vy * row_stride;
#pragma unroll
                for (int j = 0; j < SPAN; ++j) {
                    if (cmask & (1u << j)) PairIO<NP>::load(rp + j * C, dst[j]);
                    else {
#pragma unroll
                        for (int q = 0; q < NP; ++q) dst[j][q] = 0ull;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < SPAN; ++j)
#pragma unroll
                    for (int q = 0; q < NP; ++q) dst[j][q] = 0ull;
            }
        };
        f2_t acc[R][TW][NP], v[SPAN][NP], vn[SPAN][NP];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int t = 0; t < TW; ++t)
#pragma unroll
                for (int q = 0; q < NP; ++q) acc[r][t][q] = 0ull;
        load_row(row0 * S, v);
        // iteration m handles virtual rows m*S .. m*S+S-1; the oldest pending output row is m - HALF
        for (int m = row0; m < row1 + HALF; ++m) {
#pragma unroll
            for (int sub = 0; sub < S; ++sub) {
                load_row(m * S + sub + 1, vn);            // in flight while row m*S+sub is consumed
#pragma unroll
                for (int ky = sub; ky < K; ky += S) {      // ky with (vy - ky) divisible by S
                    const int slot = HALF - (ky - sub) / S;
#pragma unroll
                    for (int kx = 0; kx < K; ++kx) {
                        f2_t w[NP];
                        PairIO<NP>::load_shared(wlane + (ky * K + kx) * LXS * VEC, w);
#pragma unroll
                        for (int t = 0; t < TW; ++t)
#pragma unroll
                            for (int q = 0; q < NP; ++q) acc[slot][t][q] = f2_fma(v[kx + t * S][q], w[q], acc[slot][t][q]);
                    }
                }
                if (sub == (K - 1) % S) {      // the oldest pending output row has now seen its last input row
                    const int oy = m - HALF;
                    if (oy >= row0) {          // (oy < row1 by the loop bound)
                        float* yrow = yb + (int64_t)oy * Wo * C;
#pragma unroll
                        for (int t = 0; t < TW; ++t) {
                            if (ox0 + t < Wo) {
                                float r[VEC];
#pragma unroll
                                for (int q = 0; q < NP; ++q) {
                                    f2_unpack(f2_fma(acc[0][t][q], sc[q], sh[q]), r[2 * q], r[2 * q + 1]);
                                    r[2 * q] = act_fast(r[2 * q], act); r[2 * q + 1] = act_fast(r[2 * q + 1], act);
                                    sum[2 * q] += r[2 * q]; sum[2 * q + 1] += r[2 * q + 1];
                                }
                                PairIO<NP>::store(yrow + t * C, r);
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < SPAN; ++j)
#pragma unroll
                    for (int q = 0; q < NP; ++q) v[j][q] = vn[j][q];
            }
            // rotate the ring: slot r <- slot r+1, newest slot cleared
#pragma unroll
            for (int r = 0; r + 1 < R; ++r)
#pragma unroll
                for (int t = 0; t < TW; ++t)
#pragma unroll
                    for (int q = 0; q < NP; ++q) acc[r][t][q] = acc[r + 1][t][q];
#pragma unroll
            for (int t = 0; t < TW; ++t)
#pragma unroll
                for (int q = 0; q < NP; ++q) acc[R - 1][t][q] = 0ull;
        }
    }
    if (partial) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) s_red[(ly * LX + lx) * VEC + e] = sum[e];
        __syncthreads();
        if (ly == 0 && cv < Cv) {
            float t[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) t[e] = s_red[lx * VEC + e];
            for (int r = 1; r < LY; ++r)
#pragma unroll
                for (int e = 0; e < VEC; ++e) t[e] += s_red[(r * LX + lx) * VEC + e];
            VecIO<VEC>::store(partial + ((int64_t)b * groups + blockIdx.x) * C + cv * VEC, t);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 5x5 stride-1 depthwise at the SMALL spatial sizes (14x14, 7x7: EfficientNet-B0 blocks 4.x / 5.x / 6.x), staged through
// shared memory. dw2_kernel reads its input rows straight into registers and relies on occupancy to cover HBM latency; at
// K = 5 it needs ~112 registers, so an SM holds 16 warps with 2 KB of loads in flight each = 32 KB where ~35 KB are needed
// (measured: 2.0-2.5 TB/s on these layers). Here the loads cost no registers: a block owns one 64-channel chunk, walks over
// the frames FPB at a time, and cp.async copies the NEXT iteration's [FPB][HW*HW][64 channels] tile (50 / 25 KB) into the other
// half of a double buffer while the current one is convolved out of shared memory (two / three blocks per SM).
// Arithmetic: packed fma.rn.f32x2 on channel pairs, the 25 taps in registers, output-stationary (see below), bn2 / FiLM
// scale-shift + SiLU, deterministic SE squeeze sums per (frame, chunk) -> partial[b][0][C] (one group).
// grid (channel chunks of 64, frame slices); block 256 = 32 channel pairs x [strips (4 columns) x 2 row halves] x FPB frames.
// ------------------------------------------------------------------------------------------------
template <int HW>
__global__ void __launch_bounds__(256, 2)
dw5s_kernel(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ scale, const float* __restrict__ shift,
            float* __restrict__ y, float* __restrict__ partial, int B, int C, int act) {
    constexpr int K = 5, TW = kDwTW, SPAN = TW + K - 1, PAD = 2;
    constexpr int STRIPS = (HW + TW - 1) / TW;            // 4 (14x14) or 2 (7x7)
    constexpr int RS = 2, RH = (HW + RS - 1) / RS;        // a strip's rows are split between RS warps (RH rows each)
    constexpr int WPF = STRIPS * RS;                      // warps per frame: 8 or 4
    constexpr int FPB = 8 / WPF;                          // frames per block iteration: 1 or 2
    constexpr int PIX = HW * HW;
    constexpr int TILE_F2 = FPB * PIX * 32;               // channel pairs per stage
    extern __shared__ __align__(16) float s_dyn[];
    f2_t* s_w = reinterpret_cast<f2_t*>(s_dyn);                          // [25][32]
    f2_t* s_x = s_w + K * K * 32;                                        // [2][FPB][PIX][32]
    float* s_red = reinterpret_cast<float*>(s_x + 2 * TILE_F2);          // [256][2]
    const int chunk = blockIdx.x, c0 = chunk * 64;
    const int cw = min(64, C - c0);                       // channels of this chunk (64, or 32 in a tail chunk); C % 4 == 0
    const int lx = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int fl = wid / WPF, strip = (wid % WPF) / RS, half = wid % RS;
    const bool lane_live = 2 * lx < cw;
    for (int i = threadIdx.x; i < K * K * 32; i += 256) {
        const int tp = i >> 5, l = i & 31;
        s_w[i] = (2 * l < cw) ? __ldg(reinterpret_cast<const f2_t*>(wt + (int64_t)tp * C + c0 + 2 * l)) : 0ull;
    }
    f2_t sc = 0ull, sh = 0ull;
    if (lane_live) { sc = __ldg(reinterpret_cast<const f2_t*>(scale + c0 + 2 * lx)); sh = __ldg(reinterpret_cast<const f2_t*>(shift + c0 + 2 * lx)); }
    const int units_per_pix = cw >> 2;                    // 16-byte units per pixel of this chunk
    auto issue_tile = [&](int f0, int stage) {
        const int nf = min(FPB, B - f0);
        const int total = nf * PIX * units_per_pix;
        const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(s_x + (size_t)stage * TILE_F2);
        for (int u = threadIdx.x; u < total; u += 256) {
            const int q = u % units_per_pix, p = u / units_per_pix;       // p = frame-local pixel index (fl * PIX + pixel)
            const float* src = x + ((int64_t)f0 * PIX + p) * C + c0 + 4 * q;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)(p * 256 + q * 16)), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int ox0 = strip * TW, ixb = ox0 - PAD;
    unsigned cmask = 0;
#pragma unroll
    for (int j = 0; j < SPAN; ++j) cmask |= (ixb + j >= 0 && ixb + j < HW) ? (1u << j) : 0u;
    __syncthreads();
    f2_t wreg[K * K];                                     // the 25 taps of this lane's channel pair stay in registers
#pragma unroll
    for (int i = 0; i < K * K; ++i) wreg[i] = s_w[i * 32 + lx];
    int f0 = blockIdx.y * FPB, stage = 0;
    if (f0 < B) issue_tile(f0, 0);
    for (; f0 < B; f0 += gridDim.y * FPB, stage ^= 1) {
        const int fnext = f0 + gridDim.y * FPB;
        if (fnext < B) issue_tile(fnext, stage ^ 1); else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const int f = f0 + fl;
        float sum0 = 0.f, sum1 = 0.f;
        if (f < B && lane_live) {
            const f2_t* xs = s_x + (size_t)stage * TILE_F2 + (size_t)fl * PIX * 32 + lx;
            float* yb = y + ((int64_t)f * PIX + ox0) * C + c0 + 2 * lx;
            // Output-stationary, two output rows at a time: they gather their (up to) six input rows from shared memory --
            // re-reading a staged row costs an LDS, not HBM traffic -- so there is no accumulator ring to rotate and no row
            // phase logic. A warp is one strip of half the rows of one frame: interior strips run without column tests.
            auto out_rows = [&](int oy, bool second, auto interior) {
                f2_t acc0[TW], acc1[TW];
#pragma unroll
                for (int t = 0; t < TW; ++t) { acc0[t] = 0ull; acc1[t] = 0ull; }
#pragma unroll
                for (int r = 0; r < K + 1; ++r) {                 // input row oy - PAD + r feeds row oy (ky = r) and oy + 1 (ky = r - 1)
                    const int iy = oy - PAD + r;
                    if (iy < 0 || iy >= HW) continue;
                    f2_t v[SPAN];
#pragma unroll
                    for (int j = 0; j < SPAN; ++j)
                        v[j] = (decltype(interior)::value || (cmask & (1u << j))) ? xs[(iy * HW + ixb + j) * 32] : 0ull;
#pragma unroll
                    for (int kx = 0; kx < K; ++kx)
#pragma unroll
                        for (int t = 0; t < TW; ++t) {
                            if (r < K) acc0[t] = f2_fma(v[kx + t], wreg[(r < K ? r : 0) * K + kx], acc0[t]);
                            if (r > 0) acc1[t] = f2_fma(v[kx + t], wreg[(r > 0 ? r - 1 : 0) * K + kx], acc1[t]);
                        }
                }
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    if (rr == 1 && !second) break;
                    float* yrow = yb + (int64_t)(oy + rr) * HW * C;
#pragma unroll
                    for (int t = 0; t < TW; ++t) {
                        if (decltype(interior)::value || ox0 + t < HW) {
                            float r0, r1;
                            f2_unpack(f2_fma(rr ? acc1[t] : acc0[t], sc, sh), r0, r1);
                            r0 = act_fast(r0, act); r1 = act_fast(r1, act);
                            sum0 += r0; sum1 += r1;
                            *reinterpret_cast<float2*>(yrow + t * C) = make_float2(r0, r1);
                        }
                    }
                }
            };
            const int row_lo = half * RH, row_hi = min(HW, row_lo + RH);
            if (cmask == (1u << SPAN) - 1u && ox0 + TW <= HW) {
                for (int oy = row_lo; oy < row_hi; oy += 2) out_rows(oy, oy + 1 < row_hi, std::true_type{});
            } else {
                for (int oy = row_lo; oy < row_hi; oy += 2) out_rows(oy, oy + 1 < row_hi, std::false_type{});
            }
        }
        if (partial) {
            s_red[2 * threadIdx.x] = sum0; s_red[2 * threadIdx.x + 1] = sum1;
            __syncthreads();
            if (wid % WPF == 0 && f < B && lane_live) {
                float a = 0.f, b2 = 0.f;
#pragma unroll
                for (int r = 0; r < WPF; ++r) { a += s_red[2 * (threadIdx.x + 32 * r)]; b2 += s_red[2 * (threadIdx.x + 32 * r) + 1]; }   // fixed order
                *reinterpret_cast<float2*>(partial + (int64_t)f * C + c0 + 2 * lx) = make_float2(a, b2);
            }
        }
        __syncthreads();            // everyone is done with this stage before the next iteration's copies overwrite it
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

static int g_dw5s = 1;      // dev A/B switch (orbit_set_global_option "dw5_staged")
void set_dw5_staged(int on) { g_dw5s = on; }
int get_dw5_staged() { return g_dw5s; }

template <int HW>
static int launch_dw5s(const float* x, const float* wt, const float* scale, const float* shift, float* y, float* partial, int B, int C,
                       int act, cudaStream_t st) {
    constexpr int STRIPS = (HW + kDwTW - 1) / kDwTW, FPB = 8 / (STRIPS * 2), PIX = HW * HW;
    const size_t smem = sizeof(f2_t) * (25 * 32 + 2 * (size_t)FPB * PIX * 32) + sizeof(float) * 512;
    auto fn = dw5s_kernel<HW>;
    ORBIT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int chunks = ceil_div(C, 64);
    // 2 (14x14: 109 KB) or 3 (7x7: 59 KB) blocks per SM
    const int iters = ceil_div(B, FPB);
    const int slices = std::max(1, std::min(iters, std::max(1, ((HW == 14 ? 2 : 3) * 148) / chunks)));
    fn<<<dim3(chunks, slices), 256, smem, st>>>(x, wt, scale, shift, y, partial, B, C, act);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

int launch_depthwise(const float* x, const float* wt, const float* scale, const float* shift, float* y, float* partial,
                     int B, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad_t, int pad_l, int act,
                     cudaStream_t st) {
    if (C % 4) return ORBIT_ERR_UNSUPPORTED;
    if (g_dw5s && k == 5 && stride == 1 && H == W && pad_t == 2 && pad_l == 2 && (act == ACT_SILU || act == ACT_NONE) &&
        dw_plan(C, Ho, Wo, k, stride).groups == 1) {
        // measured on B200 (us per 1,024 frames, dw2_kernel -> staged): 7x7x1152 210 -> 155, 14x14x672 419 -> 368, 14x14x480 310 -> 266
        if (H == 7) return launch_dw5s<7>(x, wt, scale, shift, y, partial, B, C, act, st);
        if (H == 14) return launch_dw5s<14>(x, wt, scale, shift, y, partial, B, C, act, st);
    }
    const DwPlan pl = dw_plan(C, Ho, Wo, k, stride);
    dim3 grid(pl.groups, pl.nchunks, B), block(pl.LX * pl.LY);
    const size_t smem = sizeof(float) * ((size_t)pl.LY * pl.LX * pl.VEC + (size_t)k * k * 32 * pl.VEC);
#define ORBIT_DW2_LAUNCH(KK, SS, VV, MB, CC)                                                                          \
    {                                                                                                                 \
        dw2_kernel<KK, SS, VV, MB, CC><<<grid, block, smem, st>>>(x, wt, scale, shift, y, partial, H, W, C, Ho, Wo, pad_t,   \
                                                                 pad_l, act, pl.LX, pl.LY, pl.rows_per_tile, pl.strip_blocks); \
        ORBIT_RETURN_IF_LAUNCH_FAILED();                                                                              \
        return ORBIT_OK;                                                                                              \
    }
#define ORBIT_DW2_CT(KK, SS, VV, CC) if (k == KK && stride == SS && C == CC) ORBIT_DW2_LAUNCH(KK, SS, VV, 2, CC)
#define ORBIT_DW2_ANY(KK, SS, VV) if (k == KK && stride == SS) ORBIT_DW2_LAUNCH(KK, SS, VV, 2, 0)
    // the EfficientNet-B0 depthwise shapes get compile-time channel counts; everything else the generic kernels
    ORBIT_DW2_CT(5, 2, 2, 144) ORBIT_DW2_CT(5, 2, 2, 672)
    ORBIT_DW2_CT(5, 1, 2, 240) ORBIT_DW2_CT(5, 1, 2, 480) ORBIT_DW2_CT(5, 1, 2, 672) ORBIT_DW2_CT(5, 1, 2, 1152)
    ORBIT_DW2_ANY(3, 1, 4) ORBIT_DW2_ANY(3, 2, 4) ORBIT_DW2_ANY(5, 1, 2) ORBIT_DW2_ANY(5, 2, 2)
#undef ORBIT_DW2_CT
#undef ORBIT_DW2_ANY
#undef ORBIT_DW2_LAUNCH
    return ORBIT_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// Fused MBConv front half:  expand 1x1 + bn1 + SiLU  ->  depthwise KxK + bn2(+FiLM) + SiLU (+ SE squeeze partials).
//
// Reference op sites: timm InvertedResidual conv_pw/bn1 -> conv_dw/bn2 (FiLM site, model/film.py:43-44) inside the
// extractor invoked at model/few_shot_recognisers.py:114-117,143-146. At the 112x112 and 56x56 stages the 6x-expanded
// tensor is 27 % of ALL layer-boundary traffic of the network (written by the expand GEMM, read back by the depthwise
// kernel): here it never reaches HBM. Per virtual input row of a block (= a set of adjacent 4-column strips x a chunk of
// 2 LX channels of one frame):
//   1. staging   the row segment of the 16/24-channel block input arrives in shared memory two rows ahead (coalesced
//                float4 loads, one __syncthreads per row covers everything);
//   2. expand    warp w owns the 8-channel group w (its weight fragments live in registers) and runs 16-pixel x 8-channel
//                warp-level tensor-core tiles (mma.sync m16n8k8 tf32) over the segment: every (pixel, channel) is computed
//                exactly ONCE, fp32-grade through the 3xTF32 split (hi.hi + lo.hi + hi.lo; each hi.hi product tile is
//                added in fp32 registers with round-to-nearest, because tensor-core accumulation truncates), then
//                bn1 + SiLU on packed fp32x2, zeroed outside the image (the zero padding applies to the EXPANDED
//                tensor), stored [pixel][channel pair] to shared memory (double buffered);
//   3. depthwise thread (lx, ly) feeds the SPAN pixels of strip ly for channel pair lx from shared memory into the same
//                rolling accumulators as dw2_kernel and emits bn2/FiLM + SiLU outputs and the SE squeeze partial sums.
// (Round-2 history, see DESIGN.md: a CUDA-core expand -- 2 x CIN scalar FMAs per value -- was correct but ALU-bound and
// 20 % SLOWER than the unfused pair; loading the input inside the expand loop exposed an L2 round trip per pixel.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2], const float (&c)[4]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
// fp32 -> tf32 hi (round to nearest, ties away: add half a tf32 ulp, clear 13 bits) and the exact residual lo = v - hi
// (<= 13 significant bits; the tensor core reads its top 11)
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}

constexpr int kMbxXPad = 4;      // floats of padding per staged input pixel: conflict-free A-fragment loads
constexpr int kMbxEPad = 4;      // channel pairs of padding per expanded pixel: conflict-free C-fragment stores

#ifndef ORBIT_MBX_MINB
#define ORBIT_MBX_MINB 2      // 128 registers: 3 blocks per SM (96 registers) spills the fragment registers and measured 15-25 % slower
#endif
template <int K, int S, int CIN, int CT>
__global__ void __launch_bounds__(192, ORBIT_MBX_MINB)
mbx_kernel(const float* __restrict__ xin, const float* __restrict__ we, const float* __restrict__ scale1,
           const float* __restrict__ shift1, const float* __restrict__ wt, const float* __restrict__ scale,
           const float* __restrict__ shift, float* __restrict__ y, float* __restrict__ partial, int H, int W, int C_rt,
           int Ho, int Wo, int pad_t, int pad_l, int LX, int LY, int rows_per_tile, int strip_blocks) {
    constexpr int TW = kDwTW, R = (K + S - 1) / S, SPAN = (TW - 1) * S + K, HALF = (K - 1) / S, VEC = 2;
    constexpr int KS = CIN / 8;                           // k-steps of the expand MMA
    constexpr int XS = CIN + kMbxXPad;                    // staged pixel stride (floats)
    const int C = CT ? CT : C_rt;
    extern __shared__ __align__(16) float s_dyn[];
    constexpr int LXS = 32;
    const int P = (LY * TW - 1) * S + K;                  // input columns the block's strips touch
    const int PM = (P + 15) / 16 * 16;                    // ... rounded up to whole 16-pixel MMA tiles
    const int ES = LX + kMbxEPad;                         // expanded pixel stride (channel pairs)
    float* s_w = s_dyn;                                   // [K*K][LXS][2] depthwise taps of this block's channel chunk
    float* s_red = s_w + K * K * LXS * VEC;               // [LY][LX][2] partial-sum reduction
    float* s_e = s_red + LY * LX * VEC;                   // [2][PM][ES][2] expanded rows (double buffered)
    float* s_x = s_e + 2 * PM * ES * VEC;                 // [2][PM][XS] block-input row segments (double buffered)
    const int tile = blockIdx.x / strip_blocks, sb = blockIdx.x % strip_blocks, chunk = blockIdx.y, b = blockIdx.z;
    const int groups = gridDim.x;
    const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int lx = tid % LX, ly = tid / LX;               // depthwise role (threads beyond LX*LY only stage and expand)
    const int Cv = C / VEC;
    const int cv = chunk * LX + lx;
    const int strip = sb * LY + ly;
    const bool live = ly < LY && cv < Cv && strip * TW < Wo;
    for (int i = tid; i < K * K * LXS; i += nthreads) {
        const int t = i / LXS, l = i % LXS, c = chunk * LX + l;
        float wv[VEC] = {0.f, 0.f};
        if (l < LX && c < Cv) VecIO<VEC>::load(wt + (int64_t)t * C + c * VEC, wv);
        VecIO<VEC>::store(s_w + (size_t)i * VEC, wv);
    }
    // zero both staging buffers once: padding floats and pixels beyond P are read by the MMA fragments
    for (int i = tid; i < 2 * PM * XS; i += nthreads) s_x[i] = 0.f;

    // ---- expand role: warp -> 8-channel groups warp, warp + nwarps, ... (LX / 4 groups); fragments of m16n8k8:
    //      g = lane / 4, t = lane % 4;  A: (pixel g | g+8, k t | t+4);  B: (k t | t+4, channel g);  C: (pixel g | g+8, channel 2t, 2t+1)
    const int g = lane >> 2, t4 = lane & 3;
    // warp -> a PAIR of 8-channel groups (2 pair, 2 pair + 1) and a class of 16-pixel tiles (mt = mclass, mclass + mclasses, ..):
    // the tf32 split of an A fragment is shared by the two groups (a first version gave every warp one group and all tiles:
    // every warp then re-split every fragment).
    const int ngroups = LX / 4, npairs = (ngroups + 1) / 2, mclasses = max(1, nwarps / npairs);
    const int pair = warp % npairs, mclass = warp / npairs;
    constexpr int kMaxGroups = 2;
    uint32_t bh[kMaxGroups][KS][2], bl[kMaxGroups][KS][2];
    f2_t s1p[kMaxGroups], h1p[kMaxGroups];
#pragma unroll
    for (int gi = 0; gi < kMaxGroups; ++gi) {
        const int grp = 2 * pair + gi;
        const int cb = (chunk * LX + grp * 4) * VEC + g;            // channel of this lane's B column
        const int cc = (chunk * LX + grp * 4 + t4) * VEC;           // first channel of this lane's C pair
        const bool bok = grp < ngroups && cb < C, cok = grp < ngroups && cc < C;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            split_tf32(bok ? __ldg(we + (int64_t)cb * CIN + ks * 8 + t4) : 0.f, bh[gi][ks][0], bl[gi][ks][0]);
            split_tf32(bok ? __ldg(we + (int64_t)cb * CIN + ks * 8 + t4 + 4) : 0.f, bh[gi][ks][1], bl[gi][ks][1]);
        }
        s1p[gi] = cok ? __ldg(reinterpret_cast<const f2_t*>(scale1 + cc)) : 0ull;
        h1p[gi] = cok ? __ldg(reinterpret_cast<const f2_t*>(shift1 + cc)) : 0ull;
    }
    const bool second = 2 * pair + 1 < ngroups;          // warp-uniform: the pair's second group exists

    float sum[VEC] = {0.f, 0.f};
    const int row0 = tile * rows_per_tile, row1 = min(Ho, row0 + rows_per_tile);
    const int bx0 = sb * LY * TW * S - pad_l;             // input column of expanded-row slot 0
    // virtual row vy = input row + pad_t; rows that exist and that this tile needs: [vy_lo, vy_hi]
    const int vy_lo = pad_t, vy_hi = min(H - 1 + pad_t, (row1 - 1) * S + K - 1);
    const float* xfr = xin + (int64_t)b * H * W * CIN;
    const uint32_t e_base = (uint32_t)__cvta_generic_to_shared(s_e);
    const uint32_t e_row_bytes = (uint32_t)(PM * ES * VEC * 4);
    const uint32_t x_base = (uint32_t)__cvta_generic_to_shared(s_x);
    const uint32_t x_row_bytes = (uint32_t)(PM * XS * 4);

    // Everything below that does not depend on the row is computed ONCE (ncu of the first tensor-core version: 45 % of
    // the 618 instructions per warp per row were address / predicate arithmetic re-derived every row).
    // ---- input staging: the row segment [bx0, bx0 + P) x CIN of a virtual row is one contiguous byte range of the NHWC
    // tensor: every thread fetches at most kXLoads float4 of it (coalesced), two rows ahead of its use, so the L2 / HBM
    // latency hides behind a whole row of expand + depthwise work.
    constexpr int kXLoads = 3;
    const int nf4 = P * CIN / 4;
    int x_goff[kXLoads];          // float offset of this thread's j-th float4 inside the row segment (or -1: nothing to fetch)
    uint32_t x_soff[kXLoads];     // its byte offset inside a staging buffer
#pragma unroll
    for (int j = 0; j < kXLoads; ++j) {
        const int f = tid + j * nthreads;
        const int pix = (4 * f) / CIN, kk = (4 * f) % CIN, col = bx0 + pix;
        x_goff[j] = (f < nf4 && col >= 0 && col < W) ? 4 * f : -1;
        x_soff[j] = (uint32_t)((pix * XS + kk) * 4);
    }
    const int64_t x_row_stride = (int64_t)W * CIN;
    // ---- expand: image-validity of this lane's pixels per 16-pixel tile (bit 2 mt: row g, bit 2 mt + 1: row g + 8)
    uint32_t okbits = 0;
    for (int mt = 0; mt < PM / 16; ++mt) {
        const int c_lo = bx0 + mt * 16 + g, c_hi = c_lo + 8;
        okbits |= ((c_lo >= 0 && c_lo < W) ? 1u : 0u) << (2 * mt);
        okbits |= ((c_hi >= 0 && c_hi < W) ? 1u : 0u) << (2 * mt + 1);
    }
    const uint32_t a_off = (uint32_t)(((mclass * 16 + g) * XS + t4) * 4);                        // A fragment base inside a staging buffer
    const uint32_t c_off = (uint32_t)((((mclass * 16 + g) * ES + t4) + pair * 8) * VEC * 4);     // C fragment base inside an expanded-row buffer
    const bool expander = mclass < mclasses;       // warp-uniform
    const int n_mt = PM / 16;

    auto fetch_row = [&](const float* xrow, bool row_ok, float4 (&xr)[kXLoads]) {
#pragma unroll
        for (int j = 0; j < kXLoads; ++j)
            xr[j] = (row_ok && x_goff[j] >= 0) ? ldg4(xrow + x_goff[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto stash_row = [&](uint32_t xbuf, const float4 (&xr)[kXLoads]) {
#pragma unroll
        for (int j = 0; j < kXLoads; ++j)
            if (tid + j * nthreads < nf4)
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(xbuf + x_soff[j]),
                             "f"(xr[j].x), "f"(xr[j].y), "f"(xr[j].z), "f"(xr[j].w) : "memory");
    };
    // ---- expand phase: staged row at xbuf -> expanded row at ebuf ---------------------------------------------------
    auto expand_row = [&](uint32_t xbuf, uint32_t ebuf, bool row_ok) {
        if (!expander) return;
        const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t ad = xbuf + a_off, d = ebuf + c_off, ok = (row_ok ? okbits : 0u) >> (2 * mclass);
        for (int mt = mclass; mt < n_mt; mt += mclasses, ad += (uint32_t)(mclasses * 16 * XS * 4),
                 d += (uint32_t)(mclasses * 16 * ES * VEC * 4), ok >>= 2 * mclasses) {
            float acc0[4] = {0.f, 0.f, 0.f, 0.f}, cor0[4] = {0.f, 0.f, 0.f, 0.f};
            float acc1[4] = {0.f, 0.f, 0.f, 0.f}, cor1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                float a0, a1, a2, a3;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a0) : "r"(ad + (uint32_t)(ks * 32)) : "memory");
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a1) : "r"(ad + (uint32_t)(ks * 32 + 8 * XS * 4)) : "memory");
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a2) : "r"(ad + (uint32_t)(ks * 32 + 16)) : "memory");
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a3) : "r"(ad + (uint32_t)(ks * 32 + 8 * XS * 4 + 16)) : "memory");
                uint32_t ah[4], al[4];
                split_tf32(a0, ah[0], al[0]); split_tf32(a1, ah[1], al[1]);
                split_tf32(a2, ah[2], al[2]); split_tf32(a3, ah[3], al[3]);
                float m4[4];
                mma_tf32_16x8x8(m4, ah, bh[0][ks], zero4);              // hi.hi of one k-step, added below with RN
                mma_tf32_16x8x8(cor0, al, bh[0][ks], cor0);
                mma_tf32_16x8x8(cor0, ah, bl[0][ks], cor0);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc0[i] += m4[i];
                if (second) {
                    mma_tf32_16x8x8(m4, ah, bh[1][ks], zero4);
                    mma_tf32_16x8x8(cor1, al, bh[1][ks], cor1);
                    mma_tf32_16x8x8(cor1, ah, bl[1][ks], cor1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc1[i] += m4[i];
                }
            }
            // bn1 + SiLU on channel pairs (packed fp32x2 around the four MUFU ops), zero outside the image
            f2_t lo2 = f2_fma(f2_pack(acc0[0] + cor0[0], acc0[1] + cor0[1]), s1p[0], h1p[0]);
            f2_t hi2 = f2_fma(f2_pack(acc0[2] + cor0[2], acc0[3] + cor0[3]), s1p[0], h1p[0]);
            lo2 = (ok & 1u) ? f2_silu_pair(lo2) : 0ull;
            hi2 = (ok & 2u) ? f2_silu_pair(hi2) : 0ull;
            asm volatile("st.shared.b64 [%0], %1;" ::"r"(d), "l"(lo2) : "memory");
            asm volatile("st.shared.b64 [%0], %1;" ::"r"(d + (uint32_t)(8 * ES * VEC * 4)), "l"(hi2) : "memory");
            if (second) {
                lo2 = f2_fma(f2_pack(acc1[0] + cor1[0], acc1[1] + cor1[1]), s1p[1], h1p[1]);
                hi2 = f2_fma(f2_pack(acc1[2] + cor1[2], acc1[3] + cor1[3]), s1p[1], h1p[1]);
                lo2 = (ok & 1u) ? f2_silu_pair(lo2) : 0ull;
                hi2 = (ok & 2u) ? f2_silu_pair(hi2) : 0ull;
                asm volatile("st.shared.b64 [%0], %1;" ::"r"(d + (uint32_t)(4 * VEC * 4)), "l"(lo2) : "memory");
                asm volatile("st.shared.b64 [%0], %1;" ::"r"(d + (uint32_t)((8 * ES + 4) * VEC * 4)), "l"(hi2) : "memory");
            }
        }
    };

    f2_t sc[1], sh[1];
    sc[0] = sh[0] = 0ull;
    if (live) { PairIO<1>::load(scale + cv * VEC, sc); PairIO<1>::load(shift + cv * VEC, sh); }
    const int ox0 = strip * TW;
    const float* wlane = s_w + lx * VEC;
    const uint32_t e_mine = (uint32_t)((ly * TW * S * ES + lx) * VEC * 4);   // slot of this strip's first input column
    uint32_t colmask = 0;                              // bit t: output column ox0 + t exists
#pragma unroll
    for (int t = 0; t < TW; ++t) colmask |= (ox0 + t < Wo) ? (1u << t) : 0u;
    f2_t acc[R][TW], v[SPAN];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int t = 0; t < TW; ++t) acc[r][t] = 0ull;

    if (row0 < row1) {          // block-uniform
        // pipeline: s_x holds the block input of virtual rows v+1 (being expanded) and v+2 (arriving); s_e holds the expanded
        // rows v (being consumed) and v+1 (being produced). One __syncthreads per row.
        float4 xr[kXLoads];
        const int v0 = row0 * S;
        const float* xrow = xfr + ((int64_t)(v0 - pad_t) * W + bx0) * CIN;      // segment start of virtual row v0 (may point outside: guarded)
        const uint32_t xbuf[2] = {x_base, x_base + x_row_bytes}, ebuf[2] = {e_base, e_base + e_row_bytes};
        auto row_ok = [&](int vy) { return vy >= vy_lo && vy <= vy_hi; };
        fetch_row(xrow, row_ok(v0), xr);
        __syncthreads();                                  // tap table + zeroed staging buffers
        stash_row(xbuf[v0 & 1], xr);
        fetch_row(xrow + x_row_stride, row_ok(v0 + 1), xr);
        __syncthreads();
        expand_row(xbuf[v0 & 1], ebuf[v0 & 1], row_ok(v0));
        stash_row(xbuf[(v0 + 1) & 1], xr);
        __syncthreads();
        xrow += 2 * x_row_stride;                         // -> virtual row v0 + 2
        float* yrow = y + (((int64_t)b * Ho + row0) * Wo + ox0) * C + cv * VEC;   // output row row0 of this strip / channel pair
        const int64_t y_row_stride = (int64_t)Wo * C;
        int vy = v0;
        // iteration m handles virtual rows m*S .. m*S+S-1; the oldest pending output row is m - HALF
        for (int m = row0; m < row1 + HALF; ++m) {
#pragma unroll
            for (int sub = 0; sub < S; ++sub, ++vy, xrow += x_row_stride) {
                const int eb = vy & 1;
                fetch_row(xrow, row_ok(vy + 2), xr);                        // in flight during this row's work
                expand_row(xbuf[eb ^ 1], ebuf[eb ^ 1], row_ok(vy + 1));     // next row
                if (live) {
                    const uint32_t er = ebuf[eb] + e_mine;
#pragma unroll
                    for (int j = 0; j < SPAN; ++j)
                        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v[j]) : "r"(er + (uint32_t)(j * ES * VEC * 4)) : "memory");
#pragma unroll
                    for (int ky = sub; ky < K; ky += S) {      // ky with (vy - ky) divisible by S
                        const int slot = HALF - (ky - sub) / S;
#pragma unroll
                        for (int kx = 0; kx < K; ++kx) {
                            f2_t w[1];
                            PairIO<1>::load_shared(wlane + (ky * K + kx) * LXS * VEC, w);
#pragma unroll
                            for (int t = 0; t < TW; ++t) acc[slot][t] = f2_fma(v[kx + t * S], w[0], acc[slot][t]);
                        }
                    }
                    if (sub == (K - 1) % S) {      // the oldest pending output row has now seen its last input row
                        if (m - HALF >= row0) {    // (m - HALF < row1 by the loop bound)
#pragma unroll
                            for (int t = 0; t < TW; ++t) {
                                if (colmask & (1u << t)) {
                                    const f2_t o = f2_silu_pair(f2_fma(acc[0][t], sc[0], sh[0]));
                                    float r0, r1;
                                    f2_unpack(o, r0, r1);
                                    sum[0] += r0; sum[1] += r1;
                                    *reinterpret_cast<f2_t*>(yrow + t * C) = o;
                                }
                            }
                            yrow += y_row_stride;
                        }
                    }
                }
                stash_row(xbuf[eb], xr);   // s_x[eb] (row vy) was last read by the previous iteration's expand: free
                __syncthreads();           // s_e[eb] consumed by everyone; s_e[eb^1] and s_x[eb] complete
            }
            // rotate the ring: slot r <- slot r+1, newest slot cleared
#pragma unroll
            for (int r = 0; r + 1 < R; ++r)
#pragma unroll
                for (int t = 0; t < TW; ++t) acc[r][t] = acc[r + 1][t];
#pragma unroll
            for (int t = 0; t < TW; ++t) acc[R - 1][t] = 0ull;
        }
    } else {
        __syncthreads();
    }
    if (partial) {
        if (ly < LY) {
            s_red[(ly * LX + lx) * VEC] = sum[0];
            s_red[(ly * LX + lx) * VEC + 1] = sum[1];
        }
        __syncthreads();
        if (ly == 0 && cv < Cv) {
            float t[VEC] = {s_red[lx * VEC], s_red[lx * VEC + 1]};
            for (int r = 1; r < LY; ++r) { t[0] += s_red[(r * LX + lx) * VEC]; t[1] += s_red[(r * LX + lx) * VEC + 1]; }
            VecIO<VEC>::store(partial + ((int64_t)b * groups + blockIdx.x) * C + cv * VEC, t);
        }
    }
}

struct MbxPlan { int LX, LY, nchunks, tiles, rows_per_tile, strip_blocks, groups, threads, P, PM; size_t smem; };
static MbxPlan mbx_plan(int Cin, int C, int Ho, int Wo, int k, int stride) {
    MbxPlan p;
    const int Cv = C / 2;
    p.nchunks = ceil_div(Cv, 24);                         // <= 24 channel pairs = 6 eight-channel MMA groups = 6 warps per block
    p.LX = ceil_div(ceil_div(Cv, p.nchunks), 4) * 4;      // channel pairs per block: whole 8-channel MMA groups
    const int strips = ceil_div(Wo, kDwTW);
    const int ly_max = std::max(1, 192 / p.LX);
    p.strip_blocks = ceil_div(strips, ly_max);
    p.LY = ceil_div(strips, p.strip_blocks);              // balanced strip blocks (14 strips -> 7 + 7, not 10 + 4)
    const int R = ceil_div(k, stride);
    const int want_tiles = std::min(ceil_div(Ho, R), Ho >= 112 ? 4 : (Ho >= 56 ? 2 : 1));
    p.rows_per_tile = ceil_div(ceil_div(Ho, want_tiles), R) * R;
    p.tiles = ceil_div(Ho, p.rows_per_tile);
    p.groups = p.tiles * p.strip_blocks;
    p.threads = std::max(32 * ((p.LX / 4 + 1) / 2), ceil_div(p.LX * p.LY, 32) * 32);    // one warp per pair of 8-channel groups at least (<= 192)
    p.P = (p.LY * kDwTW - 1) * stride + k;
    p.PM = ceil_div(p.P, 16) * 16;
    p.smem = sizeof(float) * ((size_t)k * k * 32 * 2 + (size_t)p.LY * p.LX * 2 + (size_t)2 * p.PM * (p.LX + kMbxEPad) * 2 +
                              (size_t)2 * p.PM * (Cin + kMbxXPad));
    return p;
}

bool mbx_supported(int cin, int k, int stride) { return (cin == 16 || cin == 24) && (k == 3 || k == 5) && (stride == 1 || stride == 2); }
// whether the fused kernel's shared-memory footprint and its 3-loads-per-thread input staging cover this geometry
bool mbx_fits(int Cin, int C, int Ho, int Wo, int k, int stride) {
    if (C % 8 || !mbx_supported(Cin, k, stride)) return false;
    const MbxPlan pl = mbx_plan(Cin, C, Ho, Wo, k, stride);
    return pl.smem <= 100 * 1024 && (size_t)pl.P * Cin / 4 <= (size_t)3 * pl.threads && (pl.LX / 4 + 1) / 2 <= pl.threads / 32 && pl.threads <= 192;
}
int mbx_partial_groups(int Cin, int C, int Ho, int Wo, int k, int stride) {
    if (stride == 2 && mbs_supported(Cin, C, 2 * Ho, 2 * Wo, k, stride)) return mbs_partial_groups(Ho);   // register-resident variant
    return mbx_plan(16, C, Ho, Wo, k, stride).groups;
}

int launch_mbconv_expand_dw(const float* xin, const float* we, const float* scale1, const float* shift1, const float* wt,
                            const float* scale, const float* shift, float* y, float* partial, int B, int H, int W, int Cin,
                            int C, int Ho, int Wo, int k, int stride, int pad_t, int pad_l, cudaStream_t st) {
    // 3x3 stride-2 block with 16 input channels on an even size (TF-SAME: no padding above / left): register-resident variant
    if (pad_t == 0 && pad_l == 0 && H == 2 * Ho && W == 2 * Wo && mbs_supported(Cin, C, H, W, k, stride))
        return launch_mbconv_stream(xin, we, scale1, shift1, wt, scale, shift, y, partial, B, H, W, Cin, C, st);
    if (!mbx_fits(Cin, C, Ho, Wo, k, stride)) return ORBIT_ERR_UNSUPPORTED;
    const MbxPlan pl = mbx_plan(Cin, C, Ho, Wo, k, stride);
    dim3 grid(pl.groups, pl.nchunks, B), block(pl.threads);
    const size_t smem = pl.smem;
#define ORBIT_MBX(KK, SS, CI, CC)                                                                                     \
    if (k == KK && stride == SS && Cin == CI && (CC == 0 || C == CC)) {                                              \
        if (smem > 48 * 1024) ORBIT_CUDA(cudaFuncSetAttribute(mbx_kernel<KK, SS, CI, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        mbx_kernel<KK, SS, CI, CC><<<grid, block, smem, st>>>(xin, we, scale1, shift1, wt, scale, shift, y, partial, H, W, C, Ho, Wo, \
                                                             pad_t, pad_l, pl.LX, pl.LY, pl.rows_per_tile, pl.strip_blocks); \
        ORBIT_RETURN_IF_LAUNCH_FAILED();                                                                              \
        return ORBIT_OK;                                                                                              \
    }
    // the three EfficientNet-B0 blocks with 16 / 24 input channels get compile-time channel counts
    ORBIT_MBX(3, 2, 16, 96) ORBIT_MBX(3, 1, 24, 144) ORBIT_MBX(5, 2, 24, 144)
    ORBIT_MBX(3, 1, 16, 0) ORBIT_MBX(3, 2, 16, 0) ORBIT_MBX(5, 1, 16, 0) ORBIT_MBX(5, 2, 16, 0)
    ORBIT_MBX(3, 1, 24, 0) ORBIT_MBX(3, 2, 24, 0) ORBIT_MBX(5, 1, 24, 0) ORBIT_MBX(5, 2, 24, 0)
#undef ORBIT_MBX
    return ORBIT_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// Squeeze-excite gate: mean over the depthwise kernel's per-group sums -> FC (C -> R) + SiLU -> FC (R -> C) + sigmoid.
// Reference op site: timm SqueezeExcite inside the extractor invoked at model/few_shot_recognisers.py:114-117,143-146.
//
// The arithmetic is tiny (2 C R MACs per frame); what the launch costs is fetching the two weight matrices (up to
// 2 x 221 KB) from L2. se_gate_kernel streams them through a shared-memory ring with cp.async.bulk (one elected thread
// issues, mbarrier completion): kSeSlots chunks of whole rows are in flight from the first instruction, under the mean
// phase and under the arithmetic of the previous chunk, and no register holds a load in flight. (r02c: per-thread
// __ldg loads in unrolled batches were a chain of ~18 dependent L2 round trips per block, 42 us per wave at C = 1152.)
// A block handles F = ceil(B / SMs) frames (<= 12) so that the whole launch is ONE wave and the weights are fetched once
// per F frames; per-frame arithmetic order: means and the second FC as before (sequential over groups / over r), the
// first FC sums 4-channel groups per lane before the warp tree.
// se_gate_simple_kernel is the plain version for shapes the ring cannot take (C % 4 != 0, unaligned tensors).
// ------------------------------------------------------------------------------------------------
constexpr int kSeFrames = 8;

__global__ void __launch_bounds__(1024)
se_gate_simple_kernel(const float* __restrict__ partial, int tiles, float inv_hw, const float* __restrict__ w1,
                      const float* __restrict__ b1, const float* __restrict__ w2t, const float* __restrict__ b2,
                      float* __restrict__ gate, int B, int C, int R, int F) {
    extern __shared__ float s_se[];  // mean[F][C], hidden[F][R]
    float* s_mean = s_se;
    float* s_hid = s_se + F * C;
    const int f0 = blockIdx.x * F, nf = min(F, B - f0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < nf * C; i += blockDim.x) {
        const int f = i / C, c = i - f * C;
        const float* pb = partial + (int64_t)(f0 + f) * tiles * C + c;
        float s = 0.f;
        for (int t = 0; t < tiles; ++t) s += pb[(int64_t)t * C];
        s_mean[i] = s * inv_hw;
    }
    __syncthreads();
    for (int r = warp; r < R; r += n_warps) {
        float s[kSeFrames];
#pragma unroll
        for (int f = 0; f < kSeFrames; ++f) s[f] = 0.f;
        const float* wr = w1 + (int64_t)r * C;
#pragma unroll 12
        for (int c = lane; c < C; c += 32) {
            const float w = __ldg(wr + c);
#pragma unroll
            for (int f = 0; f < kSeFrames; ++f)
                if (f < nf) s[f] = fmaf(w, s_mean[f * C + c], s[f]);
        }
        const float bias = __ldg(b1 + r);
#pragma unroll
        for (int f = 0; f < kSeFrames; ++f) {
            if (f < nf) {
                const float t = warp_sum(s[f]);
                if (lane == 0) s_hid[f * R + r] = siluf_(t + bias);
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s[kSeFrames];
        const float bias = __ldg(b2 + c);
#pragma unroll
        for (int f = 0; f < kSeFrames; ++f) s[f] = bias;
#pragma unroll 8
        for (int r = 0; r < R; ++r) {
            const float w = __ldg(w2t + (int64_t)r * C + c);     // transposed [R][C]: coalesced, independent loads
#pragma unroll
            for (int f = 0; f < kSeFrames; ++f)
                if (f < nf) s[f] = fmaf(w, s_hid[f * R + r], s[f]);
        }
#pragma unroll
        for (int f = 0; f < kSeFrames; ++f)
            if (f < nf) gate[(int64_t)(f0 + f) * C + c] = sigmoidf_(s[f]);
    }
}

namespace seg {
constexpr int kMaxThreads = 1024, kSlots = 2, kChunkBytes = 72 * 1024, kMaxFrames = 12, kMaxSmem = 220 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {      // bounded: a protocol bug traps, never hangs
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (spin > 400000u) __trap();
    }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float dot4(const float4& w, const float4& m, float s) {
    return fmaf(w.w, m.w, fmaf(w.z, m.z, fmaf(w.y, m.y, fmaf(w.x, m.x, s))));
}

// shared memory: ring [kSlots][rows_per_chunk * C] | mean [FR][C] | hidden [FR][Rp] | mbarriers [kSlots]
// chunk k < nch: rows k * rows_per_chunk .. of w1 [R][C];  chunk nch + k: the same rows of w2t [R][C].
// FR = frames per block rounded up to a compile-time size: the rows of frames the block does not have are zeros, so that
// the hot loops carry no per-frame branch (with one, every frame's shared-memory load waited for the previous frame's FMAs).
// FR = 12 runs 640 threads (2 x 12 + 2 x 12 accumulators want ~100 registers), the smaller sizes up to 1024.
constexpr int max_threads(int fr) { return fr > 8 ? 640 : kMaxThreads; }

template <int FR>
__global__ void __launch_bounds__(max_threads(FR), 1)
se_gate_kernel(const float* __restrict__ partial, int tiles, float inv_hw, const float* __restrict__ w1,
               const float* __restrict__ b1, const float* __restrict__ w2t, const float* __restrict__ b2,
               float* __restrict__ gate, int B, int C, int R, int F, int rows_per_chunk) {
    extern __shared__ __align__(128) uint8_t s_raw[];
    const int chunk_floats = rows_per_chunk * C, Rp = (R + 3) & ~3, C4 = C >> 2, C2 = C >> 1;
    float* ring = reinterpret_cast<float*>(s_raw);
    float* s_mean = ring + kSlots * chunk_floats;
    float* s_hid = s_mean + FR * C;
    const uint32_t bars = smem_u32(s_hid + FR * Rp);
    const int nch = ceil_div(R, rows_per_chunk), NC = 2 * nch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int f0 = blockIdx.x * F, nf = min(F, B - f0);
    auto issue = [&](int k) {
        const int second = k >= nch, r0 = (k - second * nch) * rows_per_chunk, rows = min(rows_per_chunk, R - r0);
        const uint32_t bytes = (uint32_t)(rows * C) * 4u, bar = bars + 8u * (uint32_t)(k % kSlots);
        mbar_expect_tx(bar, bytes);
        bulk_load(smem_u32(ring + (k % kSlots) * chunk_floats), (second ? w2t : w1) + (int64_t)r0 * C, bytes, bar);
    };
    if (tid == 0) {
        for (int s = 0; s < kSlots; ++s) mbar_init(bars + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int k = 0; k < min(NC, kSlots); ++k) issue(k);
    }
    for (int i = tid; i < FR * Rp; i += nthreads) s_hid[i] = 0.f;      // the padding columns are read (times zero weights)
    for (int i = tid; i < FR * C4; i += nthreads) {                     // spatial means; zeros for the frames beyond nf
        const int f = i / C4, c4 = i - f * C4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < nf) {
            const float4* pb = reinterpret_cast<const float4*>(partial + (int64_t)(f0 + f) * tiles * C) + c4;
#pragma unroll 4
            for (int t = 0; t < tiles; ++t) {
                const float4 v = __ldg(pb + (int64_t)t * C4);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
        reinterpret_cast<float4*>(s_mean)[i] = make_float4(s.x * inv_hw, s.y * inv_hw, s.z * inv_hw, s.w * inv_hw);
    }
    float acc[FR][2];                         // second FC: this thread's channel pair (tid < C / 2); live from chunk nch on
    __syncthreads();
    for (int k = 0; k < NC; ++k) {
        const int slot = k % kSlots, second = k >= nch;
        const int r0 = (k - second * nch) * rows_per_chunk, rows = min(rows_per_chunk, R - r0);
        mbar_wait(bars + 8u * (uint32_t)slot, (uint32_t)(k / kSlots) & 1u);
        const float* wch = ring + slot * chunk_floats;
        if (k == nch) {
            const float2 bias = tid < C2 ? __ldg(reinterpret_cast<const float2*>(b2) + tid) : make_float2(0.f, 0.f);
#pragma unroll
            for (int f = 0; f < FR; ++f) { acc[f][0] = bias.x; acc[f][1] = bias.y; }
        }
        if (!second) {
            // hidden[f][r] = silu(b1[r] + w1[r] . mean[f]): a warp takes TWO rows, so that every mean quad it reads from shared
            // memory (the traffic that bounds this phase: 4 wavefronts per 128-bit warp load) feeds 8 FMAs instead of 4
            for (int rr = 2 * warp; rr < rows; rr += 2 * nwarps) {
                const bool two = rr + 1 < rows;
                float sa[FR], sb[FR];
#pragma unroll
                for (int f = 0; f < FR; ++f) { sa[f] = 0.f; sb[f] = 0.f; }
                const float4* wa = reinterpret_cast<const float4*>(wch + rr * C);
                const float4* wb = reinterpret_cast<const float4*>(wch + (two ? rr + 1 : rr) * C);
                for (int c4 = lane; c4 < C4; c4 += 32) {
                    const float4 va = wa[c4], vb = wb[c4];
#pragma unroll
                    for (int f = 0; f < FR; ++f) {
                        const float4 m = reinterpret_cast<const float4*>(s_mean + f * C)[c4];
                        sa[f] = dot4(va, m, sa[f]);
                        sb[f] = dot4(vb, m, sb[f]);
                    }
                }
                float mine_a = 0.f, mine_b = 0.f;      // lane f keeps frame f's totals: one SiLU per lane, not FR in sequence
#pragma unroll
                for (int f = 0; f < FR; ++f) {
                    const float ta = warp_sum(sa[f]), tb = warp_sum(sb[f]);
                    if (lane == f) { mine_a = ta; mine_b = tb; }
                }
                if (lane < FR) {
                    s_hid[lane * Rp + r0 + rr] = siluf_(mine_a + __ldg(b1 + r0 + rr));
                    if (two) s_hid[lane * Rp + r0 + rr + 1] = siluf_(mine_b + __ldg(b1 + r0 + rr + 1));
                }
            }
        } else if (tid < C2) {                // gate[f][c] += w2t[r][c] hidden[f][r], sequential over r
            const float2* wc = reinterpret_cast<const float2*>(wch) + tid;
            for (int rr = 0; rr < rows; rr += 4) {          // r0 and rows_per_chunk are multiples of 4: aligned float4 of hidden
                float2 w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = rr + j < rows ? wc[(rr + j) * C2] : make_float2(0.f, 0.f);
#pragma unroll
                for (int f = 0; f < FR; ++f) {
                    const float4 h = *reinterpret_cast<const float4*>(s_hid + f * Rp + r0 + rr);
                    acc[f][0] = fmaf(w[3].x, h.w, fmaf(w[2].x, h.z, fmaf(w[1].x, h.y, fmaf(w[0].x, h.x, acc[f][0]))));
                    acc[f][1] = fmaf(w[3].y, h.w, fmaf(w[2].y, h.z, fmaf(w[1].y, h.y, fmaf(w[0].y, h.x, acc[f][1]))));
                }
            }
        }
        __syncthreads();                      // the slot is drained (and, after the last w1 chunk, hidden is complete)
        if (tid == 0 && k + kSlots < NC) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(k + kSlots);
        }
    }
    if (tid < C2) {
#pragma unroll
        for (int f = 0; f < FR; ++f)
            if (f < nf)
                *reinterpret_cast<float2*>(gate + (int64_t)(f0 + f) * C + 2 * tid) = make_float2(sigmoidf_(acc[f][0]), sigmoidf_(acc[f][1]));
    }
}
}  // namespace seg

static int g_se_ring = 1;      // dev A/B switch (orbit_set_global_option "se_ring")
void set_se_ring(int on) { g_se_ring = on; }
int get_se_ring() { return g_se_ring; }

int launch_se_gate(const float* partial, int tiles, int hw, const float* w1, const float* b1, const float* w2t,
                   const float* b2, float* gate, int B, int C, int R, cudaStream_t st) {
    if (B <= 0) return ORBIT_OK;
    if (g_se_ring && C % 4 == 0 && C <= 2 * seg::kMaxThreads && R >= 1 && aligned16(partial) && aligned16(w1) && aligned16(w2t) &&
        aligned16(b2) && aligned16(gate)) {
        static int sms = 0;
        if (!sms) {
            int dev = 0;
            ORBIT_CUDA(cudaGetDevice(&dev));
            ORBIT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        }
        const int Rp = (R + 3) / 4 * 4;
        int rows_per_chunk = std::max(4, seg::kChunkBytes / (C * 4) / 4 * 4);
        if (rows_per_chunk >= R) rows_per_chunk = Rp;                      // one chunk per matrix
        int F = std::min(seg::kMaxFrames, std::max(1, ceil_div(B, sms)));            // one wave of blocks
        if (F > 8 && C > 2 * seg::max_threads(12)) F = 8;                            // the 12-frame instance maps C / 2 <= 640 threads
        const int FR = F <= 2 ? 2 : (F <= 4 ? 4 : (F <= 8 ? 8 : 12));
        const size_t smem = sizeof(float) * ((size_t)seg::kSlots * rows_per_chunk * C + (size_t)FR * (C + Rp)) + 8 * seg::kSlots;
        if (smem <= (size_t)seg::kMaxSmem) {
            const float inv_hw = 1.0f / (float)hw;
            const int threads = C <= 256 ? 256 : seg::max_threads(FR);
#define ORBIT_SEG(N) { ORBIT_CUDA(cudaFuncSetAttribute(seg::se_gate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                       seg::se_gate_kernel<N><<<ceil_div(B, F), threads, smem, st>>>(partial, tiles, inv_hw, w1, b1, w2t, b2, gate, B, C, R, F, rows_per_chunk); }
            if (FR == 2) ORBIT_SEG(2) else if (FR == 4) ORBIT_SEG(4) else if (FR == 8) ORBIT_SEG(8) else ORBIT_SEG(12)
#undef ORBIT_SEG
            ORBIT_RETURN_IF_LAUNCH_FAILED();
            return ORBIT_OK;
        }
    }
    int F = kSeFrames;
    while (F > 1 && sizeof(float) * (size_t)F * (C + R) > 44 * 1024) F >>= 1;   // stay inside the default 48 KB
    const size_t smem = sizeof(float) * (size_t)F * (C + R);
    se_gate_simple_kernel<<<ceil_div(B, F), C >= 256 ? 1024 : 256, smem, st>>>(partial, tiles, 1.0f / (float)hw, w1, b1, w2t, b2, gate, B, C, R, F);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// ------------------------------------------------------------------------------------------------
// Pointwise conv as an fp32 FFMA GEMM: out[M,N] = epi(A[M,K] . W[N,K]^T). 128 x BN x 16 tiles, 256 threads,
// 8 x (BN/16) register micro-tiles, register-prefetch double buffering. The A loader applies the
// SE gate (per frame, per input channel) so the gated tensor is never materialised.
// ------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(256)
pw_ffma_kernel(const float* __restrict__ A, const float* __restrict__ Wt, const float* __restrict__ scale,
               const float* __restrict__ shift, const float* __restrict__ gate, const float* __restrict__ residual,
               float* __restrict__ out, int M, int N, int K, int rows_per_frame, int act) {
    constexpr int BM = 128, BK = 16, TN = BN / 16, LDA = BM + 4, LDB = BN + 4;
    constexpr int B_LOADS = (BN * 4 + 255) / 256;  // float4 loads per thread for the W tile
    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][LDB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t block_m = (int64_t)blockIdx.x * BM;
    const int block_n = blockIdx.y * BN;

    const int a_q = tid & 3;
    int a_row[2];
    const float* a_ptr[2];
    const float* g_ptr[2];
    bool a_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        a_row[i] = (tid >> 2) + i * 64;
        const int64_t gm = block_m + a_row[i];
        a_ok[i] = gm < M;
        a_ptr[i] = A + (a_ok[i] ? gm : 0) * K + a_q * 4;
        g_ptr[i] = gate ? gate + ((a_ok[i] ? gm : 0) / rows_per_frame) * K + a_q * 4 : nullptr;
    }
    int b_row[B_LOADS];
    const float* b_ptr[B_LOADS];
    bool b_ok[B_LOADS];
#pragma unroll
    for (int i = 0; i < B_LOADS; ++i) {
        const int e = tid + i * 256;
        b_row[i] = e >> 2;
        const int gn = block_n + b_row[i];
        b_ok[i] = (e < BN * 4) && gn < N;
        b_ptr[i] = Wt + (int64_t)(b_ok[i] ? gn : 0) * K + a_q * 4;
    }

    float4 ra[2], rb[B_LOADS];
    auto load_tiles = [&](int k0) {
        const bool kin = k0 + a_q * 4 < K;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_ok[i] && kin) {
                ra[i] = ldg4(a_ptr[i] + k0);
                if (gate) { const float4 g = ldg4(g_ptr[i] + k0); ra[i].x *= g.x; ra[i].y *= g.y; ra[i].z *= g.z; ra[i].w *= g.w; }
            }
        }
#pragma unroll
        for (int i = 0; i < B_LOADS; ++i) {
            rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b_ok[i] && kin) rb[i] = ldg4(b_ptr[i] + k0);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            As[buf][a_q * 4 + 0][a_row[i]] = ra[i].x; As[buf][a_q * 4 + 1][a_row[i]] = ra[i].y;
            As[buf][a_q * 4 + 2][a_row[i]] = ra[i].z; As[buf][a_q * 4 + 3][a_row[i]] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < B_LOADS; ++i) {
            if (tid + i * 256 < BN * 4) {
                Bs[buf][a_q * 4 + 0][b_row[i]] = rb[i].x; Bs[buf][a_q * 4 + 1][b_row[i]] = rb[i].y;
                Bs[buf][a_q * 4 + 2][b_row[i]] = rb[i].z; Bs[buf][a_q * 4 + 3][b_row[i]] = rb[i].w;
            }
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int n_k = ceil_div(K, BK);
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < n_k; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_k) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], bb[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if constexpr (TN == 8) {
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
                bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
            } else if constexpr (TN == 4) {
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
            } else if constexpr (TN == 2) {
                const float2 b0 = *reinterpret_cast<const float2*>(&Bs[buf][k][tx * 2]);
                bb[0] = b0.x; bb[1] = b0.y;
            } else {
                bb[0] = Bs[buf][k][tx];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (kt + 1 < n_k) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue: folded BN/FiLM scale-shift, activation, residual
    constexpr int NG = TN == 8 ? 2 : 1;       // column groups per thread
    constexpr int GW = TN == 8 ? 4 : TN;      // group width
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const int col0 = block_n + (TN == 8 ? g * 64 + tx * 4 : tx * TN);
        if (col0 >= N) continue;
        float sc[GW], sh[GW];
#pragma unroll
        for (int j = 0; j < GW; ++j) { sc[j] = __ldg(scale + col0 + j); sh[j] = __ldg(shift + col0 + j); }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t row = block_m + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            if (row >= M) continue;
            float v[GW];
#pragma unroll
            for (int j = 0; j < GW; ++j) v[j] = apply_act(fmaf(acc[i][g * 4 + j], sc[j], sh[j]), act < 16 ? act : ACT_NONE);
            float* op = out + row * N + col0;
            if (residual) {
                const float* rp = residual + row * N + col0;
#pragma unroll
                for (int j = 0; j < GW; ++j) v[j] += __ldg(rp + j);
            }
            if (act >= 16) {
#pragma unroll
                for (int j = 0; j < GW; ++j) v[j] = apply_act(v[j], act - 16);
            }
            if constexpr (GW == 4) *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
            else if constexpr (GW == 2) *reinterpret_cast<float2*>(op) = make_float2(v[0], v[1]);
            else op[0] = v[0];
        }
    }
}

int launch_pointwise_ffma(const float* A, const float* Wt, const float* scale, const float* shift, const float* gate,
                          const float* residual, float* out, int M, int N, int K, int rows_per_frame, int act,
                          cudaStream_t st) {
    if (K % 4 || N % 4) return ORBIT_ERR_UNSUPPORTED;
    int BN;
    if (N <= 16) BN = 16;
    else if (N <= 32) BN = 32;
    else BN = (ceil_div(N, 64) * 64 < ceil_div(N, 128) * 128) ? 64 : 128;
    dim3 grid((unsigned)ceil_div(M, 128), ceil_div(N, BN));
#define ORBIT_PW_CASE(BNN)                                                                                             \
    if (BN == BNN) pw_ffma_kernel<BNN><<<grid, 256, 0, st>>>(A, Wt, scale, shift, gate, residual, out, M, N, K,        \
                                                            rows_per_frame, act);
    ORBIT_PW_CASE(16) ORBIT_PW_CASE(32) ORBIT_PW_CASE(64) ORBIT_PW_CASE(128)
#undef ORBIT_PW_CASE
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// ------------------------------------------------------------------------------------------------
// 3x3 convolution of the set encoder (reference model/set_encoders.py:91-120) = im2col + the pointwise GEMM
// (tensor cores) + 2x2 max pool. The im2col matrix is written once and read once; at the CNAPs support-set sizes
// (<= 150 frames) that is a few GB of traffic, far below the GEMM time.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ x, float* __restrict__ col, int B, int H, int W, int C, int k, int stride, int pad_t,
              int pad_l, int Ho, int Wo, int Kpad, int nchw) {
    const int kk = k * k;
    const int64_t total = (int64_t)B * Ho * Wo * Kpad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = i / Kpad;                       // the only 64-bit division; the rest is 32-bit
        const int kc = (int)(i - pix * Kpad);
        const int hw = Ho * Wo, b = (int)(pix / hw), p = (int)(pix - (int64_t)b * hw), oy = p / Wo, ox = p - oy * Wo;
        float v = 0.f;
        if (kc < kk * C) {
            const int c = nchw ? kc / kk : kc % C, tap = nchw ? kc % kk : kc / C;
            const int iy = oy * stride - pad_t + tap / k, ix = ox * stride - pad_l + tap % k;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W)
                v = nchw ? __ldg(x + (((int64_t)b * C + c) * H + iy) * W + ix) : __ldg(x + (((int64_t)b * H + iy) * W + ix) * C + c);
        }
        col[i] = v;
    }
}
// NHWC input with C % 4 == 0 (every conv but the first): one 128-bit load and store per thread, consecutive threads on
// consecutive channel quads of one tap (contiguous in x and in col), one 64-bit division per thread. The scalar kernel
// above ran 1 ms per launch on the ResNet-18 plan (40 of a 53 ms CNAPs episode); this one moves the same bytes at HBM speed.
__global__ void __launch_bounds__(256)
im2col_nhwc4_kernel(const float4* __restrict__ x, float4* __restrict__ col, int64_t num_pix, int H, int W, int C4, int k, int stride,
                    int pad_t, int pad_l, int Ho, int Wo) {
    // one block per output pixel (grid-stride): the pixel is decomposed once, a thread then owns elements r, r + 256, ... of
    // the pixel's [k*k][C/4] row: one small 32-bit division per 128-bit element. (A variant that walks along an output row
    // with the (ky, kx, c4) decomposition hoisted into registers measured slower.)
    const int kk = k * k, row4 = kk * C4, hw = Ho * Wo;
    for (int64_t pix = blockIdx.x; pix < num_pix; pix += gridDim.x) {
        const int b = (int)(pix / hw), p = (int)(pix - (int64_t)b * hw), oy = p / Wo, ox = p - oy * Wo;
        const int iy0 = oy * stride - pad_t, ix0 = ox * stride - pad_l;
        const float4* xb = x + (int64_t)b * H * W * C4;
        float4* crow = col + pix * row4;
        for (int r = threadIdx.x; r < row4; r += blockDim.x) {
            const int tap = r / C4, c4 = r - tap * C4, ky = tap / k, iy = iy0 + ky, ix = ix0 + (tap - ky * k);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xb + ((int64_t)iy * W + ix) * C4 + c4);
            crow[r] = v;
        }
    }
}

// NCHW input (the first convolution of ResNet / EfficientNet-V2 / the set encoder): a thread copies the k contiguous
// floats of one (channel, kernel row) of one output pixel; col index = c*k*k + ky*k + kx (torch weight order).
__global__ void __launch_bounds__(256)
im2col_nchw_rows_kernel(const float* __restrict__ x, float* __restrict__ col, int64_t num_pix, int H, int W, int C, int k, int stride,
                        int pad_t, int pad_l, int Ho, int Wo, int Kpad) {
    const int units = C * k, hw = Ho * Wo, per_block = blockDim.x / 32;       // one warp per pixel, lanes over (c, ky)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int64_t pix = (int64_t)blockIdx.x * per_block + wib; pix < num_pix; pix += (int64_t)gridDim.x * per_block) {
        const int b = (int)(pix / hw), p = (int)(pix - (int64_t)b * hw), oy = p / Wo, ox = p - oy * Wo;
        const int iy0 = oy * stride - pad_t, ix0 = ox * stride - pad_l;
        float* crow = col + pix * Kpad;
        for (int u = lane; u < units; u += 32) {
            const int c = u / k, ky = u - c * k, iy = iy0 + ky;
            const float* xr = x + (((int64_t)b * C + c) * H + iy) * W;
            float* dst = crow + u * k;
            const bool rok = iy >= 0 && iy < H;
            for (int kx = 0; kx < k; ++kx) {
                const int ix = ix0 + kx;
                dst[kx] = (rok && ix >= 0 && ix < W) ? __ldg(xr + ix) : 0.f;
            }
        }
        for (int z = C * k * k + lane; z < Kpad; z += 32) crow[z] = 0.f;       // padding columns
    }
}

int launch_im2col(const float* x, float* col, int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l, int Ho,
                  int Wo, int Kpad, int nchw, cudaStream_t st) {
    const int64_t total = (int64_t)B * Ho * Wo * Kpad;
    if (total == 0) return ORBIT_OK;
    if (nchw) {
        const int64_t num_pix = (int64_t)B * Ho * Wo;
        im2col_nchw_rows_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(num_pix, 8), 148 * 32), 256, 0, st>>>(
            x, col, num_pix, H, W, C, k, stride, pad_t, pad_l, Ho, Wo, Kpad);
        ORBIT_RETURN_IF_LAUNCH_FAILED();
        return ORBIT_OK;
    }
    if (!nchw && C % 4 == 0 && Kpad == k * k * C && aligned16(x) && aligned16(col)) {
        const int64_t num_pix = (int64_t)B * Ho * Wo;
        const int row4 = k * k * C / 4, threads = std::min(256, (row4 + 31) / 32 * 32);
        im2col_nhwc4_kernel<<<(unsigned)std::min<int64_t>(num_pix, 148 * 64), threads, 0, st>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(col), num_pix, H, W, C / 4, k, stride, pad_t, pad_l, Ho, Wo);
        ORBIT_RETURN_IF_LAUNCH_FAILED();
        return ORBIT_OK;
    }
    im2col_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 148 * 32), 256, 0, st>>>(x, col, B, H, W, C, k, stride, pad_t, pad_l, Ho, Wo, Kpad, nchw);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

__global__ void conv_weight_relayout_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int kk, int Kpad, int nchw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Cout * Kpad) return;
    const int co = i / Kpad, kc = i % Kpad;
    float v = 0.f;
    if (kc < kk * Cin) {
        const int c = nchw ? kc / kk : kc % Cin, tap = nchw ? kc % kk : kc / Cin;
        v = w[((int64_t)co * Cin + c) * kk + tap];
    }
    out[i] = v;
}
int launch_conv_weight_relayout(const float* w, float* out, int Cout, int Cin, int kk, int Kpad, int nchw, cudaStream_t st) {
    conv_weight_relayout_kernel<<<ceil_div(Cout * Kpad, 256), 256, 0, st>>>(w, out, Cout, Cin, kk, Kpad, nchw);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// grid (B * Ho output rows, column chunks); a thread = one output pixel x 4 channels. The row index comes from blockIdx (one
// 32-bit division per thread); the first version derived (b, oy, ox, q) from a flat 64-bit index with three 64-bit divisions
// per output (the eight pools of an S3 episode, 4.0 GB of traffic: 1.16 -> 1.11 ms; what remains is the 9-fold re-read of the 3x3 pool).
__global__ void __launch_bounds__(256)
maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo) {
    const int C4 = C >> 2;
    const int i = blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= Wo * C4) return;
    const int ox = i / C4, q = i - ox * C4;
    const int b = blockIdx.x / Ho, oy = blockIdx.x - b * Ho;
    const float* xb = x + (int64_t)b * H * W * C + q * 4;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int ky = 0; ky < k; ++ky) {
        const int iy = oy * stride - pad + ky;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < k; ++kx) {
            const int ix = ox * stride - pad + kx;
            if (ix < 0 || ix >= W) continue;
            const float4 v = ldg4(xb + ((int64_t)iy * W + ix) * C);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    *reinterpret_cast<float4*>(y + (((int64_t)b * Ho + oy) * Wo + ox) * C + q * 4) = m;
}
int launch_maxpool(const float* x, float* y, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                   cudaStream_t st) {
    if (C % 4) return ORBIT_ERR_UNSUPPORTED;
    const int64_t rows = (int64_t)B * Ho, cols = (int64_t)Wo * (C / 4);
    if (rows == 0 || cols == 0) return ORBIT_OK;
    if (rows > 0x7fffffffLL || ceil_div64(cols, 256) > 65535) return ORBIT_ERR_UNSUPPORTED;
    maxpool_kernel<<<dim3((unsigned)rows, (unsigned)ceil_div64(cols, 256)), 256, 0, st>>>(x, y, H, W, C, k, stride, pad, Ho, Wo);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// ------------------------------------------------------------------------------------------------
// Global average pool over the spatial positions: x [B,HW,C] -> y [B,C]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spatial_mean_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int HW, int C) {
    const int C4 = C >> 2;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * C4) return;
    const int b = (int)(i / C4), q = (int)(i % C4);
    const float* p = x + (int64_t)b * HW * C + q * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < HW; ++r) add4(s, ldg4_stream(p + (int64_t)r * C));
    const float d = (float)HW;
    *reinterpret_cast<float4*>(y + (int64_t)b * C + q * 4) = make_float4(s.x / d, s.y / d, s.z / d, s.w / d);
}

// The same mean with the rows of a frame split over the 8 warps of a block (fixed-order reduction: deterministic): for large
// HW and few frames (training passes keep the squeeze means of every depthwise layer) the one-thread-per-column kernel above
// is a serial chain of HW loads on a handful of warps. grid (B, ceil(C / 128)); lane = 4 channels, warp = a slice of the rows.
__global__ void __launch_bounds__(256)
spatial_mean_split_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int C) {
    __shared__ float4 s_part[8][32];
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5, b = blockIdx.x;
    const int c0 = blockIdx.y * 128 + lane * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 < C) {
        const float* p = x + (int64_t)b * HW * C + c0;
        const int per = (HW + 7) >> 3, r0 = rg * per, r1 = min(HW, r0 + per);
        for (int r = r0; r < r1; ++r) add4(s, ldg4_stream(p + (int64_t)r * C));
    }
    s_part[rg][lane] = s;
    __syncthreads();
    if (rg == 0 && c0 < C) {
        float4 t = s_part[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) add4(t, s_part[k][lane]);
        const float d = (float)HW;
        *reinterpret_cast<float4*>(y + (int64_t)b * C + c0) = make_float4(t.x / d, t.y / d, t.z / d, t.w / d);
    }
}

int launch_spatial_mean(const float* x, float* y, int B, int HW, int C, cudaStream_t st) {
    if (C % 4) return ORBIT_ERR_UNSUPPORTED;
    if (B <= 0) return ORBIT_OK;
    // enough columns to fill the GPU with the simple kernel (the 7x7x1280 pool of an inference pass), else split the rows
    if ((int64_t)B * (C / 4) >= 148 * 256 * 4 || HW < 64) {
        spatial_mean_kernel<<<(unsigned)ceil_div64((int64_t)B * (C / 4), 256), 256, 0, st>>>(x, y, B, HW, C);
    } else {
        spatial_mean_split_kernel<<<dim3(B, ceil_div(C, 128)), 256, 0, st>>>(x, y, HW, C);
    }
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

}  // namespace orbit
