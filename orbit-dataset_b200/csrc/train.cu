// Backward kernels of the EfficientNet-B0 extractor for FiLM fine-tuning (first slice).
//
// Reference: MultiStepFewShotRecogniser.personalise with adapt_features=True (model/few_shot_recognisers.py:196-198,
// 207-246): the FiLM parameters = the affine weight/bias of the tagged BatchNorms (model/film.py:38-66,76-79) and a new
// linear head are trained with gradient steps THROUGH the frozen extractor, BatchNorm in eval mode (running statistics,
// few_shot_recognisers.py:176-183). What autograd does for the reference is spelled out here:
//   BN (eval) + SiLU:   z = scale c + shift (scale = gamma rstd, shift = beta - mean scale),  y = z sigmoid(z)
//                       dz = dy * sigmoid(z) (1 + z (1 - sigmoid(z))),  dc = dz scale,
//                       dgamma = sum dz (c - mean) rstd,  dbeta = sum dz
//   1x1 conv:           dx = dc W          -> the tcgen05 GEMM on the transposed weights (gemm_tcgen05.cu)
//   depthwise conv:     dx[iy,ix] = sum_{ky,kx} dc[(iy + pad - ky)/s, (ix + pad - kx)/s] w[ky,kx]
//   squeeze-excite:     ga = a gate(mean a):  da = dga gate + dmean / HW,  dgate = sum_hw dga a,  dmean = FC^T(dgate ...)
//   pooling / head:     linear head + cross entropy (utils/optim.py:8-9) in two small kernels
// These kernels are written for correctness and coalesced access (they run a few dozen frames per grad step), not tuned.
#include "convnet.cuh"

namespace orbit {

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// y[m,c] = act(scale[c] x[m,c] + shift[c])
__global__ void __launch_bounds__(256)
bn_act_forward_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                      float* __restrict__ y, int64_t M, int C, int act) {
    const int c4 = C >> 2;
    const int64_t total = M * c4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        const float4 v = ldg4(x + 4 * i), sc = ldg4(scale + c), sh = ldg4(shift + c);
        float4 z = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
        if (act == ACT_SILU) { z.x *= sigmoid_acc(z.x); z.y *= sigmoid_acc(z.y); z.z *= sigmoid_acc(z.z); z.w *= sigmoid_acc(z.w); }
        *reinterpret_cast<float4*>(y + 4 * i) = z;
    }
}

int launch_bn_act_forward(const float* x, const float* scale, const float* shift, float* y, int64_t M, int C, int act, cudaStream_t st) {
    if (C % 4 || (act != ACT_NONE && act != ACT_SILU)) return ORBIT_ERR_UNSUPPORTED;
    if (M <= 0) return ORBIT_OK;
    const int blocks = (int)std::min<int64_t>(ceil_div64(M * (C / 4), 256), 148 * 16);
    bn_act_forward_kernel<<<blocks, 256, 0, st>>>(x, scale, shift, y, M, C, act);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// Backward of y = act(scale c + shift). dy source (mode): 0 = tensor dy[M,C]; 1 = dfeat[frame,C] / rows_per_frame (global
// average pool); 2 = dga[M,C] * gate[frame,C] + dmean[frame,C] / rows_per_frame (squeeze-excite input).
// Writes dc = dy act'(z) scale (when dc != nullptr) and, when `partial` != nullptr, per-(row block, channel) partial sums of
// dz and dz * xhat (xhat = (c - mean) rstd) at partial[blk][0][C] and partial[blk][1][C].
// grid (row blocks, ceil(C / 128)); block 256 = 8 row groups x 32 lanes, lane owns 4 consecutive channels.
constexpr int kBnBwdRows = 256;      // rows per block
__global__ void __launch_bounds__(256)
bn_act_backward_kernel(const float* __restrict__ cin, const float* __restrict__ dy, const float* __restrict__ aux0,
                       const float* __restrict__ aux1, const float* __restrict__ scale, const float* __restrict__ shift,
                       const float* __restrict__ mean, const float* __restrict__ var, float eps, float* __restrict__ dc,
                       float* __restrict__ partial, int64_t M, int C, int act, int mode, int rows_per_frame) {
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int c0 = blockIdx.y * 128 + lane * 4;
    const bool live = c0 < C;
    const int64_t r0 = (int64_t)blockIdx.x * kBnBwdRows, r1 = min(M, r0 + kBnBwdRows);
    float4 sc = make_float4(0, 0, 0, 0), sh = sc, mu = sc, rs = sc;
    if (live) {
        sc = ldg4(scale + c0); sh = ldg4(shift + c0);
        if (partial) {
            mu = ldg4(mean + c0);
            const float4 v = ldg4(var + c0);
            rs = make_float4(1.0f / sqrtf(v.x + eps), 1.0f / sqrtf(v.y + eps), 1.0f / sqrtf(v.z + eps), 1.0f / sqrtf(v.w + eps));
        }
    }
    float4 s1 = make_float4(0, 0, 0, 0), s2 = s1;
    const float inv_rows = 1.0f / (float)rows_per_frame;
    if (live) {
        for (int64_t r = r0 + rg; r < r1; r += 8) {
            const float4 c = ldg4(cin + r * C + c0);
            float4 g;
            if (mode == 0) {
                g = ldg4(dy + r * C + c0);
            } else {
                const int64_t f = r / rows_per_frame;
                if (mode == 1) {
                    g = ldg4(dy + f * C + c0);
                    g.x *= inv_rows; g.y *= inv_rows; g.z *= inv_rows; g.w *= inv_rows;
                } else {
                    const float4 d = ldg4(dy + r * C + c0), gt = ldg4(aux0 + f * C + c0), dm = ldg4(aux1 + f * C + c0);
                    g = make_float4(fmaf(d.x, gt.x, dm.x * inv_rows), fmaf(d.y, gt.y, dm.y * inv_rows),
                                    fmaf(d.z, gt.z, dm.z * inv_rows), fmaf(d.w, gt.w, dm.w * inv_rows));
                }
            }
            float4 dz = g;
            if (act == ACT_SILU) {
                const float zx = fmaf(c.x, sc.x, sh.x), zy = fmaf(c.y, sc.y, sh.y), zz = fmaf(c.z, sc.z, sh.z), zw = fmaf(c.w, sc.w, sh.w);
                const float px = sigmoid_acc(zx), py = sigmoid_acc(zy), pz = sigmoid_acc(zz), pw = sigmoid_acc(zw);
                dz.x *= px * (1.0f + zx * (1.0f - px)); dz.y *= py * (1.0f + zy * (1.0f - py));
                dz.z *= pz * (1.0f + zz * (1.0f - pz)); dz.w *= pw * (1.0f + zw * (1.0f - pw));
            }
            if (dc) *reinterpret_cast<float4*>(dc + r * C + c0) = make_float4(dz.x * sc.x, dz.y * sc.y, dz.z * sc.z, dz.w * sc.w);
            if (partial) {
                s1.x += dz.x; s1.y += dz.y; s1.z += dz.z; s1.w += dz.w;
                s2.x = fmaf(dz.x, (c.x - mu.x) * rs.x, s2.x); s2.y = fmaf(dz.y, (c.y - mu.y) * rs.y, s2.y);
                s2.z = fmaf(dz.z, (c.z - mu.z) * rs.z, s2.z); s2.w = fmaf(dz.w, (c.w - mu.w) * rs.w, s2.w);
            }
        }
    }
    if (partial) {
        __shared__ float4 s_a[8][32], s_b[8][32];
        s_a[rg][lane] = s1; s_b[rg][lane] = s2;
        __syncthreads();
        if (rg == 0 && live) {
            float4 a = s_a[0][lane], b2 = s_b[0][lane];
#pragma unroll
            for (int r = 1; r < 8; ++r) { add4(a, s_a[r][lane]); add4(b2, s_b[r][lane]); }     // fixed order: deterministic
            float* p = partial + (int64_t)blockIdx.x * 2 * C;
            *reinterpret_cast<float4*>(p + c0) = a;
            *reinterpret_cast<float4*>(p + C + c0) = b2;
        }
    }
}

// grad_gamma[c] += sum_blk partial[blk][1][c];  grad_beta[c] += sum_blk partial[blk][0][c]   (fixed order)
__global__ void bn_param_grad_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ grad_gamma,
                                     float* __restrict__ grad_beta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < nblk; ++k) { a += partial[(int64_t)k * 2 * C + c]; b += partial[(int64_t)k * 2 * C + C + c]; }
    if (grad_beta) grad_beta[c] += a;
    if (grad_gamma) grad_gamma[c] += b;
}

int64_t bn_act_backward_partial_floats(int64_t M, int C) { return ceil_div64(M, kBnBwdRows) * 2 * C; }

int launch_bn_act_backward(const float* c, const float* dy, const float* aux0, const float* aux1, const float* scale,
                           const float* shift, const float* mean, const float* var, float eps, float* dc, float* partial,
                           float* grad_gamma, float* grad_beta, int64_t M, int C, int act, int mode, int rows_per_frame,
                           cudaStream_t st) {
    if (C % 4 || (act != ACT_NONE && act != ACT_SILU) || mode < 0 || mode > 2) return ORBIT_ERR_UNSUPPORTED;
    if (M <= 0) return ORBIT_OK;
    const bool want = grad_gamma || grad_beta;
    if (want && !partial) return ORBIT_ERR_ARG;
    dim3 grid((unsigned)ceil_div64(M, kBnBwdRows), ceil_div(C, 128));
    bn_act_backward_kernel<<<grid, 256, 0, st>>>(c, dy, aux0, aux1, scale, shift, mean, var, eps, dc, want ? partial : nullptr, M, C,
                                                act, mode, rows_per_frame);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    if (want) {
        bn_param_grad_kernel<<<ceil_div(C, 256), 256, 0, st>>>(partial, (int)grid.x, C, grad_gamma, grad_beta);
        ORBIT_RETURN_IF_LAUNCH_FAILED();
    }
    return ORBIT_OK;
}

// Depthwise conv data gradient: dx [B,H,W,C] from dy [B,Ho,Wo,C], taps wt [k*k][C], forward geometry (stride, pad_t, pad_l).
__global__ void __launch_bounds__(256)
dw_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ wt, float* __restrict__ dx, int B, int H, int W, int C,
                int Ho, int Wo, int k, int stride, int pad_t, int pad_l) {
    const int c4 = C >> 2;
    const int64_t total = (int64_t)B * H * W * c4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        int64_t p = i / c4;
        const int ix = (int)(p % W); p /= W;
        const int iy = (int)(p % H);
        const int b = (int)(p / H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ky = 0; ky < k; ++ky) {
            const int ty = iy + pad_t - ky;
            if (ty < 0 || ty % stride) continue;
            const int oy = ty / stride;
            if (oy >= Ho) continue;
            for (int kx = 0; kx < k; ++kx) {
                const int tx = ix + pad_l - kx;
                if (tx < 0 || tx % stride) continue;
                const int ox = tx / stride;
                if (ox >= Wo) continue;
                fma4(acc, ldg4(dy + (((int64_t)b * Ho + oy) * Wo + ox) * C + c), ldg4(wt + (int64_t)(ky * k + kx) * C + c));
            }
        }
        *reinterpret_cast<float4*>(dx + 4 * i) = acc;
    }
}

int launch_dw_dgrad(const float* dy, const float* wt, float* dx, int B, int H, int W, int C, int Ho, int Wo, int k, int stride,
                    int pad_t, int pad_l, cudaStream_t st) {
    if (C % 4) return ORBIT_ERR_UNSUPPORTED;
    if (B <= 0) return ORBIT_OK;
    const int blocks = (int)std::min<int64_t>(ceil_div64((int64_t)B * H * W * (C / 4), 256), 148 * 32);
    dw_dgrad_kernel<<<blocks, 256, 0, st>>>(dy, wt, dx, B, H, W, C, Ho, Wo, k, stride, pad_t, pad_l);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// dgate[b,c] = sum_hw dga[b,hw,c] a[b,hw,c].  grid (B, ceil(C/128)), block 256 = 8 row groups x 32 lanes
__global__ void __launch_bounds__(256)
se_dgate_kernel(const float* __restrict__ dga, const float* __restrict__ a, float* __restrict__ dgate, int HW, int C) {
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5, b = blockIdx.x;
    const int c0 = blockIdx.y * 128 + lane * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 < C)
        for (int r = rg; r < HW; r += 8) fma4(s, ldg4(dga + ((int64_t)b * HW + r) * C + c0), ldg4(a + ((int64_t)b * HW + r) * C + c0));
    __shared__ float4 s_s[8][32];
    s_s[rg][lane] = s;
    __syncthreads();
    if (rg == 0 && c0 < C) {
        float4 t = s_s[0][lane];
#pragma unroll
        for (int r = 1; r < 8; ++r) add4(t, s_s[r][lane]);
        *reinterpret_cast<float4*>(dgate + (int64_t)b * C + c0) = t;
    }
}

// Squeeze-excite FC backward, one block per frame: gate = sigmoid(W2 silu(W1 m + b1) + b2), given dgate -> dmean.
// w1 [R,C], b1 [R], w2 [C,R], b2 [C] (torch layouts).
__global__ void __launch_bounds__(256)
se_fc_backward_kernel(const float* __restrict__ m, const float* __restrict__ dgate, const float* __restrict__ w1,
                      const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                      float* __restrict__ dmean, int C, int R) {
    extern __shared__ float s_se[];   // v[R], dv[R], du[C]
    float* s_v = s_se; float* s_dv = s_se + R; float* s_du = s_se + 2 * R;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float* mb = m + (int64_t)b * C;
    for (int r = warp; r < R; r += nw) {                      // v = W1 m + b1
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(__ldg(w1 + (int64_t)r * C + c), __ldg(mb + c), s);
        s = warp_sum(s);
        if (lane == 0) s_v[r] = s + __ldg(b1 + r);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {       // u = W2 h + b2, du = dgate g (1 - g)
        float u = __ldg(b2 + c);
        for (int r = 0; r < R; ++r) { const float v = s_v[r]; u = fmaf(__ldg(w2 + (int64_t)c * R + r), v * sigmoid_acc(v), u); }
        const float gt = sigmoid_acc(u);
        s_du[c] = __ldg(dgate + (int64_t)b * C + c) * gt * (1.0f - gt);
    }
    __syncthreads();
    for (int r = warp; r < R; r += nw) {                      // dh = W2^T du, dv = dh silu'(v)
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(__ldg(w2 + (int64_t)c * R + r), s_du[c], s);
        s = warp_sum(s);
        if (lane == 0) { const float v = s_v[r], p = sigmoid_acc(v); s_dv[r] = s * p * (1.0f + v * (1.0f - p)); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {       // dmean = W1^T dv
        float s = 0.f;
        for (int r = 0; r < R; ++r) s = fmaf(__ldg(w1 + (int64_t)r * C + c), s_dv[r], s);
        dmean[(int64_t)b * C + c] = s;
    }
}

int launch_se_backward(const float* dga, const float* a, const float* m, const float* w1, const float* b1, const float* w2,
                       const float* b2, float* dgate, float* dmean, int B, int HW, int C, int R, cudaStream_t st) {
    if (C % 4) return ORBIT_ERR_UNSUPPORTED;
    if (B <= 0) return ORBIT_OK;
    se_dgate_kernel<<<dim3(B, ceil_div(C, 128)), 256, 0, st>>>(dga, a, dgate, HW, C);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    se_fc_backward_kernel<<<B, 256, sizeof(float) * (size_t)(2 * R + C), st>>>(m, dgate, w1, b1, w2, b2, dmean, C, R);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// out[n][k] = w[k][n]   (weights [K_out, N_in] -> [N_in, K_out] for the data-gradient GEMM)
__global__ void transpose_kernel(const float* __restrict__ w, float* __restrict__ out, int rows, int cols) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    out[(int64_t)c * rows + r] = w[i];
}
int launch_transpose(const float* w, float* out, int rows, int cols, cudaStream_t st) {
    transpose_kernel<<<(unsigned)ceil_div64((int64_t)rows * cols, 256), 256, 0, st>>>(w, out, rows, cols);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// ------------------------------------------------------------------------------------------------
// Linear head + cross entropy (reduction 'mean' over the batch, times `loss_scale` = batch_len / context_size as in
// few_shot_recognisers.py:241-243): logits = logit_scale (pooled W^T + b), pooled = mean over the clip's L frame features.
//   kernel A, one block per clip: softmax, dlogits = (p - onehot) loss_scale / batch, loss contribution,
//             dfeat[frame] = logit_scale (dlogits W) / L for the clip's L frames
//   kernel B: grad_w[c,d] += logit_scale sum_n dlogits[n,c] pooled[n,d],  grad_b[c] += logit_scale sum_n dlogits[n,c]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
linear_ce_clip_kernel(const float* __restrict__ feats, const int32_t* __restrict__ labels, const float* __restrict__ w,
                      const float* __restrict__ bias, int L, int D, int C, float logit_scale, float grad_scale,
                      float* __restrict__ pooled, float* __restrict__ dlogits, float* __restrict__ dfeat) {
    extern __shared__ float s_lin[];   // pooled[D], logits[C]
    float* s_p = s_lin; float* s_l = s_lin + D;
    const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < L; ++l) s += feats[((int64_t)n * L + l) * D + d];
        s /= (float)L;
        s_p[d] = s;
        pooled[(int64_t)n * D + d] = s;
    }
    __syncthreads();
    for (int c = warp; c < C; c += nw) {
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s = fmaf(s_p[d], __ldg(w + (int64_t)c * D + d), s);
        s = warp_sum(s);
        if (lane == 0) s_l[c] = logit_scale * (s + __ldg(bias + c));
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, s_l[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += expf(s_l[c] - mx);
        se = warp_sum(se);
        const int y = labels[n];
        for (int c = lane; c < C; c += 32) {
            const float p = expf(s_l[c] - mx) / se;
            const float g = (p - (c == y ? 1.0f : 0.0f)) * grad_scale;
            dlogits[(int64_t)n * C + c] = g;
            s_l[c] = g;
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(s_l[c], __ldg(w + (int64_t)c * D + d), s);
        s *= logit_scale / (float)L;
        for (int l = 0; l < L; ++l) dfeat[((int64_t)n * L + l) * D + d] = s;
    }
}

__global__ void __launch_bounds__(256)
linear_ce_param_kernel(const float* __restrict__ pooled, const float* __restrict__ dlogits, int N, int D, int C, float logit_scale,
                       float* __restrict__ grad_w, float* __restrict__ grad_b) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < (int64_t)C * D) {
        const int c = (int)(i / D), d = (int)(i % D);
        float s = 0.f;
        for (int n = 0; n < N; ++n) s = fmaf(dlogits[(int64_t)n * C + c], pooled[(int64_t)n * D + d], s);
        grad_w[i] += logit_scale * s;
    }
    if (i < C) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s += dlogits[(int64_t)n * C + i];
        grad_b[i] += logit_scale * s;
    }
}

}  // namespace orbit

using namespace orbit;

extern "C" int64_t orbit_linear_ce_scratch_floats(int num_clips, int feat_dim, int num_classes) {
    return (int64_t)num_clips * feat_dim + (int64_t)num_clips * num_classes;
}

extern "C" int orbit_linear_ce_backward(const float* frame_feats, const int32_t* labels, const float* weight, const float* bias,
                                        int num_clips, int clip_length, int feat_dim, int num_classes, float logit_scale,
                                        float loss_scale, float* grad_weight, float* grad_bias, float* grad_frame_feats,
                                        float* scratch, void* stream) {
    if (!frame_feats || !labels || !weight || !bias || !grad_weight || !grad_bias || !grad_frame_feats || !scratch) return ORBIT_ERR_ARG;
    if (num_clips <= 0 || clip_length <= 0 || feat_dim <= 0 || num_classes <= 0) return ORBIT_ERR_ARG;
    if (num_classes > 1024 || (size_t)(feat_dim + num_classes) * sizeof(float) > 96 * 1024) return ORBIT_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    float* pooled = scratch;
    float* dlogits = scratch + (int64_t)num_clips * feat_dim;
    const size_t smem = sizeof(float) * (size_t)(feat_dim + num_classes);
    if (smem > 48 * 1024) ORBIT_CUDA(cudaFuncSetAttribute(linear_ce_clip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    linear_ce_clip_kernel<<<num_clips, 256, smem, st>>>(frame_feats, labels, weight, bias, clip_length, feat_dim, num_classes, logit_scale,
                                                       loss_scale / (float)num_clips, pooled, dlogits, grad_frame_feats);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    linear_ce_param_kernel<<<(unsigned)ceil_div64((int64_t)num_classes * feat_dim, 256), 256, 0, st>>>(
        pooled, dlogits, num_clips, feat_dim, num_classes, logit_scale, grad_weight, grad_bias);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}
