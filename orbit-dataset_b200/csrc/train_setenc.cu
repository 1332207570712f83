// Backward kernels of the set encoder, the FiLM generator and the classifier heads' query path: what CNAPs-style
// meta-training needs on top of the extractor backward of train.cu.
//
// Reference: single-step-learner.py:196-243 (train_task / train_task_with_lite). With a frozen extractor the loss reaches
//   * the FiLM generator (model/feature_adapters.py:36-78) through the FiLM parameters used on the QUERY path, and
//   * the set encoder (model/set_encoders.py:81-120: 5 x [conv3x3 + BatchNorm (eval) + ReLU + maxpool 2x2] + average pool)
//     through the task embedding;
// the support -> head path carries no gradient because configure() wraps the head weights in nn.Parameter
// (classifier_heads.py:179-180,261-263). What autograd does for the reference, spelled out:
//   avgpool:            dp[b,y,x,c] = dfeat[b,c] / (Hp Wp)
//   maxpool + ReLU:     the gradient of a 2x2 window goes to its first maximum, and only if that maximum is > 0
//   BN (eval) of conv + bias:  z = scale c + shift;  dgamma = sum dz xhat,  dbeta = sum dz,  dbias = scale sum dz,  dc = dz scale
//   conv3x3:            dW[co,ci,ky,kx] = sum dc[b,y,x,co] in[b,y+ky-1,x+kx-1,ci];   din = conv3x3(dc, W flipped / transposed)
// Written for correctness, deterministic reductions (fixed-order partial sums) and coalesced access; not tuned.
#include "convnet.cuh"

namespace orbit {

// ------------------------------------------------------------------------------------------------
// Forward: p[b,py,px,c] = max over the 2x2 window of relu(scale c + shift)      (floor mode: odd rows / columns dropped)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_relu_pool2_forward_kernel(const float* __restrict__ c, const float* __restrict__ scale, const float* __restrict__ shift,
                             float* __restrict__ p, int B, int H, int W, int C, int Hp, int Wp) {
    const int C4 = C >> 2;
    const int64_t total = (int64_t)B * Hp * Wp * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % C4) * 4;
        const int64_t pix = i / C4;
        const int px = (int)(pix % Wp), py = (int)((pix / Wp) % Hp), b = (int)(pix / ((int64_t)Wp * Hp));
        const float4 sc = ldg4(scale + q), sh = ldg4(shift + q);
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);            // relu: the maximum of the window is at least 0
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 v = ldg4(c + (((int64_t)b * H + 2 * py + (k >> 1)) * W + 2 * px + (k & 1)) * C + q);
            m.x = fmaxf(m.x, fmaf(v.x, sc.x, sh.x)); m.y = fmaxf(m.y, fmaf(v.y, sc.y, sh.y));
            m.z = fmaxf(m.z, fmaf(v.z, sc.z, sh.z)); m.w = fmaxf(m.w, fmaf(v.w, sc.w, sh.w));
        }
        *reinterpret_cast<float4*>(p + pix * C + q) = m;
    }
}

int launch_bn_relu_pool2_forward(const float* c, const float* scale, const float* shift, float* p, int B, int H, int W, int C,
                                 cudaStream_t st) {
    if (C % 4) return ORBIT_ERR_UNSUPPORTED;
    const int Hp = H / 2, Wp = W / 2;
    const int64_t total = (int64_t)B * Hp * Wp * (C / 4);
    if (total <= 0) return ORBIT_OK;
    bn_relu_pool2_forward_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 148 * 32), 256, 0, st>>>(c, scale, shift, p, B, H, W,
                                                                                                                  C, Hp, Wp);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of maxpool2x2(relu(scale c + shift)). The upstream gradient is dp [B,Hp,Wp,C] (mode 0) or dfeat [B,C] / (Hp Wp)
// (mode 1: the global average pool). Writes dc [B,H,W,C] for the 2 Hp x 2 Wp covered positions (the caller zeroes dc
// first when H or W is odd) and per-(block, channel) partial sums of dz and dz * xhat, xhat = (c + conv_bias - mean) rstd.
// block 256 = (256 / (C/4)) pooled positions x C/4 lanes; one block covers kPoolBwdRows pooled positions.
// ------------------------------------------------------------------------------------------------
constexpr int kPoolBwdRows = 128;
__global__ void __launch_bounds__(256)
pool_bn_relu_backward_kernel(const float* __restrict__ c, const float* __restrict__ dp, const float* __restrict__ scale,
                             const float* __restrict__ shift, const float* __restrict__ conv_bias, const float* __restrict__ mean,
                             const float* __restrict__ var, float eps, float* __restrict__ dc, float* __restrict__ partial, int B,
                             int H, int W, int C, int Hp, int Wp, int mode) {
    extern __shared__ float4 s_pool[];                 // [2][rows_per_pass][C/4]
    const int C4 = C >> 2;
    const int lane = threadIdx.x % C4, rg = threadIdx.x / C4, groups = 256 / C4;
    const int q = lane * 4;
    const float4 sc = ldg4(scale + q), sh = ldg4(shift + q), mu = ldg4(mean + q), vr = ldg4(var + q);
    const float4 cb = conv_bias ? ldg4(conv_bias + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 rs = make_float4(1.0f / sqrtf(vr.x + eps), 1.0f / sqrtf(vr.y + eps), 1.0f / sqrtf(vr.z + eps), 1.0f / sqrtf(vr.w + eps));
    const int64_t total = (int64_t)B * Hp * Wp;
    const int64_t r0 = (int64_t)blockIdx.x * kPoolBwdRows, r1 = min(total, r0 + kPoolBwdRows);
    const float inv = 1.0f / (float)(Hp * Wp);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    for (int64_t pix = r0 + rg; pix < r1; pix += groups) {
        const int px = (int)(pix % Wp), py = (int)((pix / Wp) % Hp), b = (int)(pix / ((int64_t)Wp * Hp));
        float4 g = mode == 0 ? ldg4(dp + pix * C + q) : ldg4(dp + (int64_t)b * C + q);
        if (mode) { g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv; }
        float4 v[4];
        float zx = 0.f, zy = 0.f, zz = 0.f, zw = 0.f;      // running maxima, start at 0: a maximum <= 0 gets no gradient
        int ix = -1, iy = -1, iz = -1, iw = -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = ldg4(c + (((int64_t)b * H + 2 * py + (k >> 1)) * W + 2 * px + (k & 1)) * C + q);
            const float a = fmaf(v[k].x, sc.x, sh.x), bb = fmaf(v[k].y, sc.y, sh.y), cc = fmaf(v[k].z, sc.z, sh.z), dd = fmaf(v[k].w, sc.w, sh.w);
            if (a > zx) { zx = a; ix = k; }                // strict: the FIRST maximum in scan order wins (torch max_pool2d)
            if (bb > zy) { zy = bb; iy = k; }
            if (cc > zz) { zz = cc; iz = k; }
            if (dd > zw) { zw = dd; iw = k; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 dz = make_float4(ix == k ? g.x : 0.f, iy == k ? g.y : 0.f, iz == k ? g.z : 0.f, iw == k ? g.w : 0.f);
            *reinterpret_cast<float4*>(dc + (((int64_t)b * H + 2 * py + (k >> 1)) * W + 2 * px + (k & 1)) * C + q) =
                make_float4(dz.x * sc.x, dz.y * sc.y, dz.z * sc.z, dz.w * sc.w);
            s1.x += dz.x; s1.y += dz.y; s1.z += dz.z; s1.w += dz.w;
            s2.x = fmaf(dz.x, (v[k].x + cb.x - mu.x) * rs.x, s2.x); s2.y = fmaf(dz.y, (v[k].y + cb.y - mu.y) * rs.y, s2.y);
            s2.z = fmaf(dz.z, (v[k].z + cb.z - mu.z) * rs.z, s2.z); s2.w = fmaf(dz.w, (v[k].w + cb.w - mu.w) * rs.w, s2.w);
        }
    }
    float4* s_a = s_pool;
    float4* s_b = s_pool + groups * C4;
    s_a[rg * C4 + lane] = s1; s_b[rg * C4 + lane] = s2;
    __syncthreads();
    if (rg == 0) {
        float4 a = s_a[lane], b2 = s_b[lane];
        for (int r = 1; r < groups; ++r) { add4(a, s_a[r * C4 + lane]); add4(b2, s_b[r * C4 + lane]); }   // fixed order
        float* pp = partial + (int64_t)blockIdx.x * 2 * C;
        *reinterpret_cast<float4*>(pp + q) = a;
        *reinterpret_cast<float4*>(pp + C + q) = b2;
    }
}

// grad_gamma += sum partial[.][1], grad_beta += sum partial[.][0], grad_conv_bias += scale * sum partial[.][0]
__global__ void conv_bn_param_grad_kernel(const float* __restrict__ partial, int nblk, int C, const float* __restrict__ scale,
                                          float* __restrict__ grad_gamma, float* __restrict__ grad_beta, float* __restrict__ grad_conv_bias) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < nblk; ++k) { a += partial[(int64_t)k * 2 * C + c]; b += partial[(int64_t)k * 2 * C + C + c]; }
    if (grad_beta) grad_beta[c] += a;
    if (grad_gamma) grad_gamma[c] += b;
    if (grad_conv_bias) grad_conv_bias[c] += a * scale[c];
}

int64_t pool_bn_relu_backward_partial_floats(int B, int H, int W, int C) {
    return ceil_div64((int64_t)B * (H / 2) * (W / 2), kPoolBwdRows) * 2 * C;
}

int launch_pool_bn_relu_backward(const float* c, const float* dp, const float* scale, const float* shift, const float* conv_bias,
                                 const float* mean, const float* var, float eps, float* dc, float* partial, float* grad_gamma,
                                 float* grad_beta, float* grad_conv_bias, int B, int H, int W, int C, int mode, cudaStream_t st) {
    if (C % 4 || C > 1024 || 256 % (C / 4) || mode < 0 || mode > 1) return ORBIT_ERR_UNSUPPORTED;
    if (!partial) return ORBIT_ERR_ARG;
    const int Hp = H / 2, Wp = W / 2;
    if (B <= 0 || Hp <= 0 || Wp <= 0) return ORBIT_OK;
    if ((H & 1) || (W & 1)) ORBIT_CUDA(cudaMemsetAsync(dc, 0, sizeof(float) * (size_t)B * H * W * C, st));
    const int nblk = (int)ceil_div64((int64_t)B * Hp * Wp, kPoolBwdRows);
    const size_t smem = sizeof(float4) * 2 * 256;
    pool_bn_relu_backward_kernel<<<nblk, 256, smem, st>>>(c, dp, scale, shift, conv_bias, mean, var, eps, dc, partial, B, H, W, C, Hp, Wp, mode);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    conv_bn_param_grad_kernel<<<ceil_div(C, 128), 128, 0, st>>>(partial, nblk, C, scale, grad_gamma, grad_beta, grad_conv_bias);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// ------------------------------------------------------------------------------------------------
// conv3x3 (pad 1, stride 1) weight gradient, NHWC input with Cin, Cout multiples of 64.
//   partial[chunk][ky][kx][co][ci] = sum over the chunk's row segments of dc[b,y,x,co] * in[b,y+ky-1,x+kx-1,ci]
// grid (chunks, 3 = ky, (Cin/64) (Cout/64)); block 256 = 16 (co groups of 4) x 16 (ci groups of 4); a row segment is 16
// consecutive pixels of one image row: the dc tile [16][64] and the input tile [18][64] (one halo pixel each side) are
// staged in shared memory and every thread accumulates a 4 x 4 x 3 (kx) register tile.
// ------------------------------------------------------------------------------------------------
constexpr int kWgSeg = 16;
__global__ void __launch_bounds__(256)
conv3_wgrad_kernel(const float* __restrict__ dc, const float* __restrict__ in, float* __restrict__ partial, int B, int H, int W,
                   int Cin, int Cout) {
    __shared__ __align__(16) float s_d[kWgSeg][64];
    __shared__ __align__(16) float s_x[kWgSeg + 2][64];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int ky = blockIdx.y;
    const int ci_blocks = Cin >> 6;
    const int ci0 = (blockIdx.z % ci_blocks) * 64, co0 = (blockIdx.z / ci_blocks) * 64;
    const int segs_per_row = ceil_div(W, kWgSeg);
    const int64_t nseg = (int64_t)B * H * segs_per_row;
    float acc[3][4][4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[k][i][j] = 0.f;
    for (int64_t s = blockIdx.x; s < nseg; s += gridDim.x) {
        const int x0 = (int)(s % segs_per_row) * kWgSeg;
        const int y = (int)((s / segs_per_row) % H), b = (int)(s / ((int64_t)segs_per_row * H));
        const int iy = y + ky - 1;
        if (iy < 0 || iy >= H) continue;                         // uniform per block: the whole input row is padding
        __syncthreads();
        {   // dc tile: 16 pixels x 16 float4; thread (ty = pixel, tx = float4)
            const int x = x0 + ty;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (x < W) v = ldg4(dc + (((int64_t)b * H + y) * W + x) * Cout + co0 + tx * 4);
            *reinterpret_cast<float4*>(&s_d[ty][tx * 4]) = v;
        }
        for (int t = threadIdx.x; t < (kWgSeg + 2) * 16; t += 256) {
            const int p = t >> 4, f = t & 15, x = x0 - 1 + p;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (x >= 0 && x < W) v = ldg4(in + (((int64_t)b * H + iy) * W + x) * Cin + ci0 + f * 4);
            *reinterpret_cast<float4*>(&s_x[p][f * 4]) = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int p = 0; p < kWgSeg; ++p) {
            const float4 d = *reinterpret_cast<const float4*>(&s_d[p][ty * 4]);
            const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 xv = *reinterpret_cast<const float4*>(&s_x[p + k][tx * 4]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[k][i][0] = fmaf(dv[i], xv.x, acc[k][i][0]); acc[k][i][1] = fmaf(dv[i], xv.y, acc[k][i][1]);
                    acc[k][i][2] = fmaf(dv[i], xv.z, acc[k][i][2]); acc[k][i][3] = fmaf(dv[i], xv.w, acc[k][i][3]);
                }
            }
        }
    }
    float* out = partial + ((int64_t)blockIdx.x * 9 + ky * 3) * Cout * Cin;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(out + ((int64_t)k * Cout + co0 + ty * 4 + i) * Cin + ci0 + tx * 4) =
                make_float4(acc[k][i][0], acc[k][i][1], acc[k][i][2], acc[k][i][3]);
}

// grad_w[co][ci][ky][kx] += sum_chunk partial[chunk][ky*3+kx][co][ci]      (fixed order)
__global__ void conv3_wgrad_reduce_kernel(const float* __restrict__ partial, int chunks, int Cin, int Cout, float* __restrict__ grad_w) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)9 * Cout * Cin;
    if (i >= n) return;
    const int ci = (int)(i % Cin), co = (int)((i / Cin) % Cout), tap = (int)(i / ((int64_t)Cin * Cout));
    float s = 0.f;
    for (int k = 0; k < chunks; ++k) s += partial[(int64_t)k * n + i];
    grad_w[((int64_t)co * Cin + ci) * 9 + tap] += s;
}

constexpr int kWgChunks = 96;
int launch_conv3_wgrad(const float* dc, const float* in, float* partial, int64_t partial_capacity, float* grad_w, int B, int H, int W,
                       int Cin, int Cout, cudaStream_t st) {
    if (Cin % 64 || Cout % 64) return ORBIT_ERR_UNSUPPORTED;
    if (B <= 0) return ORBIT_OK;
    const int64_t nseg = (int64_t)B * H * ceil_div(W, kWgSeg);
    const int chunks = (int)std::min<int64_t>(std::min<int64_t>(kWgChunks, nseg), partial_capacity / ((int64_t)9 * Cin * Cout));
    if (chunks < 1) return ORBIT_ERR_WORKSPACE;
    dim3 grid(chunks, 3, (Cin / 64) * (Cout / 64));
    conv3_wgrad_kernel<<<grid, 256, 0, st>>>(dc, in, partial, B, H, W, Cin, Cout);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    const int64_t n = (int64_t)9 * Cin * Cout;
    conv3_wgrad_reduce_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(partial, chunks, Cin, Cout, grad_w);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// ------------------------------------------------------------------------------------------------
// Weight gradient from an explicit im2col matrix (the first conv: 3 input channels, K = 27 padded to 28):
//   partial[blk][co][k] = sum over the block's rows of dc[m][co] * col[m][k];  grad_w[co][k] += sum_blk (k < K)
// block 256 = 64 output channels x 4 column groups; rows are staged 32 at a time in shared memory. Cout == 64.
// ------------------------------------------------------------------------------------------------
constexpr int kColRows = 32, kColMaxK = 64;
__global__ void __launch_bounds__(256)
col_wgrad_kernel(const float* __restrict__ dc, const float* __restrict__ col, float* __restrict__ partial, int64_t M, int Kpad,
                 int64_t rows_per_block) {
    __shared__ float s_d[kColRows][64];
    __shared__ float s_c[kColRows][kColMaxK];
    const int co = threadIdx.x & 63, kg = threadIdx.x >> 6;
    const int kper = Kpad >> 2;                                  // columns per group (Kpad is a multiple of 4)
    float acc[kColMaxK / 4];
#pragma unroll
    for (int j = 0; j < kColMaxK / 4; ++j) acc[j] = 0.f;
    const int64_t m0 = (int64_t)blockIdx.x * rows_per_block, m1 = min(M, m0 + rows_per_block);
    for (int64_t m = m0; m < m1; m += kColRows) {
        const int rows = (int)min((int64_t)kColRows, m1 - m);
        __syncthreads();
        for (int t = threadIdx.x; t < rows * 64; t += 256) s_d[t >> 6][t & 63] = dc[(m + (t >> 6)) * 64 + (t & 63)];
        for (int t = threadIdx.x; t < rows * Kpad; t += 256) s_c[t / Kpad][t % Kpad] = col[(m + t / Kpad) * Kpad + t % Kpad];
        __syncthreads();
        for (int r = 0; r < rows; ++r) {
            const float d = s_d[r][co];
#pragma unroll
            for (int j = 0; j < kColMaxK / 4; ++j)
                if (j < kper) acc[j] = fmaf(d, s_c[r][kg * kper + j], acc[j]);
        }
    }
    float* out = partial + (int64_t)blockIdx.x * 64 * Kpad + (int64_t)co * Kpad + kg * kper;
#pragma unroll
    for (int j = 0; j < kColMaxK / 4; ++j)
        if (j < kper) out[j] = acc[j];
}

__global__ void col_wgrad_reduce_kernel(const float* __restrict__ partial, int nblk, int Kpad, int K, float* __restrict__ grad_w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * K) return;
    const int co = i / K, k = i % K;
    float s = 0.f;
    for (int b = 0; b < nblk; ++b) s += partial[(int64_t)b * 64 * Kpad + co * Kpad + k];
    grad_w[i] += s;
}

constexpr int kColBlocks = 592;

int launch_col_wgrad(const float* dc, const float* col, float* partial, float* grad_w, int64_t M, int Cout, int K, int Kpad,
                     cudaStream_t st) {
    if (Cout != 64 || Kpad % 4 || Kpad > kColMaxK || K > Kpad) return ORBIT_ERR_UNSUPPORTED;
    if (M <= 0) return ORBIT_OK;
    const int64_t rows_per_block = ceil_div64(ceil_div64(M, kColBlocks), kColRows) * kColRows;
    const int nblk = (int)ceil_div64(M, rows_per_block);
    col_wgrad_kernel<<<nblk, 256, 0, st>>>(dc, col, partial, M, Kpad, rows_per_block);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    col_wgrad_reduce_kernel<<<ceil_div(64 * K, 256), 256, 0, st>>>(partial, nblk, Kpad, K, grad_w);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// conv weight [Cout, Cin, 3, 3] -> data-gradient GEMM weight [Cin][tap' * Cout + co] with tap' = (2 - ky) * 3 + (2 - kx):
// din[y,x,ci] = sum_{tap',co} dc[y + ky' - 1, x + kx' - 1, co] * W[co, ci, 2 - ky', 2 - kx']  (NHWC im2col column order)
__global__ void conv3_dgrad_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int Cin, int Cout) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)Cin * 9 * Cout) return;
    const int co = (int)(i % Cout), tap = (int)((i / Cout) % 9), ci = (int)(i / ((int64_t)9 * Cout));
    out[i] = w[((int64_t)co * Cin + ci) * 9 + (8 - tap)];
}
int launch_conv3_dgrad_weight(const float* w, float* out, int Cin, int Cout, cudaStream_t st) {
    conv3_dgrad_weight_kernel<<<(unsigned)ceil_div64((int64_t)Cin * 9 * Cout, 256), 256, 0, st>>>(w, out, Cin, Cout);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// ------------------------------------------------------------------------------------------------
// FiLM generator backward (model/feature_adapters.py:66-78, mlps.py:52-66): per generated tensor, one block.
//   h = W1 z + b1;  xhat = (h - mean) rstd;  t = xhat ln_w + ln_b;  a = relu(t);  g = W2 a + b2;
//   film = init (g r + 1)  (weights)   or   init + g r  (biases)
// Writes the gradients of W1, b1, ln_w, ln_b, W2, b2, r at the parameters' offsets of a gradient blob laid out like the
// generator's parameter blob, and dz_partial[tensor][hidden] (summed over tensors, fixed order, by the second kernel).
// ------------------------------------------------------------------------------------------------
struct FilmGenEntryT {
    int64_t w1, b1, ln_w, ln_b, w2, b2, reg, init, out;
    int32_t size, is_weight;
};

__global__ void __launch_bounds__(256)
film_generate_backward_kernel(const float* __restrict__ blob, const FilmGenEntryT* __restrict__ table, const float* __restrict__ z,
                              const float* __restrict__ dfilm, int hidden, float* __restrict__ grad_blob, float* __restrict__ dz_partial) {
    extern __shared__ float s_gb[];        // z[H], h[H], xhat[H], a[H], da[H], dh[H], red[4][H]
    float* s_z = s_gb; float* s_h = s_z + hidden; float* s_xh = s_h + hidden; float* s_a = s_xh + hidden;
    float* s_da = s_a + hidden; float* s_dh = s_da + hidden; float* s_red = s_dh + hidden;
    __shared__ float s_stat[4];
    const FilmGenEntryT e = table[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) s_z[i] = z[i];
    __syncthreads();
    for (int o = warp; o < hidden; o += n_warps) {
        const float* w = blob + e.w1 + (int64_t)o * hidden;
        float s = 0.f;
        for (int k = lane; k < hidden; k += 32) s = fmaf(__ldg(w + k), s_z[k], s);
        s = warp_sum(s);
        if (lane == 0) s_h[o] = s + blob[e.b1 + o];
    }
    __syncthreads();
    if (warp == 0) {
        float s = 0.f;
        for (int k = lane; k < hidden; k += 32) s += s_h[k];
        const float mean = warp_sum(s) / hidden;
        float v = 0.f;
        for (int k = lane; k < hidden; k += 32) { const float d = s_h[k] - mean; v = fmaf(d, d, v); }
        v = warp_sum(v) / hidden;
        if (lane == 0) { s_stat[0] = mean; s_stat[1] = 1.0f / sqrtf(v + 1e-5f); }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < hidden; k += blockDim.x) {
        const float xh = (s_h[k] - s_stat[0]) * s_stat[1];
        s_xh[k] = xh;
        s_a[k] = fmaxf(fmaf(xh, blob[e.ln_w + k], blob[e.ln_b + k]), 0.f);
    }
    for (int k = threadIdx.x; k < 4 * hidden; k += blockDim.x) s_red[k] = 0.f;
    __syncthreads();
    // output layer: thread per output o (strided), accumulating da[k] privately per 64-thread group through shared atomics-free
    // partials: group g = threadIdx.x / 64 owns s_red[g][.]; within a group the 64 threads walk o together and reduce by k.
    {
        const int grp = threadIdx.x >> 6, t = threadIdx.x & 63;      // 4 groups x 64 threads; thread t owns hidden index k = t (+64 j)
        for (int o = grp; o < e.size; o += 4) {
            // g[o] (recomputed), cooperative over the group
            const float* w = blob + e.w2 + (int64_t)o * hidden;
            float part = 0.f;
            for (int k = t; k < hidden; k += 64) part = fmaf(__ldg(w + k), s_a[k], part);
            // reduce the 64 partial sums: two warps per group
            part = warp_sum(part);
            __shared__ float s_pair[8];
            if (lane == 0) s_pair[warp] = part;
            asm volatile("bar.sync %0, 64;" :: "r"(grp + 1));
            const float gval = s_pair[grp * 2] + s_pair[grp * 2 + 1] + blob[e.b2 + o];
            const float r = blob[e.reg + o], init = blob[e.init + o];
            const float dG = dfilm[e.out + o] * (e.is_weight ? init : 1.0f);     // d loss / d (g r)
            const float dg = dG * r;
            if (t == 0) { grad_blob[e.reg + o] = dG * gval; grad_blob[e.b2 + o] = dg; }
            for (int k = t; k < hidden; k += 64) {
                grad_blob[e.w2 + (int64_t)o * hidden + k] = dg * s_a[k];
                s_red[grp * hidden + k] = fmaf(dg, __ldg(w + k), s_red[grp * hidden + k]);
            }
            asm volatile("bar.sync %0, 64;" :: "r"(grp + 1));
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < hidden; k += blockDim.x) {
        const float da = s_red[k] + s_red[hidden + k] + s_red[2 * hidden + k] + s_red[3 * hidden + k];
        const float dt = s_a[k] > 0.f ? da : 0.f;
        grad_blob[e.ln_w + k] = dt * s_xh[k];
        grad_blob[e.ln_b + k] = dt;
        s_da[k] = dt * blob[e.ln_w + k];                        // dxhat
    }
    __syncthreads();
    if (warp == 0) {
        float a = 0.f, b = 0.f;
        for (int k = lane; k < hidden; k += 32) { a += s_da[k]; b = fmaf(s_da[k], s_xh[k], b); }
        a = warp_sum(a) / hidden; b = warp_sum(b) / hidden;
        if (lane == 0) { s_stat[2] = a; s_stat[3] = b; }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < hidden; k += blockDim.x) {
        const float dh = s_stat[1] * (s_da[k] - s_stat[2] - s_xh[k] * s_stat[3]);
        s_dh[k] = dh;
        grad_blob[e.b1 + k] = dh;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < hidden * hidden; i += blockDim.x) grad_blob[e.w1 + i] = s_dh[i / hidden] * s_z[i % hidden];
    for (int j = threadIdx.x; j < hidden; j += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < hidden; ++k) s = fmaf(s_dh[k], __ldg(blob + e.w1 + (int64_t)k * hidden + j), s);
        dz_partial[(int64_t)blockIdx.x * hidden + j] = s;
    }
}

__global__ void film_dz_reduce_kernel(const float* __restrict__ dz_partial, int n, int hidden, float* __restrict__ dz) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= hidden) return;
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += dz_partial[(int64_t)i * hidden + j];
    dz[j] = s;
}

// ------------------------------------------------------------------------------------------------
// Head query path backward (classifier_heads.py:60-79,161-180,202-230): logits = s (q W^T + b) (linear / versa / proto
// euclidean) or s cos(q, W_c) (proto cosine), q = mean over the clip's L frame features. Given dlogits [N,C]:
//   dfeat[(n L + l), :] = dq[n, :] / L.   One block per clip.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_predict_backward_kernel(const float* __restrict__ feats, const float* __restrict__ weight, const float* __restrict__ dlogits,
                             int L, int D, int C, int metric, float logit_scale, float* __restrict__ dfeat) {
    extern __shared__ float s_hb[];      // q[D], coef[C], (cosine) dot[C], wn[C]
    float* s_q = s_hb; float* s_coef = s_q + D; float* s_dot = s_coef + C; float* s_wn = s_dot + C;
    __shared__ float s_qn;
    const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (metric == 1) {
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            float s = 0.f;
            for (int l = 0; l < L; ++l) s += feats[((int64_t)n * L + l) * D + d];
            s_q[d] = s / (float)L;
        }
        __syncthreads();
        for (int c = warp; c <= C; c += nw) {                    // c == C: |q|^2
            float dot = 0.f, wn = 0.f;
            for (int d = lane; d < D; d += 32) {
                const float qv = s_q[d], wv = c < C ? __ldg(weight + (int64_t)c * D + d) : qv;
                dot = fmaf(qv, wv, dot); wn = fmaf(wv, wv, wn);
            }
            dot = warp_sum(dot); wn = warp_sum(wn);
            if (lane == 0) { if (c < C) { s_dot[c] = dot; s_wn[c] = sqrtf(wn); } else s_qn = sqrtf(dot); }
        }
        __syncthreads();
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) s_coef[c] = logit_scale * dlogits[(int64_t)n * C + c];
    __syncthreads();
    const float invL = 1.0f / (float)L;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s = 0.f;
        if (metric == 0) {
            for (int c = 0; c < C; ++c) s = fmaf(s_coef[c], __ldg(weight + (int64_t)c * D + d), s);
        } else {
            // cos = dot / (max(|q|, eps) max(|w|, eps)):  d cos / dq = w / (|q||w|) - cos q / |q|^2
            const float qn = fmaxf(s_qn, 1e-8f), qv = s_q[d];
            for (int c = 0; c < C; ++c) {
                const float wn = fmaxf(s_wn[c], 1e-8f), cs = s_dot[c] / (qn * wn);
                s = fmaf(s_coef[c], __ldg(weight + (int64_t)c * D + d) / (qn * wn) - cs * qv / (qn * qn), s);
            }
        }
        s *= invL;
        for (int l = 0; l < L; ++l) dfeat[((int64_t)n * L + l) * D + d] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// Mahalanobis head query path backward (classifier_heads.py:328-350): logits[n,c] = -s (q_n - mu_c)^T P_c (q_n - mu_c), so
//   dq[n,:] = -s sum_c dlogits[n,c] (P_c + P_c^T) (q_n - mu_c),   dfeat[(n L + l),:] = dq[n,:] / L   (mean-pooled clips)
// (P_c + P_c^T, not 2 P_c: the computed inverse is symmetric only up to rounding and autograd differentiates what is there).
// One block per query clip; P_c is read twice per clip, row-wise (a warp per row) and column-wise (a thread per column).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mahalanobis_predict_backward_kernel(const float* __restrict__ feats, const float* __restrict__ means, const float* __restrict__ precisions,
                                    const float* __restrict__ dlogits, int L, int D, int C, float logit_scale, float* __restrict__ dfeat) {
    extern __shared__ float s_mb[];      // q[D], diff[D], acc[D]
    float* s_q = s_mb; float* s_diff = s_q + D; float* s_acc = s_diff + D;
    const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < L; ++l) s += feats[((int64_t)n * L + l) * D + d];
        s_q[d] = s / (float)L;
        s_acc[d] = 0.f;
    }
    __syncthreads();
    for (int c = 0; c < C; ++c) {
        const float coef = -logit_scale * dlogits[(int64_t)n * C + c];
        const float* P = precisions + (int64_t)c * D * D;
        for (int d = threadIdx.x; d < D; d += blockDim.x) s_diff[d] = s_q[d] - means[(int64_t)c * D + d];
        __syncthreads();
        for (int d = threadIdx.x; d < D; d += blockDim.x) {       // (P^T diff)[d] = sum_e P[e,d] diff[e]: coalesced over d
            float s = 0.f;
            for (int e = 0; e < D; ++e) s = fmaf(__ldg(P + (int64_t)e * D + d), s_diff[e], s);
            s_acc[d] = fmaf(coef, s, s_acc[d]);
        }
        __syncthreads();
        for (int d = warp; d < D; d += nw) {                      // (P diff)[d]: a warp per row
            float s = 0.f;
            for (int e = lane; e < D; e += 32) s = fmaf(__ldg(P + (int64_t)d * D + e), s_diff[e], s);
            s = warp_sum(s);
            if (lane == 0) s_acc[d] = fmaf(coef, s, s_acc[d]);
        }
        __syncthreads();
    }
    const float invL = 1.0f / (float)L;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float v = s_acc[d] * invL;
        for (int l = 0; l < L; ++l) dfeat[((int64_t)n * L + l) * D + d] = v;
    }
}

}  // namespace orbit

using namespace orbit;

extern "C" int orbit_mahalanobis_predict_backward(const float* frame_feats, const float* means, const float* precisions,
                                                  const float* grad_logits, int num_clips, int clip_length, int feat_dim,
                                                  int num_classes, float logit_scale, float* grad_frame_feats, void* stream) {
    if (!frame_feats || !means || !precisions || !grad_logits || !grad_frame_feats) return ORBIT_ERR_ARG;
    if (num_clips < 0 || clip_length <= 0 || feat_dim <= 0 || num_classes <= 0) return ORBIT_ERR_ARG;
    if (num_clips == 0) return ORBIT_OK;
    const size_t smem = sizeof(float) * 3 * (size_t)feat_dim;
    if (smem > 200 * 1024) return ORBIT_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) ORBIT_CUDA(cudaFuncSetAttribute(mahalanobis_predict_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mahalanobis_predict_backward_kernel<<<num_clips, 256, smem, (cudaStream_t)stream>>>(frame_feats, means, precisions, grad_logits, clip_length,
                                                                                       feat_dim, num_classes, logit_scale, grad_frame_feats);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

extern "C" int orbit_film_generate_backward(const float* gen_params, const void* table, int num_tensors, const float* task_embedding,
                                            int hidden, const float* grad_film, float* grad_gen_params, float* grad_embedding,
                                            float* scratch, void* stream) {
    if (!gen_params || !table || !task_embedding || !grad_film || !grad_gen_params || !grad_embedding || !scratch) return ORBIT_ERR_ARG;
    if (num_tensors <= 0 || hidden <= 0 || hidden % 4 || hidden > 1024) return ORBIT_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * (size_t)hidden * 10;
    if (smem > 48 * 1024) ORBIT_CUDA(cudaFuncSetAttribute(film_generate_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    film_generate_backward_kernel<<<num_tensors, 256, smem, st>>>(gen_params, reinterpret_cast<const FilmGenEntryT*>(table), task_embedding,
                                                                 grad_film, hidden, grad_gen_params, scratch);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    film_dz_reduce_kernel<<<ceil_div(hidden, 128), 128, 0, st>>>(scratch, num_tensors, hidden, grad_embedding);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

extern "C" int orbit_head_predict_backward(const float* frame_feats, const float* weight, const float* grad_logits, int num_clips,
                                           int clip_length, int feat_dim, int num_classes, int metric, float logit_scale,
                                           float* grad_frame_feats, void* stream) {
    if (!frame_feats || !weight || !grad_logits || !grad_frame_feats) return ORBIT_ERR_ARG;
    if (num_clips < 0 || clip_length <= 0 || feat_dim <= 0 || num_classes <= 0) return ORBIT_ERR_ARG;
    if (metric != ORBIT_METRIC_EUCLIDEAN && metric != ORBIT_METRIC_COSINE) return ORBIT_ERR_UNSUPPORTED;
    if (num_clips == 0) return ORBIT_OK;
    const size_t smem = sizeof(float) * (size_t)(feat_dim + 3 * num_classes);
    if (smem > 200 * 1024) return ORBIT_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) ORBIT_CUDA(cudaFuncSetAttribute(head_predict_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_predict_backward_kernel<<<num_clips, 256, smem, (cudaStream_t)stream>>>(frame_feats, weight, grad_logits, clip_length, feat_dim,
                                                                                num_classes, metric == ORBIT_METRIC_COSINE ? 1 : 0,
                                                                                logit_scale, grad_frame_feats);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}
