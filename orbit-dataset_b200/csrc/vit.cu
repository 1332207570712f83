// ViT kernels: patch gather, token assembly, LayerNorm (the ViT FiLM site, reference model/film.py:57-66),
// and the 50-token attention. Reference op sites: timm 0.6.12 VisionTransformer (vit_{small,base}_patch32_224*)
// invoked at model/few_shot_recognisers.py:114-117,143-146; all Linear layers run through the pointwise GEMM.
#include "vit.cuh"

namespace orbit {

__global__ void __launch_bounds__(256)
patch_im2col_kernel(const float* __restrict__ x, float* __restrict__ col, int B, int H, int W, int P) {
    const int gh = H / P, gw = W / P, K = 3 * P * P, P4 = P >> 2;
    const int64_t total = (int64_t)B * gh * gw * 3 * P * P4;     // float4 elements (px is the contiguous axis)
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % P4);
        int64_t r = i / P4;
        const int py = (int)(r % P); r /= P;
        const int c = (int)(r % 3); r /= 3;
        const int gx = (int)(r % gw); r /= gw;
        const int gy = (int)(r % gh);
        const int b = (int)(r / gh);
        const float4 v = ldg4_stream(x + (((int64_t)b * 3 + c) * H + gy * P + py) * W + gx * P + q * 4);
        *reinterpret_cast<float4*>(col + (((int64_t)b * gh + gy) * gw + gx) * K + (c * P + py) * P + q * 4) = v;
    }
}
int launch_patch_im2col(const float* frames, float* col, int B, int H, int W, int P, cudaStream_t st) {
    if (P % 4 || H % P || W % P || W % 4) return ORBIT_ERR_UNSUPPORTED;
    const int64_t total = (int64_t)B * (H / P) * (W / P) * 3 * P * (P / 4);
    patch_im2col_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 148 * 16), 256, 0, st>>>(frames, col, B, H, W, P);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

__global__ void __launch_bounds__(256)
assemble_tokens_kernel(const float* __restrict__ patches, const float* __restrict__ cls, const float* __restrict__ pos,
                       float* __restrict__ x, int B, int np, int D) {
    const int D4 = D >> 2, T = np + 1;
    const int64_t total = (int64_t)B * T * D4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % D4);
        const int t = (int)((i / D4) % T);
        const int b = (int)(i / ((int64_t)D4 * T));
        float4 v = t == 0 ? ldg4(cls + q * 4) : ldg4_stream(patches + ((int64_t)b * np + t - 1) * D + q * 4);
        const float4 pe = ldg4(pos + (int64_t)t * D + q * 4);
        v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
        *reinterpret_cast<float4*>(x + ((int64_t)b * T + t) * D + q * 4) = v;
    }
}
int launch_assemble_tokens(const float* patches, const float* cls, const float* pos, float* x, int B, int np, int D,
                           cudaStream_t st) {
    if (D % 4) return ORBIT_ERR_UNSUPPORTED;
    const int64_t total = (int64_t)B * (np + 1) * (D / 4);
    assemble_tokens_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 148 * 16), 256, 0, st>>>(patches, cls, pos, x, B, np, D);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// One warp per row, row held in registers (D <= 1024): mean, then centred variance (two-pass, like torch).
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int64_t row_stride, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, float* __restrict__ y, int64_t out_stride, int rows, int D) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + (int64_t)row * row_stride;
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int d = (i * 32 + lane) * 4;
        v[i] = d < D ? ldg4(xr + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int d = (i * 32 + lane) * 4;
        if (d < D) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
            q += a * a + b * b + c * c + e * e;
        }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
    float* yr = y + (int64_t)row * out_stride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int d = (i * 32 + lane) * 4;
        if (d < D) {
            const float4 g = ldg4(gamma + d), bb = ldg4(beta + d);
            float4 o;
            o.x = (v[i].x - mean) * rstd * g.x + bb.x; o.y = (v[i].y - mean) * rstd * g.y + bb.y;
            o.z = (v[i].z - mean) * rstd * g.z + bb.z; o.w = (v[i].w - mean) * rstd * g.w + bb.w;
            *reinterpret_cast<float4*>(yr + d) = o;
        }
    }
}
int launch_layernorm(const float* x, int64_t row_stride, const float* gamma, const float* beta, float eps, float* y,
                     int64_t out_stride, int rows, int D, cudaStream_t st) {
    if (D % 4 || D > 1024) return ORBIT_ERR_UNSUPPORTED;
    if (rows <= 0) return ORBIT_OK;
    layernorm_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(x, row_stride, gamma, beta, eps, y, out_stride, rows, D);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

// One block per (frame, head): Q, K, V tiles [T, dh] in shared memory, scores [T, T], row softmax by warps, then O = P V.
// T = 50, dh = 64 for ViT-*/32 at 224 px. Both products are register-tiled 4 x 4 per thread with 128-bit shared-memory loads
// (the first version did two scalar LDS per FMA and was bound by shared-memory bandwidth: 0.3 ms per layer for 240 frames).
__global__ void __launch_bounds__(256)
attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int heads, int dh) {
    extern __shared__ __align__(16) float s_att[];
    const int D = heads * dh, ld = dh + 4;          // rows stay 16-byte aligned; +4 floats: consecutive rows start 4 banks apart
    const int Tp = (T + 3) & ~3;                    // rows / columns padded to whole 4 x 4 tiles (zero-filled)
    float* s_q = s_att;                 // [Tp][ld]
    float* s_k = s_q + Tp * ld;         // [Tp][ld]
    float* s_v = s_k + Tp * ld;         // [Tp][ld]
    float* s_p = s_v + Tp * ld;         // [Tp][Tp]
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const float* base = qkv + (int64_t)b * T * 3 * D + h * dh;
    const int dh4 = dh >> 2;
    for (int i = threadIdx.x; i < Tp * dh4; i += blockDim.x) {
        const int t = i / dh4, d = (i % dh4) * 4;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f), k = q, v = q;
        if (t < T) {
            const float* p = base + (int64_t)t * 3 * D + d;
            q = ldg4(p); k = ldg4(p + D); v = ldg4(p + 2 * D);
        }
        *reinterpret_cast<float4*>(s_q + t * ld + d) = q;
        *reinterpret_cast<float4*>(s_k + t * ld + d) = k;
        *reinterpret_cast<float4*>(s_v + t * ld + d) = v;
    }
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)dh);
    const int nt = Tp >> 2;
    for (int tile = threadIdx.x; tile < nt * nt; tile += blockDim.x) {       // S = scale Q K^T, 4 x 4 tile per thread
        const int r0 = (tile / nt) * 4, c0 = (tile % nt) * 4;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int d = 0; d < dh; d += 4) {
            float4 q[4], k[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                q[i] = *reinterpret_cast<const float4*>(s_q + (r0 + i) * ld + d);
                k[i] = *reinterpret_cast<const float4*>(s_k + (c0 + i) * ld + d);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(q[i].x, k[j].x, acc[i][j]); acc[i][j] = fmaf(q[i].y, k[j].y, acc[i][j]);
                    acc[i][j] = fmaf(q[i].z, k[j].z, acc[i][j]); acc[i][j] = fmaf(q[i].w, k[j].w, acc[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(s_p + (r0 + i) * Tp + c0) = make_float4(acc[i][0] * scale, acc[i][1] * scale, acc[i][2] * scale, acc[i][3] * scale);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int r = warp; r < T; r += n_warps) {
        float m = -INFINITY;
        for (int c = lane; c < T; c += 32) m = fmaxf(m, s_p[r * Tp + c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float z = 0.f;
        for (int c = lane; c < T; c += 32) { const float e = expf(s_p[r * Tp + c] - m); s_p[r * Tp + c] = e; z += e; }
        z = warp_sum(z);
        for (int c = lane; c < Tp; c += 32) s_p[r * Tp + c] = c < T ? s_p[r * Tp + c] / z : 0.f;     // padded columns: exact zeros
    }
    __syncthreads();
    for (int tile = threadIdx.x; tile < nt * dh4; tile += blockDim.x) {      // O = P V, 4 rows x 4 channels per thread
        const int r0 = (tile / dh4) * 4, d0 = (tile % dh4) * 4;
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < Tp; c += 4) {
            float4 p[4], v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                p[i] = *reinterpret_cast<const float4*>(s_p + (r0 + i) * Tp + c);
                v[i] = *reinterpret_cast<const float4*>(s_v + (c + i) * ld + d0);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i].x = fmaf(p[i].x, v[0].x, acc[i].x); acc[i].y = fmaf(p[i].x, v[0].y, acc[i].y); acc[i].z = fmaf(p[i].x, v[0].z, acc[i].z); acc[i].w = fmaf(p[i].x, v[0].w, acc[i].w);
                acc[i].x = fmaf(p[i].y, v[1].x, acc[i].x); acc[i].y = fmaf(p[i].y, v[1].y, acc[i].y); acc[i].z = fmaf(p[i].y, v[1].z, acc[i].z); acc[i].w = fmaf(p[i].y, v[1].w, acc[i].w);
                acc[i].x = fmaf(p[i].z, v[2].x, acc[i].x); acc[i].y = fmaf(p[i].z, v[2].y, acc[i].y); acc[i].z = fmaf(p[i].z, v[2].z, acc[i].z); acc[i].w = fmaf(p[i].z, v[2].w, acc[i].w);
                acc[i].x = fmaf(p[i].w, v[3].x, acc[i].x); acc[i].y = fmaf(p[i].w, v[3].y, acc[i].y); acc[i].z = fmaf(p[i].w, v[3].z, acc[i].z); acc[i].w = fmaf(p[i].w, v[3].w, acc[i].w);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (r0 + i < T) *reinterpret_cast<float4*>(out + ((int64_t)b * T + r0 + i) * D + h * dh + d0) = acc[i];
    }
}
int launch_attention(const float* qkv, float* out, int B, int T, int heads, int dh, cudaStream_t st) {
    if (dh % 4 || (heads * dh) % 4) return ORBIT_ERR_UNSUPPORTED;
    const int Tp = (T + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)3 * Tp * (dh + 4) + (size_t)Tp * Tp);
    if (smem > 200 * 1024) return ORBIT_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) ORBIT_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (B <= 0) return ORBIT_OK;
    attention_kernel<<<B * heads, 256, smem, st>>>(qkv, out, T, heads, dh);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

__global__ void ln_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out, int D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D) { out[i] = gamma[i]; out[D + i] = beta[i]; }
}
int launch_ln_affine(const float* gamma, const float* beta, float* out, int D, cudaStream_t st) {
    ln_affine_kernel<<<ceil_div(D, 256), 256, 0, st>>>(gamma, beta, out, D);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

}  // namespace orbit
