// Launchers of the ViT-specific kernels (token tensors are row-major [frames*tokens, dim] fp32). Internal.
#pragma once
#include "common.cuh"

namespace orbit {

// frames [B,3,H,W] NCHW -> col [B*gh*gw, 3*P*P], k = c*P*P + py*P + px (order of the flattened conv weight)
int launch_patch_im2col(const float* frames, float* col, int B, int H, int W, int P, cudaStream_t st);
// x[b,0,:] = cls + pos[0]; x[b,1+p,:] = patches[b*np+p,:] + pos[1+p]
int launch_assemble_tokens(const float* patches, const float* cls, const float* pos, float* x, int B, int np, int D,
                           cudaStream_t st);
// y[r,:] = (x[r*row_stride ...] - mean) * rstd * gamma + beta for r < rows (one warp per row)
int launch_layernorm(const float* x, int64_t row_stride, const float* gamma, const float* beta, float eps, float* y,
                     int64_t out_stride, int rows, int D, cudaStream_t st);
// qkv [B,T,3*D] (which-major, then head, then dh) -> out [B,T,D]; softmax(q k^T / sqrt(dh)) v per (frame, head)
int launch_attention(const float* qkv, float* out, int B, int T, int heads, int dh, cudaStream_t st);
// dst = film ? film[...] : params[...] for one LayerNorm's gamma and beta (effective affine parameters of a task)
int launch_ln_affine(const float* gamma, const float* beta, float* out, int D, cudaStream_t st);

}  // namespace orbit
