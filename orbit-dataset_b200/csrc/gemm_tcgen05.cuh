// tcgen05 (5th-gen tensor core) pointwise-conv GEMM: TMA-fed, TMEM accumulators, FP16x3 split for
// fp32-grade accuracy, with the BN/FiLM scale-shift + activation + residual epilogue. Internal.
#pragma once
#include "common.cuh"

namespace orbit {

// w [N,K] fp32 -> out (as fp16): hi [N,Kp] | lo [N,Kp], Kp = K rounded up to 8; hi = fp16(w), lo = fp16((w - hi) * 2^11).
// `out` needs N*Kp floats (<= 2*N*K).
int launch_weight_split(const float* w, int N, int K, float* out, cudaStream_t st);

// out[M,N] = act((A[M,K] (*gate)) W[N,K]^T * scale + shift) (+ residual); w_split = [hi | lo] from launch_weight_split.
// passes = 3 (hi*hi + hi*lo + lo*hi, fp32-grade) or 1 (one fp16 product: the `fast` numerics mode).
int launch_pointwise_tcgen05(const float* A, const float* w_split, const float* scale, const float* shift,
                             const float* gate, const float* residual, float* out, int M, int N, int K,
                             int rows_per_frame, int act, int passes, cudaStream_t st);

// Implicit-GEMM 3x3 convolution (stride 1, padding 1) on the same kernel: x [B,H,W,Cin] NHWC (Cin % 64 == 0), w_split = split of the
// weights re-laid-out to [N][9*Cin] in (tap, channel) order (launch_conv_weight_relayout, nchw = 0). No im2col matrix: the TMA
// producer walks the nine taps with row-shifted boxes of the activation, the transform warps zero what falls outside the image.
// ORBIT_ERR_UNSUPPORTED when the shape / epilogue has no instance (callers fall back to im2col + launch_pointwise_tcgen05).
int launch_conv3x3_tcgen05(const float* x, const float* w_split, const float* scale, const float* shift, const float* residual,
                           float* out, int B, int H, int W, int Cin, int N, int act, cudaStream_t st);

// Row-streaming variant for the small-K / small-N layers (csrc/gemm_stream.cu): same contract and numerics scheme (FP16x3);
// returns ORBIT_ERR_UNSUPPORTED when it has no instance for the shape. launch_pointwise_tcgen05 tries it first (passes == 3).
int launch_pointwise_stream(const float* A, const float* w_split, const float* scale, const float* shift, const float* gate,
                            const float* residual, float* out, int M, int N, int K, int rows_per_frame, int act, cudaStream_t st);
void set_stream_gemm(int on);
int get_stream_gemm();

// kappa of the truncation de-biasing applied to every promoted k-block partial in FP16x3 mode (process-wide)
void set_tcgen05_debias(float kappa);
// dev aid: device buffer of 256 x 16 uint32 that CTA 0 of every following GEMM launch fills with per-role clock stamps (null = off)
void set_tcgen05_trace(unsigned* dev_buffer);
float get_tcgen05_debias();
// dev A/B switch: 4-lanes-per-row transform mapping for K <= 32 (default on)
void set_tcgen05_wide_xf(int v);   // 8 transform warps for ungated multi-n-tile layers: 0 off, 1 Linear, 2 + SiLU expands
int get_tcgen05_wide_xf();
void set_tcgen05_narrow(int on);
int get_tcgen05_narrow();
// dev A/B switches: fixed slab -> warp mapping for tiles of < 3 slabs (frees staging memory for ring stages); minimum ring
// depth at which the epilogue staging is double buffered. -1 leaves a value unchanged.
void set_tcgen05_tuning(int fixed_slabs, int double_min_stages);

}  // namespace orbit
