// tcgen05 (5th-gen tensor core) pointwise-conv GEMM: TMA-fed, TMEM accumulators, 3xTF32 split for
// fp32-grade accuracy, with the BN/FiLM scale-shift + activation + residual epilogue. Internal.
#pragma once
#include "common.cuh"

namespace orbit {

// w [n] fp32 -> out [2][n]: hi = tf32(w) (round to nearest), lo = tf32(w - hi)
int launch_tf32_split(const float* w, int64_t n, float* out, cudaStream_t st);

// out[M,N] = act((A[M,K] (*gate)) W[N,K]^T * scale + shift) (+ residual); w_split = [hi | lo] from launch_tf32_split.
// passes = 3 (hi*hi + hi*lo + lo*hi, fp32-grade) or 1 (plain tf32).
int launch_pointwise_tcgen05(const float* A, const float* w_split, const float* scale, const float* shift,
                             const float* gate, const float* residual, float* out, int M, int N, int K,
                             int rows_per_frame, int act, int passes, cudaStream_t st);

// kappa of the truncation de-biasing applied to every promoted k-block partial in 3xTF32 mode (process-wide)
void set_tcgen05_debias(float kappa);
// A-in-TMEM variant of the K-heavy 3xTF32 layers on/off (A/B switch; default on)
void set_tcgen05_atm(bool on);
bool get_tcgen05_atm();
// merged a_hi.[b_hi;b_lo] products for the gated projections on/off (A/B switch; default OFF: measured slower)
void set_tcgen05_merge(bool on);
bool get_tcgen05_merge();
// dev aid: device buffer of 256 x 16 uint32 that CTA 0 of every following GEMM launch fills with per-role clock stamps (null = off)
void set_tcgen05_trace(unsigned* dev_buffer);
float get_tcgen05_debias();

}  // namespace orbit
