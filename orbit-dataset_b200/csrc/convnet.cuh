// Launchers of the convolutional-backbone kernels (NHWC fp32 activations). Internal to the library.
#pragma once
#include "common.cuh"

namespace orbit {

enum Act { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2, ACT_ELU = 3, ACT_GELU = 4 };

// One BatchNorm (or FiLM-modulated BatchNorm) to fold into per-channel scale/shift.
struct FoldEntry {
    int64_t gamma, beta, mean, var;  // offsets into the params blob
    int64_t film_gamma, film_beta;   // offsets into the film blob, or -1
    int64_t conv_bias;               // offset of the preceding conv's bias in params (folded into shift), or -1
    int64_t out;                     // offset into derived: scale[C] then shift[C]
    int channels;
    float eps;
};

int launch_bn_fold(const FoldEntry* entries_host, int n, const float* params, const float* film, float* derived,
                   cudaStream_t st);

// a[i] += b[i]
int launch_add_vec(float* a, const float* b, int n, cudaStream_t st);

// out = [ones(n) | zeros(n)]
int launch_fill_identity(float* out, int n, cudaStream_t st);

// per-channel batch statistics of x [M,C]: mean and UNBIASED variance (double accumulation, deterministic)
int launch_channel_stats(const float* x, int64_t M, int C, float* mean, float* var, cudaStream_t st);

// depthwise weights [C,1,k,k] -> [k*k][C]
int launch_dw_relayout(const float* w, int C, int kk, float* out, cudaStream_t st);

// stem: x [B,3,H,W] NCHW -> y [B,Ho,Wo,32] NHWC, 3x3 stride 2, pad (top,left), scale/shift + act
int launch_stem(const float* x, const float* w, const float* scale, const float* shift, float* y, int B, int H, int W,
                int Ho, int Wo, int pad_t, int pad_l, int cout, int act, cudaStream_t st);

// depthwise kxk: x [B,H,W,C] -> y [B,Ho,Wo,C]; wt is [k*k][C]; also writes per-(frame,tile,channel) sums
// of the activated output into partial [B][groups][C] (deterministic SE squeeze), groups = dw_partial_groups(...)
int dw_partial_groups(int C, int Ho, int Wo, int k, int stride);
int launch_depthwise(const float* x, const float* wt, const float* scale, const float* shift, float* y, float* partial,
                     int B, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad_t, int pad_l, int act,
                     cudaStream_t st);


// dev A/B switch: shared-memory-staged 5x5 stride-1 depthwise kernel for 14x14 / 7x7 inputs (default on)
void set_dw5_staged(int on);
int get_dw5_staged();

// Fused expand 1x1 (+bn1+SiLU) -> depthwise kxk (+bn2/FiLM+SiLU, SE squeeze partials) for MBConv blocks with 16 / 24 input
// channels: xin [B,H,W,Cin] -> y [B,Ho,Wo,C]; we = expand weights [C][Cin] (torch layout), scale1/shift1 = folded bn1,
// wt = depthwise taps [k*k][C], scale/shift = folded bn2. partial as launch_depthwise with mbx_partial_groups(...).
bool mbx_supported(int cin, int k, int stride);
bool mbx_fits(int Cin, int C, int Ho, int Wo, int k, int stride);
int mbx_partial_groups(int Cin, int C, int Ho, int Wo, int k, int stride);
// register-resident variant for the 3x3 stride-2 block with 16 input channels (csrc/mbconv_stream.cu); dev A/B switch "mbconv_stream"
bool mbs_supported(int Cin, int C, int H, int W, int k, int stride);
int mbs_partial_groups(int Ho);
int launch_mbconv_stream(const float* xin, const float* we, const float* scale1, const float* shift1, const float* wt,
                         const float* scale, const float* shift, float* y, float* partial, int B, int H, int W, int Cin, int C,
                         cudaStream_t st);
void set_mbconv_stream(int on);
int get_mbconv_stream();
int launch_mbconv_expand_dw(const float* xin, const float* we, const float* scale1, const float* shift1, const float* wt,
                            const float* scale, const float* shift, float* y, float* partial, int B, int H, int W, int Cin,
                            int C, int Ho, int Wo, int k, int stride, int pad_t, int pad_l, cudaStream_t st);

// squeeze-excite gate: partial [B][tiles][C] -> gate [B][C] = sigmoid(W2 silu(W1 mean + b1) + b2);
// w2t = the expand weight [C][R] transposed to [R][C] (launch_dw_relayout(w2, C, R, w2t))
void set_stem_groups(int n);   // dev A/B switch: 128-pixel-pair groups a stem block walks over (default 8)
int get_stem_groups();
void set_se_ring(int on);   // dev A/B switch: ring-streamed SE gate kernel (default) or the plain one
int get_se_ring();
int launch_se_gate(const float* partial, int tiles, int hw, const float* w1, const float* b1, const float* w2t,
                   const float* b2, float* gate, int B, int C, int R, cudaStream_t st);

// pointwise conv as GEMM: out[M,N] = act((A[M,K] (*gate[m/rows_per_frame, k])) W[N,K]^T * scale[n] + shift[n]) (+res)
// act >= 16: (act - 16) is applied AFTER the residual add (ResNet BasicBlock: relu(bn(conv) + identity))
int launch_pointwise_ffma(const float* A, const float* Wt, const float* scale, const float* shift, const float* gate,
                          const float* residual, float* out, int M, int N, int K, int rows_per_frame, int act,
                          cudaStream_t st);

// k x k / stride im2col with top/left padding (pad_t, pad_l); taps beyond the bottom/right edge read zero. nchw=1: x [B,C,H,W] -> col [B*Ho*Wo, Kpad], col index = c*k*k + tap (torch
// weight order); nchw=0: x [B,H,W,C] -> col index = tap*C + c. Columns >= k*k*C are zero.
int launch_im2col(const float* x, float* col, int B, int H, int W, int C, int k, int stride, int pad_t, int pad_l, int Ho,
                  int Wo, int Kpad, int nchw, cudaStream_t st);
// conv weight [Cout, Cin, k, k] -> [Cout, Kpad] in the im2col column order above
int launch_conv_weight_relayout(const float* w, float* out, int Cout, int Cin, int kk, int Kpad, int nchw, cudaStream_t st);
// k x k / stride max pool with symmetric padding (floor mode), NHWC: (2,2,0) for the set encoder, (3,2,1) for resnet
int launch_maxpool(const float* x, float* y, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                   cudaStream_t st);

// first convolution on 3-channel NCHW frames (3x3 stride 1 pad 1, 7x7 stride 2 pad 3; 64 output channels) as a direct tensor-core
// convolution (csrc/conv_first.cu): x [B,3,H,W] -> y [B,Ho,Wo,64] NHWC, w in torch layout [64,3,k,k]; act ACT_NONE / ACT_RELU.
// ORBIT_ERR_UNSUPPORTED for other geometries (callers fall back to im2col + GEMM). dev A/B switch "conv_first".
int launch_conv_first(const float* x, const float* w, const float* scale, const float* shift, float* y, int B, int H, int W, int Cin,
                      int Cout, int k, int stride, int pad, int Ho, int Wo, int act, cudaStream_t st);
void set_conv_first(int on);
int get_conv_first();

// spatial mean: x [B,HW,C] -> y [B,C]
int launch_spatial_mean(const float* x, float* y, int B, int HW, int C, cudaStream_t st);

// ---- training (csrc/train.cu): forward pieces that keep pre-activations, and the backward kernels --------------------
// y = act(scale x + shift), act in {ACT_NONE, ACT_SILU}
int launch_bn_act_forward(const float* x, const float* scale, const float* shift, float* y, int64_t M, int C, int act, cudaStream_t st);
// backward of the above; see train.cu for `mode`. grad_gamma / grad_beta (nullable) are ACCUMULATED; partial needs
// bn_act_backward_partial_floats(M, C) floats when they are requested.
int64_t bn_act_backward_partial_floats(int64_t M, int C);
int launch_bn_act_backward(const float* c, const float* dy, const float* aux0, const float* aux1, const float* scale,
                           const float* shift, const float* mean, const float* var, float eps, float* dc, float* partial,
                           float* grad_gamma, float* grad_beta, int64_t M, int C, int act, int mode, int rows_per_frame,
                           cudaStream_t st);
int launch_dw_dgrad(const float* dy, const float* wt, float* dx, int B, int H, int W, int C, int Ho, int Wo, int k, int stride,
                    int pad_t, int pad_l, cudaStream_t st);
int launch_se_backward(const float* dga, const float* a, const float* m, const float* w1, const float* b1, const float* w2,
                       const float* b2, float* dgate, float* dmean, int B, int HW, int C, int R, cudaStream_t st);
int launch_transpose(const float* w, float* out, int rows, int cols, cudaStream_t st);

// ---- set-encoder training (csrc/train_setenc.cu) -----------------------------------------------------------------------
// p = maxpool2x2(relu(scale c + shift)), floor mode
int launch_bn_relu_pool2_forward(const float* c, const float* scale, const float* shift, float* p, int B, int H, int W, int C,
                                 cudaStream_t st);
// backward of the above (mode 0: dp [B,H/2,W/2,C]; mode 1: dp = dfeat [B,C] / (H/2 * W/2)); dc [B,H,W,C]; the BatchNorm weight /
// bias and conv-bias gradients are ACCUMULATED; partial: pool_bn_relu_backward_partial_floats(...) floats
int64_t pool_bn_relu_backward_partial_floats(int B, int H, int W, int C);
int launch_pool_bn_relu_backward(const float* c, const float* dp, const float* scale, const float* shift, const float* conv_bias,
                                 const float* mean, const float* var, float eps, float* dc, float* partial, float* grad_gamma,
                                 float* grad_beta, float* grad_conv_bias, int B, int H, int W, int C, int mode, cudaStream_t st);
// conv3x3 (pad 1, stride 1) weight gradient, NHWC input: grad_w [Cout,Cin,3,3] += ...; partial: scratch of `partial_capacity`
// floats (at least 9 Cin Cout; more lets more blocks work in parallel)
int launch_conv3_wgrad(const float* dc, const float* in, float* partial, int64_t partial_capacity, float* grad_w, int B, int H, int W,
                       int Cin, int Cout, cudaStream_t st);
// weight gradient from an im2col matrix col [M,Kpad]: grad_w [64,K] += dc^T col; partial: at most 56 M floats
int launch_col_wgrad(const float* dc, const float* col, float* partial, float* grad_w, int64_t M, int Cout, int K, int Kpad,
                     cudaStream_t st);
// conv weight [Cout,Cin,3,3] -> [Cin][9*Cout] for the data-gradient GEMM over im2col(dc)
int launch_conv3_dgrad_weight(const float* w, float* out, int Cin, int Cout, cudaStream_t st);

}  // namespace orbit
