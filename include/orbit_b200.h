/* orbit_b200.h -- C ABI of liborbit_b200.so: the sm_100a implementation of the ORBIT few-shot
 * recogniser's inner episodic loop (reference: microsoft/ORBIT-Dataset @ 97ccae1,
 * model/few_shot_recognisers.py personalise()/predict()).
 *
 * Conventions
 *   - every function returns int: 0 = ok, >0 = cudaError_t, <0 = ORBIT_ERR_* (argument errors).
 *   - no exceptions, no hidden allocation, no stream synchronisation across the ABI: all device
 *     memory (parameters, workspace, outputs) is caller-owned; every call takes the cudaStream_t it
 *     enqueues on (as void*). Pointers are DEVICE pointers unless the name ends in _host.
 *   - activations are fp32; frames come in as the reference delivers them (fp32, NCHW, contiguous,
 *     data/datasets.py:384,428-431), features/logits go out as the reference returns them
 *     (fp32 row-major).
 *   - thread-safety: functions are re-entrant; an orbit_engine is immutable after create and may be
 *     shared by threads that use distinct workspaces.
 */
#ifndef ORBIT_B200_H
#define ORBIT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBIT_ABI_VERSION 1

#define ORBIT_OK               0
#define ORBIT_ERR_ARG         (-1)  /* null pointer / non-positive size / misaligned pointer   */
#define ORBIT_ERR_UNSUPPORTED (-2)  /* shape or option outside what the kernels implement      */
#define ORBIT_ERR_WORKSPACE   (-3)  /* caller's workspace is smaller than *_workspace_bytes()   */
#define ORBIT_ERR_NO_DEVICE   (-4)  /* not running on an sm_100 device                         */

#define ORBIT_METRIC_EUCLIDEAN 0    /* classifier 'proto'        (few_shot_recognisers.py:78-79) */
#define ORBIT_METRIC_COSINE    1    /* classifier 'proto_cosine' (few_shot_recognisers.py:80-81) */

#define ORBIT_ARCH_EFFICIENTNET_B0 0 /* timm tf_efficientnet_b0 (feature_extractors.py:39-43)   */
#define ORBIT_ARCH_VIT_S_32        1 /* feature_extractors.py:49-53                             */
#define ORBIT_ARCH_VIT_B_32        2 /* feature_extractors.py:54-58                             */
#define ORBIT_ARCH_VIT_B_32_CLIP   3 /* feature_extractors.py:59-64                             */
#define ORBIT_ARCH_RESNET18        4 /* BASELINE.json extension (not in the reference)          */
#define ORBIT_ARCH_EFFICIENTNET_V2_S 5 /* timm tf_efficientnetv2_s_in21k (feature_extractors.py:44-48) */
#define ORBIT_ARCH_SET_ENCODER   100 /* SetEncoder / SimplePrePoolNet (model/set_encoders.py:34-120): frames ->
                                        64-d per-frame embedding, run through the same engine API               */

#define ORBIT_MAX_CLASSES 64

int         orbit_abi_version(void);
/* 1 when the library was compiled with a timing-experiment flag (-DORBIT_EXP_SKIP_A / -DORBIT_EXP_SKIP_B: wrong results by design);
 * the Python host side refuses to load such a build unless ORBIT_ALLOW_EXPERIMENT_BUILD is set.                                   */
int         orbit_experiment_build(void);
const char* orbit_error_string(int code);
/* 0 if the current CUDA device is sm_100 and the library's kernels can run, else ORBIT_ERR_NO_DEVICE */
int         orbit_device_check(void);

/* Process-wide numerics options. "tc_debias_x1000": kappa*1000 of the truncation de-biasing of the tcgen05
 * 3xTF32 GEMM (every tcgen05.mma result is truncated, not rounded, to fp32; kappa*ulp is added back to each
 * promoted k-block partial).                                                                            */
int orbit_set_global_option(const char* key, int value);
int orbit_get_global_option(const char* key, int* value);
/* Development aid: device buffer of 256 x 16 uint32 that CTA 0 of every following tcgen05 GEMM launch fills with
 * per-role clock stamps per k-block step (0/1 TMA producer, 2/3 transform, 4-7 MMA issuer, 8/9 epilogue); NULL = off. */
int orbit_debug_set_gemm_trace(void* dev_buffer);

/* ------------------------------------------------------------------------------------------------
 * Head: frame pooling + prototype build + all-pairs scoring.
 * ---------------------------------------------------------------------------------------------- */

/* Replaces MeanPooler.forward (model/poolers.py:13-16): out[n,:] = mean_l in[n*L+l,:].  */
int orbit_pool_clips(const float* frame_feats, int num_clips, int clip_length, int feat_dim,
                     float* clip_feats, void* stream);

/* Sliding-window dedupe: clip features of attach_frame_history(frames, L) + MeanPooler
 * (data/utils.py:8-28, model/poolers.py:13-16) from per-frame features computed ONCE:
 *   clip_feats[t,:] = mean over l = 0..L-1 of frame_feats[max(t-L+1+l, 0),:]   (left-padded with frame 0).
 * Bit-identical to orbit_pool_clips over the materialised [F, L] clips; must not run in place.      */
int orbit_pool_history(const float* frame_feats, int num_frames, int history_length, int feat_dim,
                       float* clip_feats, void* stream);

/* Replaces FewShotRecogniser._pool_features + HeadClassifier._build_class_reps +
 * PrototypicalClassifier.configure (few_shot_recognisers.py:155-166; classifier_heads.py:94-105,
 * 232-263) in one launch.
 *   frame_feats [num_clips*clip_length, feat_dim]   support FRAME features (pooling is fused)
 *   class_index [num_clips] int32 in [0,num_classes): rank of the clip's label among the sorted
 *               distinct labels (torch.unique order, classifier_heads.py:96,246-248)
 *   weight [num_classes, feat_dim] <- 2*mu_c ; bias [num_classes] <- -mu_c.mu_c (euclidean only,
 *               may be NULL for cosine); proto [num_classes, feat_dim] <- mu_c (nullable)
 *   scratch: >= orbit_proto_configure_scratch_bytes(num_classes, feat_dim) bytes, ZERO-FILLED once by the
 *               caller before first use (the kernel restores it to zero).                        */
int64_t orbit_proto_configure_scratch_bytes(int num_classes, int feat_dim);
int orbit_proto_configure(const float* frame_feats, const int32_t* class_index, int num_clips,
                          int clip_length, int feat_dim, int num_classes, int metric,
                          float* weight, float* bias, float* proto, void* scratch, void* stream);

/* Replaces _pool_features + PrototypicalClassifier.predict (classifier_heads.py:202-230), and also
 * LinearClassifier.predict / VersaClassifier.predict (classifier_heads.py:63-75,137-143) which are
 * the same s*(qW^T+b) form.
 *   euclidean: logits = logit_scale * (q W^T + b)          (classifier_heads.py:213)
 *   cosine   : logits = logit_scale * cos(q, W_c), each norm clamped at 1e-8 (:215-217)
 *   logits [num_clips, num_classes]; argmax [num_clips] int32 (nullable; first maximal column,
 *   as torch.argmax)                                                                              */
int orbit_head_predict(const float* frame_feats, int num_clips, int clip_length, int feat_dim,
                       const float* weight, const float* bias, int num_classes, int metric,
                       float logit_scale, float* logits, int32_t* argmax, void* stream);

/* FiLM parameter generator in ONE launch. Replaces FilmParameterGenerator.forward (model/feature_adapters.py:66-78):
 * per FiLM tensor i (sorted-name order) g = Linear(hidden,size_i)(ReLU(LayerNorm(Linear(hidden,hidden)(z)))),
 * gamma' = gamma0*(g*r+1) for '...weight', beta' = beta0 + g*r for '...bias'; written at film[out_i ...].
 *   gen_params: flat fp32 blob of all generators; table: num_tensors device records of
 *   orbit_film_table_entry_bytes() bytes = {int64 w1,b1,ln_w,ln_b,w2,b2,reg,init,out; int32 size,is_weight}
 *   (offsets into gen_params / film); task_embedding z [hidden] (set_encoders.py:61-75 aggregate).         */
int orbit_film_table_entry_bytes(void);
int orbit_film_generate(const float* gen_params, const void* table, int num_tensors, int max_size,
                        const float* task_embedding, int hidden, float* film, void* stream);

/* out[r, o] = act(in[r, :] . weight[o, :] + bias[o]) (+ skip[r, o]) for r < rows <= 64. act: 0 none, 2 ReLU,
 * 3 ELU. The dense layers of DenseResidualBlock (model/mlps.py:33-50) used by VersaClassifier.configure
 * (classifier_heads.py:171-180) on the class-mean rows.                                                   */
int orbit_dense_rows(const float* in, const float* weight, const float* bias, const float* skip, float* out,
                     int rows, int in_dim, int out_dim, int act, void* stream);

/* Mahalanobis head (Simple CNAPs). Replaces MahalanobisClassifier.configure/predict/_estimate_cov
 * (classifier_heads.py:282-368): Sigma_c = n_c/(n_c+1) cov(class c) + 1/(n_c+1) cov(all) + I, P_c = Sigma_c^-1,
 * logits[n,c] = -logit_scale (mu_c - q_n)^T P_c (mu_c - q_n); a one-clip class uses the reference's scalar estimate.
 *   clip_feats [num_clips, feat_dim] pooled features; order [num_clips] int32 (device): clip indices grouped by class
 *   rank; counts_host [num_classes] (HOST): clips per class. means [C,D], precisions [C,D,D], task_mean [D],
 *   task_precision [D,D] = (cov(all) + I)^-1 (place it at precisions + C*D*D to invert all matrices in one sweep).  */
int64_t orbit_mahalanobis_configure_workspace_bytes(int num_clips, int feat_dim, int num_classes);
int orbit_mahalanobis_configure(const float* clip_feats, const int32_t* order, const int32_t* counts_host, int num_clips,
                                int feat_dim, int num_classes, float* means, float* precisions, float* task_mean,
                                float* task_precision, void* workspace, void* stream);
int64_t orbit_mahalanobis_predict_workspace_bytes(int num_clips, int feat_dim);
int orbit_mahalanobis_predict(const float* clip_feats, int num_clips, int feat_dim, const float* means,
                              const float* precisions, int num_classes, float logit_scale, float* logits,
                              void* workspace, void* stream);

/* FineTuner inner loop in one launch. Replaces the num_grad_steps x batches loop of
 * MultiStepFewShotRecogniser.personalise (few_shot_recognisers.py:231-246) for the default FineTuner (frozen
 * extractor, so the clip features are loop-invariant): LinearClassifier.predict + cross_entropy (mean, each
 * batch re-weighted by batch_len/N  ==  mean over all N clips) + torch.optim.Adam / SGD (utils/optim.py:8-32).
 *   clip_feats [num_clips, feat_dim] pooled support features; class_index as in orbit_proto_configure
 *   optimizer 0 = Adam(lr, betas, eps, weight_decay), 1 = SGD(lr, momentum, weight_decay)
 *   weight [num_classes, feat_dim], bias [num_classes]: in = initial head (zeros, classifier_heads.py:59-60),
 *   out = personalised head.  scratch >= orbit_linear_finetune_scratch_bytes() (optimiser moments, the gradient of the
 *   logits and, for the cooperative-grid kernel that runs when feat_dim % 16 == 0, the per-CTA partial logits).           */
int64_t orbit_linear_finetune_scratch_bytes(int num_clips, int feat_dim, int num_classes);
int orbit_linear_finetune(const float* clip_feats, const int32_t* class_index, int num_clips, int feat_dim,
                          int num_classes, int num_grad_steps, int optimizer, float lr, float beta1, float beta2,
                          float eps, float weight_decay, float momentum, float logit_scale, float* weight,
                          float* bias, void* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Single layer entry point (what the engine runs for every conv_pw / conv_pwl / conv_head + BatchNormAct2d
 * of the timm extractor, few_shot_recognisers.py:114-117): 1x1 convolution on NHWC activations as a GEMM,
 *   out[M,N] = act( (A[M,K] * gate[m / rows_per_frame, k]) . W[N,K]^T * scale[n] + shift[n] ) (+ residual[M,N])
 * act: 0 none, 1 SiLU, 2 ReLU.  gate / residual may be NULL.
 * mode 0: fp32 FFMA tiles; mode 1: tcgen05 3xTF32 (fp32-grade); mode 2: tcgen05 single-pass TF32.
 * w_split: scratch of 2*N*K floats (modes 1,2; receives the tf32 hi/lo split of W), may be NULL in mode 0.
 * ---------------------------------------------------------------------------------------------- */
int orbit_pointwise_conv(const float* A, const float* W, const float* scale, const float* shift,
                         const float* gate, const float* residual, float* out, int M, int N, int K,
                         int rows_per_frame, int act, int mode, float* w_split, void* stream);

/* 3x3 convolution, stride 1, padding 1, on NHWC activations with the folded BatchNorm scale-shift, activation (0 none, 1 SiLU,
 * 2 ReLU, 18 = ReLU after the residual add) and residual fused: the BasicBlock convolutions of the resnet18 extension
 * (BASELINE.json configs 1 and 3), SimplePrePoolNet layers 2-5 of the set encoder (model/set_encoders.py:91-105) and the
 * EdgeResidual expand of tf_efficientnetv2_s. weight [Cout,Cin,3,3] (torch layout).
 *   implicit = 1: implicit GEMM on the tcgen05 kernel (Cin % 64 == 0): no im2col matrix, the nine taps are row-shifted TMA
 *                 boxes of the activation; implicit = 0: explicit im2col + the same GEMM (any Cin % 4 == 0). Same arithmetic.
 *   scratch: orbit_conv3x3_scratch_floats(...) floats.                                                                     */
int64_t orbit_conv3x3_scratch_floats(int B, int H, int W, int Cin, int Cout, int implicit);
int orbit_conv3x3(const float* x, const float* weight, const float* scale, const float* shift, const float* residual,
                  float* out, int B, int H, int W, int Cin, int Cout, int act, int implicit, float* scratch,
                  int64_t scratch_floats, void* stream);

/* First convolution on the 3-channel NCHW frames, 64 output channels, folded BatchNorm scale-shift (+ ReLU: act 2) fused, NHWC
 * output [B,Ho,Wo,64]: SimplePrePoolNet.layer1 of the set encoder (k 3, stride 1, pad 1; model/set_encoders.py:91-105) and conv1 of
 * the resnet18 extension (k 7, stride 2, pad 3). Direct tensor-core convolution, no im2col matrix. weight [64,3,k,k] (torch).
 * Other geometries: ORBIT_ERR_UNSUPPORTED.                                                                                  */
int orbit_conv_first(const float* x, const float* weight, const float* scale, const float* shift, float* y, int B, int H,
                     int W, int k, int stride, int pad, int act, void* stream);

/* Squeeze-excite gate of one MBConv block: timm SqueezeExcite (conv_reduce + SiLU + conv_expand + sigmoid on the spatial mean),
 * inside the extractor invoked at model/few_shot_recognisers.py:114-117,143-146.
 *   partial [B][groups][C]: per-group channel sums of the depthwise output (orbit_depthwise_conv), hw = pixels per frame;
 *   w1 [R,C] (conv_reduce.weight), b1 [R], w2t [R,C] (conv_expand.weight TRANSPOSED), b2 [C]  ->  gate [B,C] in (0,1).
 * The weight matrices stream through shared memory (cp.async.bulk ring) when C % 4 == 0 and every tensor is 16-byte aligned.  */
int orbit_se_gate(const float* partial, int groups, int hw, const float* w1, const float* b1, const float* w2t, const float* b2,
                  float* gate, int B, int C, int R, void* stream);

/* EfficientNet stem: timm conv_stem (3x3, stride 2, TF "SAME" padding, 3 -> 32) + bn1 (folded scale / shift) + activation on the
 * NCHW fp32 frames the recogniser is handed (model/few_shot_recognisers.py:114-117,143-146).
 *   x [B,3,H,W] -> y [B,ceil(H/2),ceil(W/2),32] NHWC; weight [32,3,3,3] (torch); act: 0 none, 1 SiLU, 2 ReLU.
 * fp32 FMA kernel (stem_kernel). A tensor-core variant (FP16x3, the conv_first scheme with 27 of 48 k slots used) measured
 * 1,597 us per 1,600 frames against 1,322 us and was not kept.                                                                    */
int orbit_stem_conv(const float* x, const float* weight, const float* scale, const float* shift, float* y, int B, int H, int W,
                    int act, void* stream);

/* Depthwise k x k convolution (k in {3,5}, stride in {1,2}, TF "SAME" padding) on NHWC activations with the folded
 * BatchNorm/FiLM scale-shift and activation fused: timm conv_dw + BatchNormAct2d of every MBConv block.
 *   x [B,H,W,C] -> y [B,ceil(H/s),ceil(W/s),C]; weight [C,1,k,k] (torch layout); weight_scratch: k*k*C floats.
 *   partial (nullable): [B][groups][C] per-block channel sums of y (the squeeze-excite squeeze), groups*C*B =
 *   orbit_depthwise_partial_floats().                                                                    */
int64_t orbit_depthwise_partial_floats(int B, int H, int W, int C, int k, int stride);
int orbit_depthwise_conv(const float* x, const float* weight, const float* scale, const float* shift, float* y,
                         float* partial, float* weight_scratch, int B, int H, int W, int C, int k, int stride,
                         int act, void* stream);

/* Fused front half of a timm InvertedResidual block with 16 or 24 input channels (EfficientNet-B0 blocks 1.0, 1.1, 2.0):
 * conv_pw 1x1 + bn1 + SiLU -> conv_dw kxk (TF-SAME) + bn2 (FiLM site, model/film.py:43-44) + SiLU, one launch; the
 * 6x-expanded tensor stays in shared memory. Same result as orbit_pointwise_conv (mode 0) followed by
 * orbit_depthwise_conv up to fp32 summation order.
 *   x [B,H,W,Cin] NHWC; w_expand [C,Cin]; scale1/shift1 [C] folded bn1; w_dw [C,1,k,k]; scale2/shift2 [C] folded bn2;
 *   y [B,ceil(H/s),ceil(W/s),C]; partial: orbit_mbconv_partial_floats(...) floats (SE squeeze sums per block; nullable);
 *   weight_scratch: k*k*C floats.                                                                                    */
int64_t orbit_mbconv_partial_floats(int B, int H, int W, int C, int k, int stride);   /* upper bound over Cin in {16, 24} */
/* the partial sums are laid out [B][groups][C] with groups = orbit_mbconv_partial_groups(...) (depends on the kernel variant
 * the shape selects: the register-resident 3x3 stride-2 kernel for 16 input channels writes one group per band of rows)      */
int orbit_mbconv_partial_groups(int H, int W, int Cin, int C, int k, int stride);
int orbit_mbconv_expand_dw(const float* x, const float* w_expand, const float* scale1, const float* shift1,
                           const float* w_dw, const float* scale2, const float* shift2, float* y, float* partial,
                           float* weight_scratch, int B, int H, int W, int Cin, int C, int k, int stride, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backbone engine: the feature extractor forward (reference: timm model called at
 * few_shot_recognisers.py:114-117,143-146, with FiLM by functional_call parameter substitution).
 * ---------------------------------------------------------------------------------------------- */
typedef struct orbit_engine orbit_engine;

int  orbit_engine_create(orbit_engine** out, int arch);
void orbit_engine_destroy(orbit_engine* e);
int  orbit_engine_feat_dim(const orbit_engine* e);

/* Parameter blob layout. The caller packs the extractor's state_dict (timm key names, e.g.
 * "blocks.1.0.bn2.weight") into ONE fp32 device array; entry i lives at [offset, offset+numel). */
int     orbit_engine_num_params(const orbit_engine* e);
int     orbit_engine_param_info(const orbit_engine* e, int i, char* name, int name_cap,
                                int64_t* numel, int64_t* offset);
int     orbit_engine_param_shape(const orbit_engine* e, int i, int* ndim, int64_t* dims4);
int64_t orbit_engine_param_floats(const orbit_engine* e);

/* FiLM tensors (reference model/film.py:38-74), in the SORTED-name order the reference's generator
 * uses (feature_adapters.py:43-44); the film blob is their concatenation.                        */
int     orbit_engine_num_film(const orbit_engine* e);
int     orbit_engine_film_info(const orbit_engine* e, int i, char* name, int name_cap,
                               int64_t* numel, int64_t* offset);
int64_t orbit_engine_film_floats(const orbit_engine* e);

/* Derived per-task tensors (BatchNorm folded to per-channel scale/shift with the FiLM gamma'/beta'
 * substituted, re-laid-out depthwise weights, tf32 hi/lo weight splits). `film` may be NULL
 * (no adaptation). Must be re-run when params or film change.                                    */
int64_t orbit_engine_derived_floats(const orbit_engine* e);
int     orbit_engine_prepare(const orbit_engine* e, const float* params, const float* film,
                             float* derived, void* stream);

/* Options: "chunk_frames" (frames per pass, sized to keep inter-layer tensors in L2),
 *          "gemm" 0 = fp32 FFMA tiles, 1 = tcgen05 3xTF32 (fp32-accurate), 2 = tcgen05 1xTF32,
 *          "profile" 0/1 = CUDA-event timing of every launch (see orbit_engine_profile_read)      */
int orbit_engine_set_option(orbit_engine* e, const char* key, int value);
int orbit_engine_get_option(const orbit_engine* e, const char* key, int* value);

int64_t orbit_engine_workspace_bytes(const orbit_engine* e, int height, int width);

/* Multiply-accumulates of one forward pass over ONE height x width frame (convolutions, dense layers, attention
 * matmuls, pooling adds; normalisation and activations are not counted). Replaces the thop trace behind
 * OpsCounter.compute_macs (reference utils/ops_counter.py:82-88, few_shot_recognisers.py:119-120). Host-only; < 0 on error. */
int64_t orbit_engine_macs(const orbit_engine* e, int height, int width);

/* frames [num_frames,3,height,width] fp32 NCHW  ->  feats [num_frames, feat_dim] fp32.          */
int orbit_engine_forward(const orbit_engine* e, const float* params, const float* derived,
                         const float* frames, int num_frames, int height, int width,
                         float* feats, void* workspace, int64_t workspace_bytes, void* stream);

/* Batch-statistics pass (the train-mode BatchNorm of few_shot_recognisers.py:181-183, used here to
 * calibrate synthetic checkpoints): one forward over `frames` (num_frames <= chunk_frames) in which every
 * BatchNorm normalises with the statistics of this batch; the batch mean / unbiased variance are WRITTEN
 * into the running_mean / running_var entries of `params`, and `derived` is refolded accordingly.       */
int orbit_engine_calibrate(const orbit_engine* e, float* params, float* derived, const float* frames,
                           int num_frames, int height, int width, float* feats, void* workspace,
                           int64_t workspace_bytes, void* stream);

/* Per-kernel-family timing. With option "profile"=1 every launch of orbit_engine_forward is bracketed by
 * CUDA events on the caller's stream (the only place the library creates CUDA objects). profile_read blocks
 * until those launches finished and returns, per family, the summed device time [ms], launch count and the
 * ALGORITHMIC bytes / flops (each input and output tensor counted once) since the previous read.
 * Families: 0 stem conv, 1 depthwise conv, 2 squeeze-excite gate, 3 pointwise-conv GEMM, 4 spatial mean,
 * 5 calibration statistics.                                                                           */
#define ORBIT_PROFILE_FAMILIES 6
int orbit_engine_profile_read(const orbit_engine* e, double* ms, int64_t* launches, double* bytes, double* flops);

/* number of kernels the last orbit_engine_forward on this engine enqueued (for bench accounting) */
int64_t orbit_engine_last_launches(const orbit_engine* e);

/* ---------------------------------------------------------------------------------------------
 * Device-side evaluator. Replaces the per-video statistics of the reference's
 * Evaluator.get_frame_accuracy / get_frames_to_recognition / get_video_prediction and the softmax + .cpu() of
 * TestEvaluator.append_video (utils/eval_metrics.py:27-68,260-276) for a batch of videos in one launch.
 *   logits [rows, num_classes] fp32 (nullable) or predictions [rows] int32 (used when logits is NULL);
 *   frame_index [video_offsets[V]] int32 (nullable): row of every scored frame, in order (the reference drops padded
 *   duplicate frames with np.unique, eval_metrics.py:262-264); NULL = rows video_offsets[v] .. video_offsets[v+1]-1;
 *   video_offsets [V+1], video_labels [V] int32 (device);
 *   stats [V,4] int32 = {correct frames, frames, index of first correct frame (= frames if none), most frequent
 *   prediction (lowest index on ties)}; pred_out [video_offsets[V]] int32 (nullable): arg-max per scored frame.
 * num_classes <= 64.                                                                               */
int orbit_video_stats(const float* logits, const int32_t* predictions, int num_classes, const int32_t* frame_index,
                      const int32_t* video_offsets, const int32_t* video_labels, int num_videos, int32_t* stats,
                      int32_t* pred_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training through the frozen extractor: FineTuner + FiLM, i.e. MultiStepFewShotRecogniser.personalise
 * with adapt_features=True (model/few_shot_recognisers.py:196-198,207-246), and CNAPs-style meta-training of the set
 * encoder + FiLM generator (single-step-learner.py:196-243, few_shot_recognisers.py:313-343,388-437).
 * ORBIT_ARCH_EFFICIENTNET_B0 (gradients of the FiLM-site BatchNorm weight / bias) and ORBIT_ARCH_SET_ENCODER (gradients
 * of every parameter; `frames` = the input of the forward pass, needed for the first conv's weight gradient; `tderived`
 * is rewritten from the current weights on every call); other architectures return ORBIT_ERR_UNSUPPORTED.
 *   orbit_engine_train_saved_floats   floats PER FRAME of the activation arena forward_train fills
 *   orbit_engine_train_derived_floats floats of the transposed / split 1x1 weights (prepare_train; weights are frozen)
 *   orbit_engine_forward_train        same result as orbit_engine_forward (layer at a time, BatchNorm in eval mode) for
 *                                     num_frames <= chunk_frames, keeping the pre-activations in `saved`
 *   orbit_engine_backward_train       dfeats [num_frames, feat_dim] -> ACCUMULATES d loss / d (BatchNorm weight, bias) of
 *                                     every FiLM site into grad_params, a blob with the layout of `params`
 *   orbit_linear_ce_backward          linear head + cross entropy (utils/optim.py:8-9; mean over the batch times
 *                                     loss_scale = batch_len / context_size, few_shot_recognisers.py:241-243): accumulates
 *                                     grad_weight [C,D] and grad_bias [C], writes grad_frame_feats [num_clips*L, D];
 *                                     labels int32 in [0,C); scratch: orbit_linear_ce_scratch_floats(...) floats          */
int64_t orbit_engine_train_saved_floats(const orbit_engine* engine, int height, int width);
int64_t orbit_engine_train_derived_floats(const orbit_engine* engine);
int orbit_engine_prepare_train(const orbit_engine* engine, const float* params, float* tderived, void* stream);
int orbit_engine_forward_train(const orbit_engine* engine, const float* params, const float* derived, const float* frames,
                               int num_frames, int height, int width, float* feats, float* saved, int64_t saved_floats,
                               void* workspace, int64_t workspace_bytes, void* stream);
int orbit_engine_backward_train(const orbit_engine* engine, const float* params, const float* derived, float* tderived,
                                const float* saved, const float* frames, const float* dfeats, int num_frames, int height, int width,
                                float* grad_params, void* workspace, int64_t workspace_bytes, void* stream);
/* copies the FiLM-site gradients out of a grad_params blob into the film-blob layout (orbit_engine_film_info order):
 * d loss / d (generated gamma', beta') for the generator backward                                                     */
int orbit_engine_film_grad(const orbit_engine* engine, const float* grad_params, float* grad_film, void* stream);
/* FilmParameterGenerator backward (model/feature_adapters.py:66-78): grad_film [film floats] -> grad_gen_params (layout of
 * gen_params, WRITTEN for every trainable tensor of the table) and grad_embedding [hidden]; scratch: num_tensors * hidden */
int orbit_film_generate_backward(const float* gen_params, const void* table, int num_tensors, const float* task_embedding,
                                 int hidden, const float* grad_film, float* grad_gen_params, float* grad_embedding,
                                 float* scratch, void* stream);
/* query path of the linear-form heads (linear / versa / proto: metric EUCLIDEAN -> logits = s (q W^T + b); proto_cosine:
 * metric COSINE), classifier_heads.py:60-79,161-180,202-230: grad_logits [num_clips, C] -> grad_frame_feats [num_clips*L, D]
 * through the clip mean-pool (poolers.py:13-16)                                                                        */
int orbit_head_predict_backward(const float* frame_feats, const float* weight, const float* grad_logits, int num_clips,
                                int clip_length, int feat_dim, int num_classes, int metric, float logit_scale,
                                float* grad_frame_feats, void* stream);
/* the same for the Mahalanobis head (classifier_heads.py:328-350): logits = -s (q - mu_c)^T P_c (q - mu_c)              */
int orbit_mahalanobis_predict_backward(const float* frame_feats, const float* means, const float* precisions,
                                       const float* grad_logits, int num_clips, int clip_length, int feat_dim,
                                       int num_classes, float logit_scale, float* grad_frame_feats, void* stream);
int64_t orbit_linear_ce_scratch_floats(int num_clips, int feat_dim, int num_classes);
int orbit_linear_ce_backward(const float* frame_feats, const int32_t* labels, const float* weight, const float* bias,
                             int num_clips, int clip_length, int feat_dim, int num_classes, float logit_scale,
                             float loss_scale, float* grad_weight, float* grad_bias, float* grad_frame_feats,
                             float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ORBIT_B200_H */
