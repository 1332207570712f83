// FineTuner inner loop: num_grad_steps of Adam/SGD on a linear head over FIXED clip features, in ONE launch.
//
// Replaces the loop of MultiStepFewShotRecogniser.personalise (model/few_shot_recognisers.py:231-246) with
// LinearClassifier.predict (model/classifier_heads.py:63-75), cross_entropy (utils/optim.py:8-9) and
// torch.optim.Adam / SGD as configured by init_optimizer (utils/optim.py:11-32) for the default FineTuner
// (frozen extractor => the support features are loop-invariant).
//
// Per grad step the reference accumulates, over the support batches, d/dtheta of
//     sum_batches (batch_len / N) * mean_{i in batch} CE(s * (W x_i + b), y_i)  =  (1/N) sum_i CE_i ,
// i.e. the batching only changes the fp32 summation order. So one step is
//     G[i,c] = (s / N) * (softmax(s (W x_i + b))_c - [y_i == c])
//     dW = G^T X ,  db = sum_i G[i,:]          then ONE optimiser update.
// A single persistent CTA runs all steps (the problem is a few MFLOP per step; launch latency of
// 50 x ~10 kernels is what the reference pays). Deterministic: fixed summation order, no atomics.
#include "common.cuh"

namespace orbit {

struct FinetuneParams {
    const float* x;          // [N, D] clip features
    const int32_t* y;        // [N] class index in [0, C)
    float* w;                // [C, D] in/out
    float* b;                // [C]    in/out
    float* g;                // scratch [N, C]
    float* mw; float* vw;    // scratch [C, D] each (Adam moments / SGD momentum buffer in mw)
    float* mb; float* vb;    // scratch [C] each
    int N, D, C, steps, optimizer;
    float lr, beta1, beta2, eps, weight_decay, momentum, logit_scale;
};

__global__ void __launch_bounds__(1024, 1) linear_finetune_kernel(const FinetuneParams p) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    const int CD = p.C * p.D;
    for (int i = tid; i < CD; i += blockDim.x) { p.mw[i] = 0.f; p.vw[i] = 0.f; }
    for (int i = tid; i < p.C; i += blockDim.x) { p.mb[i] = 0.f; p.vb[i] = 0.f; }
    __syncthreads();
    const float inv_n = p.logit_scale / (float)p.N;
    float b1t = 1.f, b2t = 1.f;
    for (int step = 1; step <= p.steps; ++step) {
        // ---- phase A: logits -> softmax -> G (one warp per sample; lane c (+32) owns class c) ----
        for (int i = warp; i < p.N; i += n_warps) {
            const float* xi = p.x + (int64_t)i * p.D;
            float l0 = -INFINITY, l1 = -INFINITY;
            for (int c = 0; c < p.C; ++c) {
                const float* wc = p.w + (int64_t)c * p.D;
                float s = 0.f;
                for (int d = lane * 4; d < p.D; d += 128) {
                    const float4 a = *reinterpret_cast<const float4*>(xi + d);
                    const float4 w = *reinterpret_cast<const float4*>(wc + d);
                    s = fmaf(a.x, w.x, s); s = fmaf(a.y, w.y, s); s = fmaf(a.z, w.z, s); s = fmaf(a.w, w.w, s);
                }
                s = p.logit_scale * (warp_sum(s) + p.b[c]);
                if (lane == (c & 31)) { if (c < 32) l0 = s; else l1 = s; }
            }
            float mx = fmaxf(l0, l1);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float e0 = lane < p.C ? expf(l0 - mx) : 0.f;
            const float e1 = lane + 32 < p.C ? expf(l1 - mx) : 0.f;
            const float z = warp_sum(e0 + e1);
            const int yi = p.y[i];
            if (lane < p.C) p.g[(int64_t)i * p.C + lane] = inv_n * (e0 / z - (yi == lane ? 1.f : 0.f));
            if (lane + 32 < p.C) p.g[(int64_t)i * p.C + lane + 32] = inv_n * (e1 / z - (yi == lane + 32 ? 1.f : 0.f));
        }
        __syncthreads();
        // ---- phase B: gradient + optimiser update, one thread per parameter ----
        b1t *= p.beta1; b2t *= p.beta2;
        const float bc1 = 1.f - b1t, bc2_sqrt = sqrtf(1.f - b2t);
        for (int e = tid; e < CD + p.C; e += blockDim.x) {
            const bool is_bias = e >= CD;
            const int c = is_bias ? e - CD : e / p.D, d = is_bias ? 0 : e % p.D;
            float grad = 0.f;
            if (is_bias) for (int i = 0; i < p.N; ++i) grad += p.g[(int64_t)i * p.C + c];
            else for (int i = 0; i < p.N; ++i) grad = fmaf(p.g[(int64_t)i * p.C + c], p.x[(int64_t)i * p.D + d], grad);
            float* param = is_bias ? p.b + c : p.w + e;
            float* m = is_bias ? p.mb + c : p.mw + e;
            float* v = is_bias ? p.vb + c : p.vw + e;
            float w = *param;
            if (p.weight_decay != 0.f) grad = fmaf(p.weight_decay, w, grad);
            if (p.optimizer == 0) {   // torch.optim.Adam (no amsgrad)
                const float mm = p.beta1 * *m + (1.f - p.beta1) * grad;
                const float vv = p.beta2 * *v + (1.f - p.beta2) * grad * grad;
                *m = mm; *v = vv;
                w -= (p.lr / bc1) * (mm / (sqrtf(vv) / bc2_sqrt + p.eps));
            } else {                  // torch.optim.SGD with momentum (dampening 0, no nesterov)
                float buf = grad;
                if (p.momentum != 0.f) { buf = step == 1 ? grad : p.momentum * *m + grad; *m = buf; }
                w -= p.lr * buf;
            }
            *param = w;
        }
        __syncthreads();
    }
}

}  // namespace orbit

using namespace orbit;

extern "C" int64_t orbit_linear_finetune_scratch_bytes(int num_clips, int feat_dim, int num_classes) {
    if (num_clips <= 0 || feat_dim <= 0 || num_classes <= 0) return 0;
    return (int64_t)sizeof(float) * ((int64_t)num_clips * num_classes + 2LL * num_classes * feat_dim + 2LL * num_classes + 16);
}

extern "C" int orbit_linear_finetune(const float* clip_feats, const int32_t* class_index, int num_clips, int feat_dim,
                                     int num_classes, int num_grad_steps, int optimizer, float lr, float beta1, float beta2,
                                     float eps, float weight_decay, float momentum, float logit_scale, float* weight,
                                     float* bias, void* scratch, void* stream) {
    if (!clip_feats || !class_index || !weight || !bias || !scratch) return ORBIT_ERR_ARG;
    if (num_clips <= 0 || feat_dim <= 0 || num_classes <= 0 || num_grad_steps < 0) return ORBIT_ERR_ARG;
    if (optimizer != 0 && optimizer != 1) return ORBIT_ERR_ARG;
    if (num_classes > ORBIT_MAX_CLASSES || feat_dim % 4) return ORBIT_ERR_UNSUPPORTED;
    if (!aligned16(clip_feats) || !aligned16(weight) || !aligned16(scratch)) return ORBIT_ERR_UNSUPPORTED;
    FinetuneParams p;
    p.x = clip_feats; p.y = class_index; p.w = weight; p.b = bias;
    float* s = reinterpret_cast<float*>(scratch);
    const int64_t cd = (int64_t)num_classes * feat_dim;
    p.mw = s; p.vw = s + cd; p.mb = s + 2 * cd; p.vb = p.mb + num_classes; p.g = p.vb + num_classes;
    p.N = num_clips; p.D = feat_dim; p.C = num_classes; p.steps = num_grad_steps; p.optimizer = optimizer;
    p.lr = lr; p.beta1 = beta1; p.beta2 = beta2; p.eps = eps; p.weight_decay = weight_decay; p.momentum = momentum;
    p.logit_scale = logit_scale;
    linear_finetune_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}
