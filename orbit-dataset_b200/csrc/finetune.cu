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
// All steps run inside ONE launch (the problem is a few MFLOP per step; launch latency of 50 x ~10 kernels is what the
// reference pays). Deterministic: fixed summation order, no atomics.
//   linear_finetune_grid_kernel (default): a cooperative grid of D / 16 CTAs; each owns 16 feature columns -- its slice of X,
//     W and the optimiser moments stay in shared memory for the whole loop -- and per step publishes its partial logits
//     [N, C], meets the others at ONE grid barrier, sums the partials in CTA order, and updates its own 16 C parameters (the
//     bias redundantly in every CTA). r02c profile of the S4 episode: the single-CTA kernel below took 3.29 ms of 17.6 ms
//     (66 us per step, a chain of L2 latencies: X is 245 KB, W changes every step).
//   linear_finetune_kernel: one CTA, everything through L1 / L2; serves D % 16 != 0 and shapes beyond the grid kernel's
//     shared memory.
#include <cooperative_groups.h>

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

// one optimiser update (shared by both kernels): returns the new parameter value
__device__ __forceinline__ float finetune_update(const FinetuneParams& p, float w, float grad, float& m, float& v, int step, float bc1,
                                                 float bc2_sqrt) {
    if (p.weight_decay != 0.f) grad = fmaf(p.weight_decay, w, grad);
    if (p.optimizer == 0) {   // torch.optim.Adam (no amsgrad)
        const float mm = p.beta1 * m + (1.f - p.beta1) * grad;
        const float vv = p.beta2 * v + (1.f - p.beta2) * grad * grad;
        m = mm; v = vv;
        return w - (p.lr / bc1) * (mm / (sqrtf(vv) / bc2_sqrt + p.eps));
    }
    float buf = grad;         // torch.optim.SGD with momentum (dampening 0, no nesterov)
    if (p.momentum != 0.f) { buf = step == 1 ? grad : p.momentum * m + grad; m = buf; }
    return w - p.lr * buf;
}

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
            *param = finetune_update(p, *param, grad, is_bias ? p.mb[c] : p.mw[e], is_bias ? p.vb[c] : p.vw[e], step, bc1, bc2_sqrt);
        }
        __syncthreads();
    }
}

constexpr int kFtCols = 16, kFtThreads = 256;

// grid = D / 16 CTAs (cooperative launch). part: [2][grid][N * C] partial logits (double-buffered across steps), then
// [2][N * C] totals (two_level).
// shared: xs [N][16] | ws, mws, vws [C][16] each | bs, mbs, vbs [C] each | lg [N][C]
__global__ void __launch_bounds__(kFtThreads) linear_finetune_grid_kernel(const FinetuneParams p, float* __restrict__ part, int two_level) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) float s_ft[];
    const int N = p.N, C = p.C, NC = N * C, CK = C * kFtCols;
    float* xs = s_ft;
    float* ws = xs + (size_t)N * kFtCols;
    float* mws = ws + CK;
    float* vws = mws + CK;
    float* bs = vws + CK;
    float* mbs = bs + C;
    float* vbs = mbs + C;
    float* lg = vbs + C;
    const int tid = threadIdx.x, d0 = blockIdx.x * kFtCols, ncta = gridDim.x;
    for (int e = tid; e < N * (kFtCols / 4); e += kFtThreads) {
        const int i = e / (kFtCols / 4), q = e % (kFtCols / 4);
        reinterpret_cast<float4*>(xs)[e] = *reinterpret_cast<const float4*>(p.x + (int64_t)i * p.D + d0 + 4 * q);
    }
    for (int e = tid; e < CK; e += kFtThreads) {
        ws[e] = p.w[(int64_t)(e / kFtCols) * p.D + d0 + e % kFtCols];
        mws[e] = 0.f; vws[e] = 0.f;
    }
    for (int e = tid; e < C; e += kFtThreads) { bs[e] = p.b[e]; mbs[e] = 0.f; vbs[e] = 0.f; }
    __syncthreads();
    const float inv_n = p.logit_scale / (float)N;
    float b1t = 1.f, b2t = 1.f;
    for (int step = 1; step <= p.steps; ++step) {
        float* mine = part + ((size_t)(step & 1) * ncta + blockIdx.x) * NC;
        for (int e = tid; e < NC; e += kFtThreads) {          // partial logits of this CTA's 16 columns
            const int i = e / C, c = e - i * C;
            const float4* a = reinterpret_cast<const float4*>(xs + (size_t)i * kFtCols);
            const float4* w = reinterpret_cast<const float4*>(ws + c * kFtCols);
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < kFtCols / 4; ++q) {
                const float4 av = a[q], wv = w[q];
                s = fmaf(av.x, wv.x, s); s = fmaf(av.y, wv.y, s); s = fmaf(av.z, wv.z, s); s = fmaf(av.w, wv.w, s);
            }
            __stcg(mine + e, s);
        }
        grid.sync();                                           // (release / acquire of the partials included)
        const float* all = part + (size_t)(step & 1) * ncta * NC;
        if (!two_level) {
            for (int e = tid; e < NC; e += kFtThreads) {      // full logits: the partials in CTA order
                float s = 0.f;
#pragma unroll 16
                for (int j = 0; j < ncta; ++j) s += __ldcg(all + (size_t)j * NC + e);  // (16 independent L2 loads in flight per batch)
                lg[e] = p.logit_scale * (s + bs[e % C]);
            }
        } else {
            // many logits: every CTA summing all of them would read ncta x N x C floats per CTA per step. Each CTA sums a slice
            // (same CTA order), publishes it, and a second barrier later everyone reads the N x C totals.
            float* tot = part + (size_t)2 * ncta * NC + (size_t)(step & 1) * NC;
            const int per = (NC + ncta - 1) / ncta, e0 = blockIdx.x * per, e1 = min(NC, e0 + per);
            for (int e = e0 + tid; e < e1; e += kFtThreads) {
                float s = 0.f;
#pragma unroll 16
                for (int j = 0; j < ncta; ++j) s += __ldcg(all + (size_t)j * NC + e);
                __stcg(tot + e, s);
            }
            grid.sync();
            for (int e = tid; e < NC; e += kFtThreads) lg[e] = p.logit_scale * (__ldcg(tot + e) + bs[e % C]);
        }
        __syncthreads();
        for (int i = tid; i < N; i += kFtThreads) {           // G[i, :] = (scale / N) (softmax - onehot), in place
            float* li = lg + (size_t)i * C;
            float mx = -INFINITY;
            for (int c = 0; c < C; ++c) mx = fmaxf(mx, li[c]);
            float z = 0.f;
            for (int c = 0; c < C; ++c) { const float ev = expf(li[c] - mx); li[c] = ev; z += ev; }
            const int yi = p.y[i];
            for (int c = 0; c < C; ++c) li[c] = inv_n * (li[c] / z - (yi == c ? 1.f : 0.f));
        }
        __syncthreads();
        b1t *= p.beta1; b2t *= p.beta2;
        const float bc1 = 1.f - b1t, bc2_sqrt = sqrtf(1.f - b2t);
        for (int e = tid; e < CK + C; e += kFtThreads) {       // this CTA's 16 C weights, and (in every CTA alike) the bias
            const bool is_bias = e >= CK;
            const int c = is_bias ? e - CK : e / kFtCols, k = is_bias ? 0 : e % kFtCols;
            float grad = 0.f;
            if (is_bias) for (int i = 0; i < N; ++i) grad += lg[(size_t)i * C + c];
            else for (int i = 0; i < N; ++i) grad = fmaf(lg[(size_t)i * C + c], xs[(size_t)i * kFtCols + k], grad);
            float* param = is_bias ? bs + c : ws + e;
            *param = finetune_update(p, *param, grad, is_bias ? mbs[c] : mws[e], is_bias ? vbs[c] : vws[e], step, bc1, bc2_sqrt);
        }
        __syncthreads();
    }
    for (int e = tid; e < CK; e += kFtThreads) p.w[(int64_t)(e / kFtCols) * p.D + d0 + e % kFtCols] = ws[e];
    if (blockIdx.x == 0) for (int e = tid; e < C; e += kFtThreads) p.b[e] = bs[e];
}

}  // namespace orbit

using namespace orbit;

static int g_finetune_grid = 1;      // dev A/B switch (orbit_set_global_option "finetune_grid")
namespace orbit {
void set_finetune_grid(int on) { g_finetune_grid = on; }
int get_finetune_grid() { return g_finetune_grid; }
}

extern "C" int64_t orbit_linear_finetune_scratch_bytes(int num_clips, int feat_dim, int num_classes) {
    if (num_clips <= 0 || feat_dim <= 0 || num_classes <= 0) return 0;
    const int64_t ctas = (feat_dim + kFtCols - 1) / kFtCols;       // + the grid kernel's double-buffered partial logits
    return (int64_t)sizeof(float) * ((int64_t)num_clips * num_classes * (3 + 2 * ctas) + 2LL * num_classes * feat_dim + 2LL * num_classes + 16);
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
    if (g_finetune_grid && feat_dim % kFtCols == 0 && num_grad_steps > 0) {
        const int ctas = feat_dim / kFtCols;
        const size_t smem = sizeof(float) * ((size_t)num_clips * kFtCols + 3u * num_classes * kFtCols + 3u * num_classes +
                                             (size_t)num_clips * num_classes);
        int dev = 0, sms = 0, coop = 0, per_sm = 0;
        ORBIT_CUDA(cudaGetDevice(&dev));
        ORBIT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        ORBIT_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        if (coop && smem <= 200 * 1024) {
            ORBIT_CUDA(cudaFuncSetAttribute(linear_finetune_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ORBIT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, linear_finetune_grid_kernel, kFtThreads, smem));
            if ((int64_t)per_sm * sms >= ctas) {               // the whole grid is resident: the grid barrier cannot deadlock
                float* part = p.g + (int64_t)num_clips * num_classes;
                int two_level = (int64_t)num_clips * num_classes * ctas > 65536 ? 1 : 0;
                void* args[] = {(void*)&p, (void*)&part, (void*)&two_level};
                ORBIT_CUDA(cudaLaunchCooperativeKernel((const void*)linear_finetune_grid_kernel, dim3(ctas), dim3(kFtThreads), args, smem,
                                                       (cudaStream_t)stream));
                return ORBIT_OK;
            }
        }
    }
    linear_finetune_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}
