// Head kernels: clip mean-pooling, prototype build and all-pairs scoring.
//
// Replaces (reference microsoft/ORBIT-Dataset @ 97ccae1):
//   MeanPooler.forward                      model/poolers.py:13-16
//   HeadClassifier._build_class_reps        model/classifier_heads.py:94-105
//   PrototypicalClassifier.configure        model/classifier_heads.py:232-263
//   PrototypicalClassifier.predict          model/classifier_heads.py:202-230
//   LinearClassifier/VersaClassifier.predict  model/classifier_heads.py:63-75,137-143
//
// All of this is HBM-bound streaming work (a few flop per byte): coalesced 128-bit loads, warp
// shuffles and shared memory for the reductions; no tensor cores.
#include "common.cuh"

namespace orbit {

// ------------------------------------------------------------------------------------------------
// pool: out[n,:] = mean_l in[n*L+l,:]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_clips_kernel(const float* __restrict__ in, int n_clips, int L,
                                                         int D, float* __restrict__ out) {
    const int d4 = D >> 2;
    const int64_t total = (int64_t)n_clips * d4;
    const float invL = (float)L;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / d4), q = (int)(i % d4);
        const float* p = in + ((int64_t)n * L) * D + q * 4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < L; ++l) add4(s, ldg4_stream(p + (int64_t)l * D));
        s.x /= invL; s.y /= invL; s.z /= invL; s.w /= invL;
        *reinterpret_cast<float4*>(out + (int64_t)n * D + q * 4) = s;
    }
}

// ------------------------------------------------------------------------------------------------
// causal history pool: out[t,:] = mean_l in[max(t-L+1+l, 0),:], l = 0..L-1 -- the clip features of
// attach_frame_history(frames, L) (reference data/utils.py:8-28) followed by MeanPooler (model/poolers.py:13-16),
// from per-FRAME features computed once. Same summation order and division as pool_clips_kernel, so the result is
// bit-identical to pooling the materialised clips.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_history_kernel(const float* __restrict__ in, int n_frames, int L, int D,
                                                           float* __restrict__ out) {
    const int d4 = D >> 2;
    const int64_t total = (int64_t)n_frames * d4;
    const float len = (float)L;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / d4), q = (int)(i % d4);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < L; ++l) {
            const int src = max(t - L + 1 + l, 0);
            add4(s, ldg4(in + (int64_t)src * D + q * 4));
        }
        s.x /= len; s.y /= len; s.z /= len; s.w /= len;
        *reinterpret_cast<float4*>(out + (int64_t)t * D + q * 4) = s;
    }
}

// ------------------------------------------------------------------------------------------------
// configure: one launch builds mu_c, W = 2 mu_c, b = -mu_c.mu_c from support FRAME features.
// grid (C, ceil(D/128)); block 256 = 8 row groups x 32 lanes, lane owns 4 consecutive columns.
// ------------------------------------------------------------------------------------------------
struct ConfigureScratch {
    int counters[ORBIT_MAX_CLASSES];
    float partial[1];  // [C * n_tiles]
};

__global__ void __launch_bounds__(256)
proto_configure_kernel(const float* __restrict__ feats, const int32_t* __restrict__ class_index, int n_clips,
                       int L, int D, int euclidean, float* __restrict__ weight, float* __restrict__ bias,
                       float* __restrict__ proto, ConfigureScratch* scratch) {
    const int c = blockIdx.x, tile = blockIdx.y, n_tiles = gridDim.y;
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int d0 = tile * 128 + lane * 4;
    const bool live = d0 < D;

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int count = 0;
    const float fL = (float)L;
    for (int n = rg; n < n_clips; n += 8) {
        if (__ldg(class_index + n) != c) continue;  // warp-uniform
        ++count;
        if (live) {
            const float* p = feats + ((int64_t)n * L) * D + d0;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int l = 0; l < L; ++l) add4(s, ldg4_stream(p + (int64_t)l * D));
            acc.x += s.x / fL; acc.y += s.y / fL; acc.z += s.z / fL; acc.w += s.w / fL;
        }
    }
    __shared__ float4 s_acc[8][32];
    __shared__ int s_cnt[8];
    __shared__ int s_last;
    s_acc[rg][lane] = acc;
    if (lane == 0) s_cnt[rg] = count;
    __syncthreads();
    if (rg == 0) {
        int total = 0;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 8; ++r) { add4(t, s_acc[r][lane]); total += s_cnt[r]; }
        const float fn = (float)(total > 0 ? total : 1);
        float4 mu = make_float4(t.x / fn, t.y / fn, t.z / fn, t.w / fn);
        float dot = 0.f;
        if (live) {
            const int64_t o = (int64_t)c * D + d0;
            *reinterpret_cast<float4*>(weight + o) = make_float4(2.f * mu.x, 2.f * mu.y, 2.f * mu.z, 2.f * mu.w);
            if (proto) *reinterpret_cast<float4*>(proto + o) = mu;
            dot = mu.x * mu.x + mu.y * mu.y + mu.z * mu.z + mu.w * mu.w;
        }
        if (euclidean) {
            dot = warp_sum(dot);
            if (lane == 0) {
                scratch->partial[c * n_tiles + tile] = dot;
                __threadfence();
                s_last = (atomicAdd(&scratch->counters[c], 1) == n_tiles - 1);
            }
            __syncwarp();
            if (s_last && lane == 0) {  // last tile of this class: fixed-order sum => deterministic bias
                __threadfence();
                float b = 0.f;
                for (int t2 = 0; t2 < n_tiles; ++t2) b += *((volatile float*)&scratch->partial[c * n_tiles + t2]);
                bias[c] = -b;
                scratch->counters[c] = 0;  // restore the zero state for the next launch
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// predict: one warp per query clip. Streams the clip's L frame rows once (mean-pool on the fly) and
// dots against every class row held in shared memory; writes logits[n, :] and argmax[n].
// ------------------------------------------------------------------------------------------------
constexpr int kPredictGroup = 16;  // classes scored per pass over the query row

template <bool kSmemW>
__global__ void __launch_bounds__(256)
head_predict_kernel(const float* __restrict__ feats, int n_clips, int L, int D, const float* __restrict__ weight,
                    const float* __restrict__ bias, int C, int cosine, float logit_scale,
                    float* __restrict__ logits, int32_t* __restrict__ argmax) {
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                       // [C][D] when kSmemW
    float* s_wn = smem + (kSmemW ? (size_t)C * D : 0);  // [C] row norms (cosine)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;

    if (kSmemW) {
        const int total4 = (C * D) >> 2;
        for (int i = threadIdx.x; i < total4; i += blockDim.x)
            reinterpret_cast<float4*>(s_w)[i] = ldg4(weight + 4 * (int64_t)i);
        __syncthreads();
    }
    if (cosine) {
        for (int c = warp; c < C; c += n_warps) {
            float ss = 0.f;
            for (int d = lane * 4; d < D; d += 128) {
                const float4 w = kSmemW ? *reinterpret_cast<const float4*>(s_w + (size_t)c * D + d)
                                        : ldg4(weight + (int64_t)c * D + d);
                ss += w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
            }
            ss = warp_sum(ss);
            if (lane == 0) s_wn[c] = fmaxf(sqrtf(ss), 1e-8f);
        }
        __syncthreads();
    }

    const float fL = (float)L;
    for (int n = blockIdx.x * n_warps + warp; n < n_clips; n += gridDim.x * n_warps) {
        const float* row = feats + ((int64_t)n * L) * D;
        float best_v = -INFINITY;
        int best_i = 0;
        for (int c0 = 0; c0 < C; c0 += kPredictGroup) {
            const int cg = min(kPredictGroup, C - c0);
            float acc[kPredictGroup];
#pragma unroll
            for (int j = 0; j < kPredictGroup; ++j) acc[j] = 0.f;
            float qq = 0.f;
            for (int d = lane * 4; d < D; d += 128) {
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int l = 0; l < L; ++l) add4(q, ldg4_stream(row + (int64_t)l * D + d));
                q.x /= fL; q.y /= fL; q.z /= fL; q.w /= fL;
                qq += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
#pragma unroll
                for (int j = 0; j < kPredictGroup; ++j) {
                    if (j < cg) {
                        const float4 w = kSmemW ? *reinterpret_cast<const float4*>(s_w + (size_t)(c0 + j) * D + d)
                                                : ldg4(weight + (int64_t)(c0 + j) * D + d);
                        acc[j] = fmaf(q.x, w.x, acc[j]); acc[j] = fmaf(q.y, w.y, acc[j]);
                        acc[j] = fmaf(q.z, w.z, acc[j]); acc[j] = fmaf(q.w, w.w, acc[j]);
                    }
                }
            }
            qq = warp_sum(qq);
            float mine = -INFINITY;
#pragma unroll
            for (int j = 0; j < kPredictGroup; ++j) {
                const float v = warp_sum(acc[j]);
                if (lane == j) mine = v;
            }
            if (lane < cg) {
                const int c = c0 + lane;
                float v;
                if (cosine) v = logit_scale * (mine / (fmaxf(sqrtf(qq), 1e-8f) * s_wn[c]));
                else        v = logit_scale * (mine + (bias ? __ldg(bias + c) : 0.f));
                logits[(int64_t)n * C + c] = v;
                mine = v;
            } else {
                mine = -INFINITY;
            }
            // warp arg-max, first maximal index wins (torch.argmax convention)
            float v = mine;
            int i = c0 + lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, i, o);
                if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
            }
            if (v > best_v) { best_v = v; best_i = i; }
        }
        if (argmax && lane == 0) argmax[n] = best_i;
    }
}

}  // namespace orbit

using namespace orbit;

extern "C" int orbit_pool_clips(const float* frame_feats, int num_clips, int clip_length, int feat_dim,
                                float* clip_feats, void* stream) {
    if (num_clips < 0 || clip_length <= 0 || feat_dim <= 0) return ORBIT_ERR_ARG;
    if (num_clips == 0) return ORBIT_OK;
    if (!frame_feats || !clip_feats) return ORBIT_ERR_ARG;
    if (feat_dim % 4 || !aligned16(frame_feats) || !aligned16(clip_feats)) return ORBIT_ERR_UNSUPPORTED;
    if (num_clips == 0) return ORBIT_OK;
    const int64_t total = (int64_t)num_clips * (feat_dim / 4);
    const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), 148 * 8);
    pool_clips_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(frame_feats, num_clips, clip_length, feat_dim, clip_feats);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

extern "C" int orbit_pool_history(const float* frame_feats, int num_frames, int history_length, int feat_dim, float* clip_feats,
                                  void* stream) {
    if (num_frames < 0 || history_length <= 0 || feat_dim <= 0) return ORBIT_ERR_ARG;
    if (num_frames == 0) return ORBIT_OK;
    if (!frame_feats || !clip_feats || frame_feats == clip_feats) return ORBIT_ERR_ARG;
    if (feat_dim % 4 || !aligned16(frame_feats) || !aligned16(clip_feats)) return ORBIT_ERR_UNSUPPORTED;
    const int64_t total = (int64_t)num_frames * (feat_dim / 4);
    const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), 148 * 8);
    pool_history_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(frame_feats, num_frames, history_length, feat_dim, clip_feats);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

extern "C" int64_t orbit_proto_configure_scratch_bytes(int num_classes, int feat_dim) {
    if (num_classes <= 0 || feat_dim <= 0) return 0;
    return (int64_t)sizeof(int) * ORBIT_MAX_CLASSES + (int64_t)sizeof(float) * num_classes * ceil_div(feat_dim, 128);
}

extern "C" int orbit_proto_configure(const float* frame_feats, const int32_t* class_index, int num_clips,
                                     int clip_length, int feat_dim, int num_classes, int metric, float* weight,
                                     float* bias, float* proto, void* scratch, void* stream) {
    if (!frame_feats || !class_index || !weight || !scratch) return ORBIT_ERR_ARG;
    if (num_clips <= 0 || clip_length <= 0 || feat_dim <= 0 || num_classes <= 0) return ORBIT_ERR_ARG;
    if (metric != ORBIT_METRIC_EUCLIDEAN && metric != ORBIT_METRIC_COSINE) return ORBIT_ERR_ARG;
    if (metric == ORBIT_METRIC_EUCLIDEAN && !bias) return ORBIT_ERR_ARG;
    if (num_classes > ORBIT_MAX_CLASSES || feat_dim % 4) return ORBIT_ERR_UNSUPPORTED;
    if (!aligned16(frame_feats) || !aligned16(weight) || (proto && !aligned16(proto))) return ORBIT_ERR_UNSUPPORTED;
    dim3 grid(num_classes, ceil_div(feat_dim, 128));
    proto_configure_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        frame_feats, class_index, num_clips, clip_length, feat_dim, metric == ORBIT_METRIC_EUCLIDEAN, weight, bias,
        proto, reinterpret_cast<ConfigureScratch*>(scratch));
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

extern "C" int orbit_head_predict(const float* frame_feats, int num_clips, int clip_length, int feat_dim,
                                  const float* weight, const float* bias, int num_classes, int metric,
                                  float logit_scale, float* logits, int32_t* argmax, void* stream) {
    if (num_clips < 0 || clip_length <= 0 || feat_dim <= 0 || num_classes <= 0) return ORBIT_ERR_ARG;
    if (num_clips == 0 && weight) return ORBIT_OK;       // empty query set: nothing to write (pointers may be null)
    if (!frame_feats || !weight || !logits) return ORBIT_ERR_ARG;
    if (metric != ORBIT_METRIC_EUCLIDEAN && metric != ORBIT_METRIC_COSINE) return ORBIT_ERR_ARG;
    if (num_classes > ORBIT_MAX_CLASSES || feat_dim % 4) return ORBIT_ERR_UNSUPPORTED;
    if (!aligned16(frame_feats) || !aligned16(weight)) return ORBIT_ERR_UNSUPPORTED;
    if (num_clips == 0) return ORBIT_OK;
    const int warps = 8;
    const int blocks = std::min(ceil_div(num_clips, warps), 148 * 4);
    const size_t w_bytes = (size_t)num_classes * feat_dim * sizeof(float);
    const size_t wn_bytes = (size_t)num_classes * sizeof(float);
    const int cosine = metric == ORBIT_METRIC_COSINE;
    cudaStream_t st = (cudaStream_t)stream;
    if (w_bytes + wn_bytes <= 160 * 1024) {
        const size_t smem = w_bytes + wn_bytes;
        if (smem > 48 * 1024)
            ORBIT_CUDA(cudaFuncSetAttribute(head_predict_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        head_predict_kernel<true><<<blocks, warps * 32, smem, st>>>(frame_feats, num_clips, clip_length, feat_dim, weight,
                                                                   bias, num_classes, cosine, logit_scale, logits, argmax);
    } else {
        head_predict_kernel<false><<<blocks, warps * 32, wn_bytes, st>>>(frame_feats, num_clips, clip_length, feat_dim, weight,
                                                                       bias, num_classes, cosine, logit_scale, logits, argmax);
    }
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}
