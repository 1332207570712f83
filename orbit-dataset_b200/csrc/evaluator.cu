// Device-side per-video metric pass: replaces the softmax -> .cpu() -> numpy statistics of the
// reference's TestEvaluator.append_video / stat functions (utils/eval_metrics.py:27-68,260-276). Integer outputs only
// (correct frames, frames, first correct frame, most frequent prediction): every statistic the reference reports is a
// ratio of these, formed on the host in float64 exactly as numpy does, so results are bit-identical and independent of
// how episodes are sharded over ranks. One CTA per video; integer shared-memory atomics => deterministic.
#include "common.cuh"

namespace orbit {

constexpr int kMaxEvalClasses = 64;

__global__ void __launch_bounds__(256)
video_stats_kernel(const float* __restrict__ logits, const int32_t* __restrict__ predictions, int C,
                   const int32_t* __restrict__ frame_index, const int32_t* __restrict__ video_offsets,
                   const int32_t* __restrict__ video_labels, int32_t* __restrict__ stats, int32_t* __restrict__ pred_out) {
    __shared__ int hist[kMaxEvalClasses];
    __shared__ int s_correct, s_first;
    const int v = blockIdx.x;
    const int begin = video_offsets[v], end = video_offsets[v + 1], n = end - begin;
    const int label = video_labels[v];
    if (threadIdx.x < kMaxEvalClasses) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_correct = 0; s_first = n; }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int row = frame_index ? frame_index[begin + i] : begin + i;
        int p;
        if (logits) {   // softmax is monotone: arg-max of the probabilities == arg-max of the logits; first maximum
            const float* lr = logits + (int64_t)row * C;   // wins on ties, as np.argmax (eval_metrics.py:34)
            float best = lr[0];
            p = 0;
            for (int c = 1; c < C; ++c) {
                const float x = lr[c];
                if (x > best) { best = x; p = c; }
            }
        } else {
            p = predictions[row];
        }
        if (pred_out) pred_out[begin + i] = p;
        if (p >= 0 && p < kMaxEvalClasses) atomicAdd(&hist[p], 1);
        if (p == label) { atomicAdd(&s_correct, 1); atomicMin(&s_first, i); }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int mode = 0;
        for (int c = 1; c < kMaxEvalClasses; ++c) if (hist[c] > hist[mode]) mode = c;   // np.bincount(...).argmax()
        stats[4 * v + 0] = s_correct;
        stats[4 * v + 1] = n;
        stats[4 * v + 2] = s_first;     // == n when no frame was recognised
        stats[4 * v + 3] = mode;
    }
}

}  // namespace orbit

using namespace orbit;

extern "C" int orbit_video_stats(const float* logits, const int32_t* predictions, int num_classes, const int32_t* frame_index,
                                 const int32_t* video_offsets, const int32_t* video_labels, int num_videos, int32_t* stats,
                                 int32_t* pred_out, void* stream) {
    if (num_videos == 0) return ORBIT_OK;
    if (num_videos < 0 || (!logits && !predictions) || !video_offsets || !video_labels || !stats) return ORBIT_ERR_ARG;
    if (num_classes < 1 || num_classes > kMaxEvalClasses) return ORBIT_ERR_UNSUPPORTED;
    video_stats_kernel<<<num_videos, 256, 0, (cudaStream_t)stream>>>(logits, predictions, num_classes, frame_index, video_offsets,
                                                                     video_labels, stats, pred_out);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}
