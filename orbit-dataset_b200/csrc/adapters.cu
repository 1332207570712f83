// FiLM parameter generator and the small dense layers of the Versa head.
//
// Replaces (reference microsoft/ORBIT-Dataset @ 97ccae1):
//   FilmParameterGenerator.forward        model/feature_adapters.py:66-78  (+ DenseBlock model/mlps.py:52-63)
//   DenseResidualBlock.forward            model/mlps.py:41-50  (VersaClassifier hyper-nets, classifier_heads.py:171-180)
// The reference runs ~6 tiny launches per FiLM tensor in a Python loop (34 tensors for EfficientNet-B0); here the
// whole generator is ONE launch that writes gamma'/beta' straight into the film blob the engine folds.
#include "common.cuh"

namespace orbit {

// One FiLM tensor's generator (all offsets into the generator's flat parameter blob).
struct FilmGenEntry {
    int64_t w1, b1, ln_w, ln_b, w2, b2, reg, init;   // Linear(64,64) | LayerNorm(64) | Linear(64,size) | r | gamma0/beta0
    int64_t out;                                     // offset into the film blob
    int32_t size, is_weight;                         // is_weight: gamma' = g0*(g*r+1) else beta' = b0 + g*r
};

// grid (ceil(max_size/256), n_tensors); every block recomputes the 64-wide hidden layer (cheap) and then its outputs.
__global__ void __launch_bounds__(256)
film_generate_kernel(const float* __restrict__ blob, const FilmGenEntry* __restrict__ table, const float* __restrict__ z,
                     int hidden, float* __restrict__ film) {
    extern __shared__ float s_fg[];       // z[hidden], h[hidden]
    float* s_z = s_fg;
    float* s_h = s_fg + hidden;
    __shared__ float s_stat[2];
    const FilmGenEntry e = table[blockIdx.y];
    if (blockIdx.x * blockDim.x >= e.size) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) s_z[i] = z[i];
    __syncthreads();
    for (int o = warp; o < hidden; o += n_warps) {          // Linear(hidden, hidden)
        const float* w = blob + e.w1 + (int64_t)o * hidden;
        float s = 0.f;
        for (int k = lane; k < hidden; k += 32) s = fmaf(__ldg(w + k), s_z[k], s);
        s = warp_sum(s);
        if (lane == 0) s_h[o] = s + blob[e.b1 + o];
    }
    __syncthreads();
    if (warp == 0) {                                        // LayerNorm statistics (biased variance, eps 1e-5)
        float s = 0.f;
        for (int k = lane; k < hidden; k += 32) s += s_h[k];
        const float mean = warp_sum(s) / hidden;
        float v = 0.f;
        for (int k = lane; k < hidden; k += 32) { const float d = s_h[k] - mean; v = fmaf(d, d, v); }
        v = warp_sum(v) / hidden;
        if (lane == 0) { s_stat[0] = mean; s_stat[1] = 1.0f / sqrtf(v + 1e-5f); }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < hidden; k += blockDim.x) { // affine + ReLU
        const float t = (s_h[k] - s_stat[0]) * s_stat[1] * blob[e.ln_w + k] + blob[e.ln_b + k];
        s_z[k] = fmaxf(t, 0.f);
    }
    __syncthreads();
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o < e.size) {                                       // Linear(hidden, size) + regulariser + FiLM rule
        const float* w = blob + e.w2 + (int64_t)o * hidden;
        float g = blob[e.b2 + o];
        for (int k = 0; k < hidden; k += 4) {
            const float4 w4 = ldg4(w + k);
            g = fmaf(w4.x, s_z[k], g); g = fmaf(w4.y, s_z[k + 1], g); g = fmaf(w4.z, s_z[k + 2], g); g = fmaf(w4.w, s_z[k + 3], g);
        }
        const float r = blob[e.reg + o], init = blob[e.init + o];
        film[e.out + o] = e.is_weight ? init * (g * r + 1.0f) : init + g * r;
    }
}

// out[c, o] = act(in[c, :] . W[o, :] + b[o]) (+ skip[c, o]); one warp per output column o, all C rows at once so that
// every weight row is read exactly once. C <= ORBIT_MAX_CLASSES.
template <int CMAX>
__global__ void __launch_bounds__(256)
dense_rows_kernel(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ b,
                  const float* __restrict__ skip, float* __restrict__ out, int C, int I, int O, int act) {
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= O) return;
    float acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
    const float* w = W + (int64_t)o * I;
    for (int k = lane * 4; k < I; k += 128) {
        const float4 w4 = ldg4(w + k);
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (c < C) {
                const float4 x = ldg4(in + (int64_t)c * I + k);
                acc[c] = fmaf(w4.x, x.x, acc[c]); acc[c] = fmaf(w4.y, x.y, acc[c]);
                acc[c] = fmaf(w4.z, x.z, acc[c]); acc[c] = fmaf(w4.w, x.w, acc[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            float v = warp_sum(acc[c]);
            if (lane == 0) {
                v += b ? b[o] : 0.f;
                if (act == 3) v = v > 0.f ? v : expm1f(v);      // ELU(alpha = 1)
                else if (act == 2) v = fmaxf(v, 0.f);
                if (skip) v += skip[(int64_t)c * O + o];
                out[(int64_t)c * O + o] = v;
            }
        }
    }
}

}  // namespace orbit

using namespace orbit;

extern "C" int orbit_film_generate(const float* gen_params, const void* table, int num_tensors, int max_size,
                                   const float* task_embedding, int hidden, float* film, void* stream) {
    if (!gen_params || !table || !task_embedding || !film || num_tensors <= 0 || max_size <= 0) return ORBIT_ERR_ARG;
    if (hidden <= 0 || hidden % 4 || hidden > 1024) return ORBIT_ERR_UNSUPPORTED;
    dim3 grid(ceil_div(max_size, 256), num_tensors);
    film_generate_kernel<<<grid, 256, 2 * hidden * sizeof(float), (cudaStream_t)stream>>>(
        gen_params, reinterpret_cast<const FilmGenEntry*>(table), task_embedding, hidden, film);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

extern "C" int orbit_film_table_entry_bytes(void) { return (int)sizeof(FilmGenEntry); }

extern "C" int orbit_dense_rows(const float* in, const float* weight, const float* bias, const float* skip, float* out,
                                int rows, int in_dim, int out_dim, int act, void* stream) {
    if (!in || !weight || !out || rows <= 0 || in_dim <= 0 || out_dim <= 0) return ORBIT_ERR_ARG;
    if (rows > ORBIT_MAX_CLASSES || in_dim % 4 || !aligned16(in) || !aligned16(weight)) return ORBIT_ERR_UNSUPPORTED;
    const int warps = 8;
    dim3 grid(ceil_div(out_dim, warps));
    cudaStream_t st = (cudaStream_t)stream;
    if (rows <= 16) dense_rows_kernel<16><<<grid, warps * 32, 0, st>>>(in, weight, bias, skip, out, rows, in_dim, out_dim, act);
    else dense_rows_kernel<ORBIT_MAX_CLASSES><<<grid, warps * 32, 0, st>>>(in, weight, bias, skip, out, rows, in_dim, out_dim, act);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}
