// Mahalanobis head (Simple CNAPs): per-class regularised covariances, their inverses, quadratic-form logits.
//
// Replaces (reference microsoft/ORBIT-Dataset @ 97ccae1) MahalanobisClassifier.configure / predict / _estimate_cov,
// model/classifier_heads.py:282-368:
//   task_cov = cov(all support features);  per class c (sorted labels):  mu_c = mean,  lambda = n_c/(n_c+1),
//   Sigma_c = lambda*cov(class c) + (1-lambda)*task_cov + I ,  P_c = Sigma_c^-1  (torch.inverse)
//   logits[n,c] = -s * (mu_c - q_n)^T P_c (mu_c - q_n)
// Covariances and the D x D products run on the tcgen05 3xTF32 GEMM; the inverses are an in-place Gauss-Jordan sweep
// without pivoting (every Sigma is symmetric positive definite with eigenvalues >= 1), batched over all classes.
#include "convnet.cuh"
#include "gemm_tcgen05.cuh"

namespace orbit {

// mean over a gathered set of rows: out[d] = mean_j x[rows[j], d]
__global__ void __launch_bounds__(256)
rows_mean_kernel(const float* __restrict__ x, const int32_t* __restrict__ rows, int n, int D, float* __restrict__ out) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    float s = 0.f;
    for (int j = 0; j < n; ++j) s += x[(int64_t)rows[j] * D + d];
    out[d] = s / (float)n;
}

// t[d, j] = x[rows[j], d] - mean[d]  for j < n, zero for n <= j < n_pad  (the K-major operand of cov = T T^T / (n-1))
__global__ void __launch_bounds__(256)
center_transpose_kernel(const float* __restrict__ x, const int32_t* __restrict__ rows, int n, int n_pad, int D,
                        const float* __restrict__ mean, float* __restrict__ t) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)D * n_pad) return;
    const int d = (int)(i / n_pad), j = (int)(i % n_pad);
    t[i] = j < n ? x[(int64_t)rows[j] * D + d] - mean[d] : 0.f;
}

// the reference's single-example branch (classifier_heads.py:360-363): a SCALAR sum_d (x_d - mean_d(x))^2 / (D-1)
__global__ void __launch_bounds__(256)
single_example_scalar_kernel(const float* __restrict__ x, int D, float* __restrict__ out) {
    __shared__ float s_part[8];
    __shared__ float s_mean;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float s = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) s += x[d];
    s = warp_sum(s);
    if (lane == 0) s_part[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += s_part[i]; s_mean = t / (float)D; }
    __syncthreads();
    float q = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) { const float c = x[d] - s_mean; q = fmaf(c, c, q); }
    q = warp_sum(q);
    __syncthreads();
    if (lane == 0) s_part[warp] = q;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += s_part[i]; out[0] = t / (float)(D - 1); }
}

__global__ void fill_kernel(float* __restrict__ p, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// sigma[i,j] = lambda * (class_cov ? class_cov[i,j] : *class_scalar) + (1-lambda) * task_cov[i,j] + (i == j)
__global__ void __launch_bounds__(256)
cov_combine_kernel(const float* __restrict__ class_cov, const float* __restrict__ class_scalar, const float* __restrict__ task_cov,
                   float lambda, int D, float* __restrict__ sigma) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)D * D) return;
    const float cc = class_cov ? class_cov[i] : class_scalar[0];
    const float tc = task_cov ? (1.0f - lambda) * task_cov[i] : 0.f;
    sigma[i] = lambda * cc + tc + ((i / D) == (i % D) ? 1.0f : 0.f);
}

// ---- batched in-place Gauss-Jordan inverse, step k ----------------------------------------------------------
// (a) scale the pivot row, save the pivot column and clear it
__global__ void __launch_bounds__(256)
gj_pivot_kernel(float* __restrict__ mats, int D, int k, float* __restrict__ col) {
    float* A = mats + (int64_t)blockIdx.x * D * D;
    float* c = col + (int64_t)blockIdx.x * D;
    __shared__ float s_inv;
    if (threadIdx.x == 0) s_inv = 1.0f / A[(int64_t)k * D + k];
    __syncthreads();
    const float inv = s_inv;
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        c[j] = j == k ? 0.f : A[(int64_t)j * D + k];        // multipliers of the other rows
        if (j != k) A[(int64_t)j * D + k] = 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        const float v = j == k ? 1.0f : A[(int64_t)k * D + j];
        A[(int64_t)k * D + j] = v * inv;
    }
}
// (b) rank-1 update of every other row: A[i,:] -= col[i] * A[k,:]
__global__ void __launch_bounds__(256)
gj_update_kernel(float* __restrict__ mats, int D, int k, const float* __restrict__ col) {
    float* A = mats + (int64_t)blockIdx.z * D * D;
    const float* c = col + (int64_t)blockIdx.z * D;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j >= D) return;
    const float4 rk = *reinterpret_cast<const float4*>(A + (int64_t)k * D + j);
    const int i0 = blockIdx.y * 16;
#pragma unroll 4
    for (int ii = 0; ii < 16; ++ii) {
        const int i = i0 + ii;
        if (i >= D || i == k) continue;
        const float f = c[i];
        float4* p = reinterpret_cast<float4*>(A + (int64_t)i * D + j);
        float4 v = *p;
        v.x = fmaf(-f, rk.x, v.x); v.y = fmaf(-f, rk.y, v.y); v.z = fmaf(-f, rk.z, v.z); v.w = fmaf(-f, rk.w, v.w);
        *p = v;
    }
}

// diff[n,:] = mu - q[n,:]
__global__ void __launch_bounds__(256)
diff_rows_kernel(const float* __restrict__ mu, const float* __restrict__ q, int Nq, int D, float* __restrict__ diff) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)Nq * D) return;
    diff[i] = mu[i % D] - q[i];
}
// logits[n, c] = -s * sum_j t[n,j] * diff[n,j]   (one warp per row)
__global__ void __launch_bounds__(256)
rowdot_logits_kernel(const float* __restrict__ t, const float* __restrict__ diff, int Nq, int D, float neg_scale,
                     float* __restrict__ logits, int C, int c) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= Nq) return;
    float s = 0.f;
    for (int j = lane * 4; j < D; j += 128) {
        const float4 a = ldg4(t + (int64_t)n * D + j), b = ldg4(diff + (int64_t)n * D + j);
        s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    }
    s = warp_sum(s);
    if (lane == 0) logits[(int64_t)n * C + c] = neg_scale * s;
}

static int64_t align4(int64_t v) { return (v + 3) / 4 * 4; }

}  // namespace orbit

using namespace orbit;

// workspace floats: scale/shift vectors (2D) + transposed centred set (D * n_pad_max) + its tf32 split (2x) +
// task covariance (D*D) + class covariance (D*D) + pivot columns ((C+1)*D) + scalar
extern "C" int64_t orbit_mahalanobis_configure_workspace_bytes(int num_clips, int feat_dim, int num_classes) {
    if (num_clips <= 0 || feat_dim <= 0 || num_classes <= 0) return 0;
    const int64_t D = feat_dim, np = align4(num_clips);
    return (int64_t)sizeof(float) * (2 * D + 3 * D * np + 2 * D * D + (int64_t)(num_classes + 1) * D + 64);
}

// clip_feats [N, D]; order_host [N]: clip indices grouped by class (class 0's clips first, ...); counts_host [C];
// order_dev: the same index array on the device. Outputs: means [C, D], precisions [C, D, D], task_mean [D],
// task_precision [D, D].
extern "C" int orbit_mahalanobis_configure(const float* clip_feats, const int32_t* order_dev, const int32_t* counts_host,
                                           int num_clips, int feat_dim, int num_classes, float* means, float* precisions,
                                           float* task_mean, float* task_precision, void* workspace, void* stream) {
    if (!clip_feats || !order_dev || !counts_host || !means || !precisions || !task_mean || !task_precision || !workspace)
        return ORBIT_ERR_ARG;
    if (num_clips < 2 || feat_dim <= 1 || num_classes <= 0) return ORBIT_ERR_ARG;
    if (feat_dim % 4 || num_classes > ORBIT_MAX_CLASSES) return ORBIT_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = feat_dim, N = num_clips, C = num_classes;
    const int64_t np_max = align4(N);
    float* ws = reinterpret_cast<float*>(workspace);
    float* scale = ws;                         ws += D;
    float* shift = ws;                         ws += D;
    float* tmat = ws;                          ws += (int64_t)D * np_max;
    float* tsplit = ws;                        ws += 2 * (int64_t)D * np_max;
    float* task_cov = ws;                      ws += (int64_t)D * D;
    float* class_cov = ws;                     ws += (int64_t)D * D;
    float* cols = ws;                          ws += (int64_t)(C + 1) * D;
    float* scalar = ws;
    const int dd_blocks = (int)ceil_div64((int64_t)D * D, 256);
    int rc;
    fill_kernel<<<ceil_div(D, 256), 256, 0, st>>>(shift, D, 0.f);

    auto covariance = [&](const int32_t* rows, int n, float* mean_out, float* cov_out) -> int {
        const int n_pad = (int)align4(n);
        rows_mean_kernel<<<ceil_div(D, 256), 256, 0, st>>>(clip_feats, rows, n, D, mean_out);
        center_transpose_kernel<<<(unsigned)ceil_div64((int64_t)D * n_pad, 256), 256, 0, st>>>(clip_feats, rows, n, n_pad, D, mean_out, tmat);
        fill_kernel<<<ceil_div(D, 256), 256, 0, st>>>(scale, D, 1.0f / (float)(n - 1));
        int r = launch_weight_split(tmat, D, n_pad, tsplit, st);
        if (r) return r;
        return launch_pointwise_tcgen05(tmat, tsplit, scale, shift, nullptr, nullptr, cov_out, D, D, n_pad, D, ACT_NONE, 3, st);
    };

    // identity permutation for the whole support set = order_dev (any order gives the same covariance)
    rc = covariance(order_dev, N, task_mean, task_cov);
    if (rc) return rc;
    // task precision = (task_cov + I)^-1 (classifier_heads.py:297): lambda = 1 with the task covariance as "class" term
    cov_combine_kernel<<<dd_blocks, 256, 0, st>>>(task_cov, nullptr, nullptr, 1.0f, D, task_precision);
    int off = 0;
    for (int c = 0; c < C; ++c) {
        const int n = counts_host[c];
        if (n <= 0) return ORBIT_ERR_ARG;
        const float lambda = (float)n / (float)(n + 1);
        float* sigma = precisions + (int64_t)c * D * D;
        if (n > 1) {
            rc = covariance(order_dev + off, n, means + (int64_t)c * D, class_cov);
            if (rc) return rc;
            cov_combine_kernel<<<dd_blocks, 256, 0, st>>>(class_cov, nullptr, task_cov, lambda, D, sigma);
        } else {
            rows_mean_kernel<<<ceil_div(D, 256), 256, 0, st>>>(clip_feats, order_dev + off, 1, D, means + (int64_t)c * D);
            single_example_scalar_kernel<<<1, 256, 0, st>>>(means + (int64_t)c * D, D, scalar);
            cov_combine_kernel<<<dd_blocks, 256, 0, st>>>(nullptr, scalar, task_cov, lambda, D, sigma);
        }
        off += n;
    }
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    // batched in-place inverses; when the caller placed the task matrix right behind the C class matrices (as
    // orbit_b200 does) all C+1 are swept together: 2*D launches in total
    const bool together = task_precision == precisions + (int64_t)C * D * D;
    const int nb = together ? C + 1 : C;
    dim3 ugrid(ceil_div(D / 4, 256), ceil_div(D, 16), nb);
    for (int k = 0; k < D; ++k) {
        gj_pivot_kernel<<<nb, 256, 0, st>>>(precisions, D, k, cols);
        gj_update_kernel<<<ugrid, 256, 0, st>>>(precisions, D, k, cols);
    }
    if (!together) {
        dim3 tgrid(ceil_div(D / 4, 256), ceil_div(D, 16), 1);
        for (int k = 0; k < D; ++k) {
            gj_pivot_kernel<<<1, 256, 0, st>>>(task_precision, D, k, cols + (int64_t)C * D);
            gj_update_kernel<<<tgrid, 256, 0, st>>>(task_precision, D, k, cols + (int64_t)C * D);
        }
    }
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

extern "C" int64_t orbit_mahalanobis_predict_workspace_bytes(int num_clips, int feat_dim) {
    if (num_clips <= 0 || feat_dim <= 0) return 0;
    const int64_t D = feat_dim;
    return (int64_t)sizeof(float) * (2 * D + 2 * (int64_t)num_clips * D + 2 * D * D + 64);
}

extern "C" int orbit_mahalanobis_predict(const float* clip_feats, int num_clips, int feat_dim, const float* means,
                                         const float* precisions, int num_classes, float logit_scale, float* logits,
                                         void* workspace, void* stream) {
    if (num_clips < 0 || feat_dim <= 0 || num_classes <= 0) return ORBIT_ERR_ARG;
    if (num_clips == 0) return ORBIT_OK;     // empty query set: nothing to score (pointers may be null)
    if (!clip_feats || !means || !precisions || !logits || !workspace) return ORBIT_ERR_ARG;
    if (feat_dim % 4) return ORBIT_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = feat_dim, Nq = num_clips;
    float* ws = reinterpret_cast<float*>(workspace);
    float* scale = ws;   ws += D;
    float* shift = ws;   ws += D;
    float* diff = ws;    ws += (int64_t)Nq * D;
    float* tmp = ws;     ws += (int64_t)Nq * D;
    float* psplit = ws;
    fill_kernel<<<ceil_div(D, 256), 256, 0, st>>>(scale, D, 1.0f);
    fill_kernel<<<ceil_div(D, 256), 256, 0, st>>>(shift, D, 0.f);
    for (int c = 0; c < num_classes; ++c) {
        diff_rows_kernel<<<(unsigned)ceil_div64((int64_t)Nq * D, 256), 256, 0, st>>>(means + (int64_t)c * D, clip_feats, Nq, D, diff);
        int rc = launch_weight_split(precisions + (int64_t)c * D * D, D, D, psplit, st);
        if (rc) return rc;
        // tmp = diff . P_c^T  (= diff . P_c, P_c symmetric)
        rc = launch_pointwise_tcgen05(diff, psplit, scale, shift, nullptr, nullptr, tmp, Nq, D, D, Nq, ACT_NONE, 3, st);
        if (rc) return rc;
        rowdot_logits_kernel<<<ceil_div(Nq, 8), 256, 0, st>>>(tmp, diff, Nq, D, -logit_scale, logits, num_classes, c);
    }
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}
