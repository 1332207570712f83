#include "gemm_tcgen05.cuh"

namespace orbit {

__global__ void tf32_split_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = w[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const float lo = v - hi;  // exact
    out[i] = hi;
    out[n + i] = __uint_as_float(__float_as_uint(lo) & 0xffffe000u);
}

int launch_tf32_split(const float* w, int64_t n, float* out, cudaStream_t st) {
    tf32_split_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(w, n, out);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

int launch_pointwise_tcgen05(const float*, const float*, const float*, const float*, const float*, const float*, float*,
                             int, int, int, int, int, int, cudaStream_t) {
    return ORBIT_ERR_UNSUPPORTED;
}

}  // namespace orbit
