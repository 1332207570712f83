// Library-level entry points of the C ABI (include/orbit_b200.h).
#include <cstring>
#include "convnet.cuh"
#include "gemm_tcgen05.cuh"

extern "C" int orbit_abi_version(void) { return ORBIT_ABI_VERSION; }

extern "C" const char* orbit_error_string(int code) {
    switch (code) {
        case ORBIT_OK: return "ok";
        case ORBIT_ERR_ARG: return "orbit: invalid argument (null pointer, bad size or enum)";
        case ORBIT_ERR_UNSUPPORTED: return "orbit: unsupported shape/option or misaligned pointer";
        case ORBIT_ERR_WORKSPACE: return "orbit: workspace too small";
        case ORBIT_ERR_NO_DEVICE: return "orbit: no sm_100 CUDA device";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "orbit: unknown error";
    }
}

extern "C" int orbit_device_check(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return ORBIT_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return ORBIT_ERR_NO_DEVICE; }
    return prop.major == 10 ? ORBIT_OK : ORBIT_ERR_NO_DEVICE;
}

extern "C" int orbit_pointwise_conv(const float* A, const float* W, const float* scale, const float* shift,
                                    const float* gate, const float* residual, float* out, int M, int N, int K,
                                    int rows_per_frame, int act, int mode, float* w_split, void* stream) {
    using namespace orbit;
    if (!A || !W || !scale || !shift || !out || M < 0 || N <= 0 || K <= 0 || rows_per_frame <= 0) return ORBIT_ERR_ARG;
    if (mode < 0 || mode > 2 || act < 0 || act > 4 || act == 3) return ORBIT_ERR_ARG;
    if (!aligned16(A) || !aligned16(W) || !aligned16(out) || !aligned16(scale) || !aligned16(shift)) return ORBIT_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) return launch_pointwise_ffma(A, W, scale, shift, gate, residual, out, M, N, K, rows_per_frame, act, st);
    if (!w_split || !aligned16(w_split)) return ORBIT_ERR_ARG;
    const int rc = launch_tf32_split(W, (int64_t)N * K, w_split, st);
    if (rc) return rc;
    return launch_pointwise_tcgen05(A, w_split, scale, shift, gate, residual, out, M, N, K, rows_per_frame, act,
                                    mode == 1 ? 3 : 1, st);
}

extern "C" int orbit_set_global_option(const char* key, int value) {
    if (!key) return ORBIT_ERR_ARG;
    if (!strcmp(key, "tc_debias_x1000")) { orbit::set_tcgen05_debias((float)value / 1000.0f); return ORBIT_OK; }
    return ORBIT_ERR_UNSUPPORTED;
}
extern "C" int orbit_get_global_option(const char* key, int* value) {
    if (!key || !value) return ORBIT_ERR_ARG;
    if (!strcmp(key, "tc_debias_x1000")) { *value = (int)(orbit::get_tcgen05_debias() * 1000.0f + 0.5f); return ORBIT_OK; }
    return ORBIT_ERR_UNSUPPORTED;
}
