// Library-level entry points of the C ABI (include/orbit_b200.h).
#include <algorithm>
#include <cstring>
#include "convnet.cuh"
#include "gemm_tcgen05.cuh"

namespace orbit {
void set_finetune_grid(int on);
int get_finetune_grid();
}

extern "C" int orbit_abi_version(void) { return ORBIT_ABI_VERSION; }
extern "C" int orbit_experiment_build(void) {
#if defined(ORBIT_EXP_SKIP_A) || defined(ORBIT_EXP_SKIP_B)
    return 1;
#else
    return 0;
#endif
}

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
    const int rc = launch_weight_split(W, N, K, w_split, st);
    if (rc) return rc;
    return launch_pointwise_tcgen05(A, w_split, scale, shift, gate, residual, out, M, N, K, rows_per_frame, act,
                                    mode == 1 ? 3 : 1, st);
}

extern "C" int orbit_conv3x3(const float* x, const float* weight, const float* scale, const float* shift, const float* residual,
                             float* out, int B, int H, int W, int Cin, int Cout, int act, int implicit, float* scratch,
                             int64_t scratch_floats, void* stream) {
    using namespace orbit;
    if (!x || !weight || !scale || !shift || !out || !scratch || B < 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return ORBIT_ERR_ARG;
    if (Cin % 4 || Cout % 4 || !aligned16(x) || !aligned16(out) || !aligned16(scratch)) return ORBIT_ERR_UNSUPPORTED;
    if (scratch_floats < orbit_conv3x3_scratch_floats(B, H, W, Cin, Cout, implicit)) return ORBIT_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int K = 9 * Cin;
    float* w_gemm = scratch;                       // [Cout][K] in (tap, channel) order
    float* w_split = w_gemm + (int64_t)Cout * K;   // fp16 hi | lo
    int rc = launch_conv_weight_relayout(weight, w_gemm, Cout, Cin, 9, K, 0, st);
    if (rc) return rc;
    rc = launch_weight_split(w_gemm, Cout, K, w_split, st);
    if (rc || B == 0) return rc;
    if (implicit) return launch_conv3x3_tcgen05(x, w_split, scale, shift, residual, out, B, H, W, Cin, Cout, act, st);
    float* col = w_split + 2 * (int64_t)Cout * K;
    rc = launch_im2col(x, col, B, H, W, Cin, 3, 1, 1, 1, H, W, K, 0, st);
    if (rc) return rc;
    return launch_pointwise_tcgen05(col, w_split, scale, shift, nullptr, residual, out, B * H * W, Cout, K, H * W, act, 3, st);
}

extern "C" int orbit_se_gate(const float* partial, int groups, int hw, const float* w1, const float* b1, const float* w2t,
                             const float* b2, float* gate, int B, int C, int R, void* stream) {
    using namespace orbit;
    if (!partial || !w1 || !b1 || !w2t || !b2 || !gate || B < 0 || C <= 0 || R <= 0 || groups <= 0 || hw <= 0) return ORBIT_ERR_ARG;
    return launch_se_gate(partial, groups, hw, w1, b1, w2t, b2, gate, B, C, R, (cudaStream_t)stream);
}

extern "C" int orbit_stem_conv(const float* x, const float* weight, const float* scale, const float* shift, float* y, int B, int H,
                               int W, int act, void* stream) {
    using namespace orbit;
    if (!x || !weight || !scale || !shift || !y || B < 0 || H <= 0 || W <= 0) return ORBIT_ERR_ARG;
    const int ho = (H + 1) / 2, wo = (W + 1) / 2;
    const int pad_t = std::max((ho - 1) * 2 + 3 - H, 0) / 2, pad_l = std::max((wo - 1) * 2 + 3 - W, 0) / 2;   // TF "SAME": the odd pixel goes below / right
    return launch_stem(x, weight, scale, shift, y, B, H, W, ho, wo, pad_t, pad_l, 32, act, (cudaStream_t)stream);
}

extern "C" int orbit_conv_first(const float* x, const float* weight, const float* scale, const float* shift, float* y, int B, int H,
                                int W, int k, int stride, int pad, int act, void* stream) {
    using namespace orbit;
    if (!x || !weight || !scale || !shift || !y || B < 0 || H <= 0 || W <= 0) return ORBIT_ERR_ARG;
    const int ho = (H + 2 * pad - k) / stride + 1, wo = (W + 2 * pad - k) / stride + 1;
    if (ho < 1 || wo < 1) return ORBIT_ERR_ARG;
    return launch_conv_first(x, weight, scale, shift, y, B, H, W, 3, 64, k, stride, pad, ho, wo, act, (cudaStream_t)stream);
}

extern "C" int64_t orbit_conv3x3_scratch_floats(int B, int H, int W, int Cin, int Cout, int implicit) {
    const int64_t K = 9 * (int64_t)Cin;
    return 3 * (int64_t)Cout * K + 16 + (implicit ? 0 : (int64_t)B * H * W * K);
}

extern "C" int orbit_debug_set_gemm_trace(void* dev_buffer) { orbit::set_tcgen05_trace(static_cast<unsigned*>(dev_buffer)); return ORBIT_OK; }

extern "C" int orbit_set_global_option(const char* key, int value) {
    if (!key) return ORBIT_ERR_ARG;
    if (!strcmp(key, "tc_debias_x1000")) { orbit::set_tcgen05_debias((float)value / 1000.0f); return ORBIT_OK; }
    if (!strcmp(key, "tc_narrow")) { orbit::set_tcgen05_narrow(value != 0); return ORBIT_OK; }
    if (!strcmp(key, "tc_stream")) { orbit::set_stream_gemm(value != 0); return ORBIT_OK; }
    if (!strcmp(key, "dw5_staged")) { orbit::set_dw5_staged(value); return ORBIT_OK; }
    if (!strcmp(key, "mbconv_stream")) { orbit::set_mbconv_stream(value != 0); return ORBIT_OK; }
    if (!strcmp(key, "se_ring")) { orbit::set_se_ring(value != 0); return ORBIT_OK; }
    if (!strcmp(key, "stem_groups")) { orbit::set_stem_groups(value); return ORBIT_OK; }
    if (!strcmp(key, "tc_wide_xf")) { orbit::set_tcgen05_wide_xf(value); return ORBIT_OK; }
    if (!strcmp(key, "finetune_grid")) { orbit::set_finetune_grid(value != 0); return ORBIT_OK; }
    if (!strcmp(key, "conv_first")) { orbit::set_conv_first(value != 0); return ORBIT_OK; }
    if (!strcmp(key, "tc_fixed_slabs")) { orbit::set_tcgen05_tuning(value != 0, -1); return ORBIT_OK; }
    if (!strcmp(key, "tc_double_min_stages")) { orbit::set_tcgen05_tuning(-1, value); return ORBIT_OK; }
    return ORBIT_ERR_UNSUPPORTED;
}
extern "C" int orbit_get_global_option(const char* key, int* value) {
    if (!key || !value) return ORBIT_ERR_ARG;
    if (!strcmp(key, "tc_debias_x1000")) { *value = (int)(orbit::get_tcgen05_debias() * 1000.0f + 0.5f); return ORBIT_OK; }
    if (!strcmp(key, "tc_narrow")) { *value = orbit::get_tcgen05_narrow(); return ORBIT_OK; }
    if (!strcmp(key, "tc_stream")) { *value = orbit::get_stream_gemm(); return ORBIT_OK; }
    if (!strcmp(key, "dw5_staged")) { *value = orbit::get_dw5_staged(); return ORBIT_OK; }
    if (!strcmp(key, "mbconv_stream")) { *value = orbit::get_mbconv_stream(); return ORBIT_OK; }
    if (!strcmp(key, "se_ring")) { *value = orbit::get_se_ring(); return ORBIT_OK; }
    if (!strcmp(key, "stem_groups")) { *value = orbit::get_stem_groups(); return ORBIT_OK; }
    if (!strcmp(key, "tc_wide_xf")) { *value = orbit::get_tcgen05_wide_xf(); return ORBIT_OK; }
    if (!strcmp(key, "finetune_grid")) { *value = orbit::get_finetune_grid(); return ORBIT_OK; }
    if (!strcmp(key, "conv_first")) { *value = orbit::get_conv_first(); return ORBIT_OK; }
    return ORBIT_ERR_UNSUPPORTED;
}

static void same_geom(int in, int k, int s, int* out, int* pad_before) {
    *out = (in + s - 1) / s;
    const int total = (*out - 1) * s + k - in;
    *pad_before = (total > 0 ? total : 0) / 2;
}

extern "C" int64_t orbit_depthwise_partial_floats(int B, int H, int W, int C, int k, int stride) {
    int ho, wo, p;
    same_geom(H, k, stride, &ho, &p); same_geom(W, k, stride, &wo, &p);
    return (int64_t)B * orbit::dw_partial_groups(C, ho, wo, k, stride) * C;
}

extern "C" int orbit_mbconv_partial_groups(int H, int W, int Cin, int C, int k, int stride) {
    int ho, wo, p;
    same_geom(H, k, stride, &ho, &p); same_geom(W, k, stride, &wo, &p);
    return orbit::mbx_partial_groups(Cin, C, ho, wo, k, stride);
}

extern "C" int64_t orbit_mbconv_partial_floats(int B, int H, int W, int C, int k, int stride) {
    int ho, wo, p;
    same_geom(H, k, stride, &ho, &p); same_geom(W, k, stride, &wo, &p);
    return (int64_t)B * std::max(orbit::mbx_partial_groups(16, C, ho, wo, k, stride), orbit::mbx_partial_groups(24, C, ho, wo, k, stride)) * C;
}

extern "C" int orbit_mbconv_expand_dw(const float* x, const float* w_expand, const float* scale1, const float* shift1,
                                      const float* w_dw, const float* scale2, const float* shift2, float* y, float* partial,
                                      float* weight_scratch, int B, int H, int W, int Cin, int C, int k, int stride, void* stream) {
    using namespace orbit;
    if (!x || !w_expand || !scale1 || !shift1 || !w_dw || !scale2 || !shift2 || !y || !weight_scratch) return ORBIT_ERR_ARG;
    if (B < 0 || H <= 0 || W <= 0 || Cin <= 0 || C <= 0) return ORBIT_ERR_ARG;
    if (C % 4 || !mbx_supported(Cin, k, stride)) return ORBIT_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(w_expand) || !aligned16(y) || !aligned16(weight_scratch)) return ORBIT_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int ho, wo, pt, pl;
    same_geom(H, k, stride, &ho, &pt); same_geom(W, k, stride, &wo, &pl);
    int rc = launch_dw_relayout(w_dw, C, k * k, weight_scratch, st);
    if (rc) return rc;
    if (B == 0) return ORBIT_OK;
    return launch_mbconv_expand_dw(x, w_expand, scale1, shift1, weight_scratch, scale2, shift2, y, partial, B, H, W, Cin, C, ho, wo,
                                   k, stride, pt, pl, st);
}

extern "C" int orbit_depthwise_conv(const float* x, const float* weight, const float* scale, const float* shift, float* y,
                                    float* partial, float* weight_scratch, int B, int H, int W, int C, int k, int stride,
                                    int act, void* stream) {
    using namespace orbit;
    if (!x || !weight || !scale || !shift || !y || !weight_scratch || B < 0 || H <= 0 || W <= 0 || C <= 0) return ORBIT_ERR_ARG;
    if ((k != 3 && k != 5) || (stride != 1 && stride != 2) || C % 4) return ORBIT_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(y) || !aligned16(scale) || !aligned16(shift) || !aligned16(weight_scratch)) return ORBIT_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int ho, wo, pt, pl;
    same_geom(H, k, stride, &ho, &pt); same_geom(W, k, stride, &wo, &pl);
    int rc = launch_dw_relayout(weight, C, k * k, weight_scratch, st);
    if (rc) return rc;
    if (B == 0) return ORBIT_OK;
    return launch_depthwise(x, weight_scratch, scale, shift, y, partial, B, H, W, C, ho, wo, k, stride, pt, pl, act, st);
}
