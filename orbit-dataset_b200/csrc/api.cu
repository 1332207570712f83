// Library-level entry points of the C ABI (include/orbit_b200.h).
#include "common.cuh"

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
