// Dev microbenchmark (round 2): issue rate of tcgen05.mma.kind::f16 (fp16 operands, fp32 accumulate, M=128, K=16) next to
// kind::tf32 (K=8) as a function of N; A in shared memory (SWIZZLE_128B K-major) or in tensor memory. One CTA per SM,
// one issuing thread, garbage operands.  Decides whether the 3-product split should run on fp16 pairs instead of tf32 pairs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate_f16 scripts/mma_rate_f16.cu && gpurun_out/mma_rate_f16
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t a) {
    uint64_t d = (uint64_t)((a & 0x3FFFFu) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
// c_format F32 [4,6)=1; a/b format [7,10),[10,13): tf32 = 2, f16 = 0 (kind::f16: 0 = F16, 1 = BF16)
__device__ __forceinline__ uint32_t make_idesc(int fmt, int m, int n) { return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
template <int KIND>   // 0 = tf32, 1 = f16
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int a_tmem, long long* out) {
    extern __shared__ __align__(1024) uint8_t raw[];
    const uint32_t smem = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t slot; __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory"); }
    for (int i = threadIdx.x; i < 60 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw)[i + 256] = 0x3c003c00u;   // fp16 1.0 pairs (a small tf32 too)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(KIND == 0 ? 2 : 0, 128, N);
        const uint64_t a = make_desc_sw128(smem), b = make_desc_sw128(smem + 16384);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                if (KIND == 0) {
                    if (a_tmem)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                     ::"r"(tmem), "r"(tmem + 256 + k * 8), "l"(b + adv), "r"(idesc), "r"(1u) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(tmem), "l"(a + adv), "l"(b + adv), "r"(idesc), "r"(1u) : "memory");
                } else {
                    if (a_tmem)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                     ::"r"(tmem), "r"(tmem + 256 + k * 8), "l"(b + adv), "r"(idesc), "r"(1u) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(tmem), "l"(a + adv), "l"(b + adv), "r"(idesc), "r"(1u) : "memory");
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        for (long long spin = 0; !done; ++spin) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (spin > 50000000ll) __trap();
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    long long* out; cudaMalloc(&out, 8);
    cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 4096;
    printf("clocks per tcgen05.mma, M=128 (148 CTAs, %d MMAs each); tf32: K=8, f16: K=16\n   N   tf32 A=smem  tf32 A=tmem   f16 A=smem   f16 A=tmem   MAC/clk/SM (f16, smem)\n", iters);
    for (int N : {16, 32, 48, 64, 96, 128, 160, 192, 256}) {
        double c[4];
        for (int v = 0; v < 4; ++v) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (v < 2) rate_kernel<0><<<148, 128, 64 * 1024>>>(N, iters, v & 1, out);
                else rate_kernel<1><<<148, 128, 64 * 1024>>>(N, iters, v & 1, out);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            }
            c[v] = (double)h / iters;
        }
        printf("%4d  %10.1f  %11.1f  %11.1f  %11.1f   %8.0f\n", N, c[0], c[1], c[2], c[3], 128.0 * N * 16 / c[2]);
    }
    return 0;
}
