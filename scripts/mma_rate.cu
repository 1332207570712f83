// Dev microbenchmark: issue rate of tcgen05.mma.kind::tf32 (M=128, K=8) as a function of N, operands A/B in shared
// memory (SWIZZLE_128B K-major) or A in tensor memory. One CTA per SM, one issuing thread, garbage operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate scripts/mma_rate.cu && gpurun_out/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t a) {
    uint64_t d = (uint64_t)((a & 0x3FFFFu) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int a_tmem, int chains, long long* out) {
    extern __shared__ __align__(1024) uint8_t raw[];
    const uint32_t smem = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t slot; __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory"); }
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(raw)[i + 256] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint64_t a = make_desc_sw128(smem), b = make_desc_sw128(smem + 16384);
        const long long t0 = clock64();
        const uint32_t stride = (uint32_t)((N + 31) / 32 * 32);
        for (int i = 0; i < iters; i += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t d = tmem + ((chains == 2 && (k & 1)) ? stride : 0u);   // `chains` independent accumulators
                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                if (a_tmem)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(d), "r"(tmem + 256 + k * 8), "l"(b + adv), "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d), "l"(a + adv), "l"(b + adv), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    long long* out; cudaMalloc(&out, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 4096;
    printf("clocks per tcgen05.mma.kind::tf32 M=128 K=8 (148 CTAs, %d MMAs each)\n   N  chains  A=smem  A=tmem   MAC/clk/SM(smem)\n", iters);
    for (int chains = 1; chains <= 2; ++chains)
        for (int N : {16, 32, 64, 96, 128, 192, 256}) {
            if (chains == 2 && N > 128) continue;
            double c[2];
            for (int at = 0; at < 2; ++at) {
                long long h = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    rate_kernel<<<148, 128, 64 * 1024>>>(N, iters, at, chains, out);
                    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
                }
                c[at] = (double)h / iters;
            }
            printf("%4d  %6d  %6.1f  %6.1f   %8.0f\n", N, chains, c[0], c[1], 128.0 * N * 8 / c[0]);
        }
    return 0;
}
