// Pointwise (1x1) convolution as a tcgen05 GEMM for sm_100a:  out[M,N] = epi( (A[M,K] * gate) . W[N,K]^T )
//
// Reference op sites: timm conv_pw / conv_pwl / conv_head + BatchNormAct2d inside the extractor invoked at
// model/few_shot_recognisers.py:114-117,143-146 (88% of EfficientNet-B0's MACs); the
// FiLM gamma'/beta' (model/film.py, feature_adapters.py:66-78) arrive folded into `scale`/`shift`.
//
// Design (one persistent CTA per SM, warp-specialised, 512 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B, zero OOB fill) of the fp32 A tile
//               [128 rows x 32 k] and the weight tiles [BN x 32 k] (tf32 hi and lo parts) into a
//               multi-stage shared-memory ring, completion on mbarriers.
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) with the
//               accumulator in TMEM (double buffered, 2 x BN columns); tcgen05.commit frees ring slots.
//   warps 12-15 A transform: multiply the landed tile by the squeeze-excite gate (per frame, per input
//               channel) and split it into tf32 hi / lo parts in shared memory (3xTF32: hi*hi + hi*lo + lo*hi
//               gives fp32-grade products; the tensor core accumulates in fp32), then fence.proxy.async.
//   warps 4-11  epilogue: tcgen05.ld the accumulator rows, apply folded BN/FiLM scale-shift, SiLU, residual,
//               and store fp32 rows (16-byte vector stores).
// The kernel is HBM-bound by design (A read once, out written once); the tensor pipe has the headroom
// for the three passes.
#include <cuda.h>

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

namespace tc {

constexpr int BM = 128;          // rows per tile (= UMMA M, one TMEM lane per row)
constexpr int BK = 32;           // fp32 elements per k-block = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 8;        // tf32: 32 bytes per instruction along K
constexpr int A_TILE_BYTES = BM * BK * 4;  // 16 KB
constexpr int NUM_THREADS = 512;
constexpr int EPI_WARP0 = 4, NUM_EPI_WARPS = 8;
constexpr int XF_WARP0 = 12, NUM_XF_WARPS = 4;
constexpr int MAX_STAGES = 8;
constexpr int TMEM_COLS = 256;   // 2 accumulators x BN (<= 128) fp32 columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded spin: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if ((spin & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();   // ~2 s: deadlock => launch failure, not a hang
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 = 1024 B between 8-row groups
//   [46,48) version = 1 (Blackwell) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1, a/b format TF32 [7,10)=[10,13)=2, K-major A and B,
// n_dim [17,23) = N>>3, m_dim [24,29) = M>>4
__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct Params {
    const float* scale;
    const float* shift;
    const float* gate;       // [frames, K] or null
    const float* residual;   // [M, N] or null
    float* out;              // [M, N]
    int M, N, K, rows_per_frame, act, passes;
    int BN, n_tiles, m_tiles, stages;
    int b_tile_bytes;        // BN * 128 rounded up to 1024
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
pw_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                  const __grid_constant__ CUtensorMap map_blo, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const bool split = p.passes == 3;
    const bool transform = split || p.gate != nullptr;
    const int a_bytes = A_TILE_BYTES * (split ? 2 : 1);
    const int stage_bytes = a_bytes + p.b_tile_bytes * (split ? 2 : 1);
    auto stage_a_hi = [&](int s) { return smem + (size_t)s * stage_bytes; };
    auto stage_a_lo = [&](int s) { return smem + (size_t)s * stage_bytes + A_TILE_BYTES; };
    auto stage_b_hi = [&](int s) { return smem + (size_t)s * stage_bytes + a_bytes; };
    auto stage_b_lo = [&](int s) { return smem + (size_t)s * stage_bytes + a_bytes + p.b_tile_bytes; };
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;                      // [stages] TMA landed
    uint64_t* ready = bars + MAX_STAGES;        // [stages] transform done
    uint64_t* empty = bars + 2 * MAX_STAGES;    // [stages] MMAs that read the slot retired
    uint64_t* tmem_full = bars + 3 * MAX_STAGES;       // [2]
    uint64_t* tmem_empty = bars + 3 * MAX_STAGES + 2;  // [2]
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_k = ceil_div(p.K, BK);
    const int num_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], NUM_XF_WARPS); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], NUM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            uint32_t it = 0;
            const uint32_t tx = A_TILE_BYTES + (uint32_t)p.BN * BK * 4 * (split ? 2 : 1);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / p.n_tiles) * BM, n0 = (tile % p.n_tiles) * p.BN;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % p.stages;
                    mbar_wait(&empty[s], ((it / p.stages) & 1) ^ 1);
                    mbar_expect_tx(&full[s], tx);
                    tma_load_2d(stage_a_hi(s), &map_a, &full[s], kb * BK, m0);
                    tma_load_2d(stage_b_hi(s), &map_bhi, &full[s], kb * BK, n0);
                    if (split) tma_load_2d(stage_b_lo(s), &map_blo, &full[s], kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, p.BN);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
                const int acc = tcount & 1;
                mbar_wait(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(acc * p.BN);
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(&full[s], ph);
                    if (transform) mbar_wait(&ready[s], ph);
                    tc_fence_after();
                    const uint64_t a_hi = make_desc_sw128(smem_u32(stage_a_hi(s)));
                    const uint64_t b_hi = make_desc_sw128(smem_u32(stage_b_hi(s)));
                    const uint64_t a_lo = make_desc_sw128(smem_u32(stage_a_lo(s)));
                    const uint64_t b_lo = make_desc_sw128(smem_u32(stage_b_lo(s)));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);   // +32 B per k step inside the swizzle row
                        const uint32_t first = (kb | k) ? 1u : 0u;
                        if (split) {
                            umma_tf32(d_addr, a_lo + adv, b_hi + adv, idesc, first);
                            umma_tf32(d_addr, a_hi + adv, b_lo + adv, idesc, 1u);
                            umma_tf32(d_addr, a_hi + adv, b_hi + adv, idesc, 1u);
                        } else {
                            umma_tf32(d_addr, a_hi + adv, b_hi + adv, idesc, first);
                        }
                    }
                    umma_commit(&empty[s]);                       // slot reusable once these MMAs retire
                    if (kb == num_k - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
                }
            }
        }
    } else if (warp >= XF_WARP0) {
        // ================================ A transform ================================
        if (transform) {
            const int t = threadIdx.x - XF_WARP0 * 32;       // 0..127
            const int pchunk = t & 7;                        // physical 16-byte chunk inside the 128-byte row
            const int rbase = t >> 3;                        // rows rbase + 16*i
            const int jchunk = pchunk ^ (rbase & 7);         // logical chunk (SWIZZLE_128B: chunk ^= row & 7)
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / p.n_tiles) * BM;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % p.stages;
                    mbar_wait(&full[s], (it / p.stages) & 1);
                    float4* hi = reinterpret_cast<float4*>(stage_a_hi(s));
                    float4* lo = reinterpret_cast<float4*>(stage_a_lo(s));
                    const int kcol = kb * BK + jchunk * 4;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = rbase + 16 * i;
                        const int idx = row * 8 + pchunk;
                        float4 v = hi[idx];
                        if (p.gate && kcol < p.K) {
                            const int m = min(m0 + row, p.M - 1);
                            const float4 g = ldg4(p.gate + (int64_t)(m / p.rows_per_frame) * p.K + kcol);
                            v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
                        }
                        if (split) {
                            float4 h;
                            h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                            h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                            h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                            h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                            hi[idx] = h;
                            lo[idx] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                        } else {
                            hi[idx] = v;
                        }
                    }
                    fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ready[s]);
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ================================ epilogue ================================
        const int ew = warp - EPI_WARP0;            // 0..7
        const int lane_grp = warp & 3;              // TMEM lanes 32*lane_grp .. +31 are accessible to this warp
        const int col_half = ew >> 2;               // two warps share a lane group and split the columns
        const int row_in_tile = lane_grp * 32 + lane;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const int acc = tcount & 1;
            const int m0 = (tile / p.n_tiles) * BM, n0 = (tile % p.n_tiles) * p.BN;
            mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
            tc_fence_after();
            const int64_t row = (int64_t)m0 + row_in_tile;
            const int n_groups = p.BN / 16;
            for (int g = col_half; g < n_groups; g += 2) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * p.BN + g * 16), v);
                const int col = n0 + g * 16;
                if (row < p.M) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int c = col + q * 4;
                        if (c < p.N) {
                            const float4 sc = ldg4(p.scale + c), sh = ldg4(p.shift + c);
                            float4 o;
                            o.x = fmaf(v[q * 4 + 0], sc.x, sh.x); o.y = fmaf(v[q * 4 + 1], sc.y, sh.y);
                            o.z = fmaf(v[q * 4 + 2], sc.z, sh.z); o.w = fmaf(v[q * 4 + 3], sc.w, sh.w);
                            if (p.act == 1) { o.x = siluf_(o.x); o.y = siluf_(o.y); o.z = siluf_(o.z); o.w = siluf_(o.w); }
                            else if (p.act == 2) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                            if (p.residual) {
                                const float4 r = ldg4_stream(p.residual + row * p.N + c);
                                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                            }
                            *reinterpret_cast<float4*>(p.out + row * p.N + c) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 row-major [rows, cols] tensor, box = [box_rows, 32 cols] (128 bytes), SWIZZLE_128B, zero OOB fill
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return ORBIT_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ORBIT_OK : ORBIT_ERR_UNSUPPORTED;
}

}  // namespace tc

int launch_pointwise_tcgen05(const float* A, const float* w_split, const float* scale, const float* shift,
                             const float* gate, const float* residual, float* out, int M, int N, int K,
                             int rows_per_frame, int act, int passes, cudaStream_t st) {
    using namespace tc;
    if (K % 4 || N % 4 || (passes != 1 && passes != 3)) return ORBIT_ERR_UNSUPPORTED;
    if (M <= 0) return ORBIT_OK;
    Params p;
    p.scale = scale; p.shift = shift; p.gate = gate; p.residual = residual; p.out = out;
    p.M = M; p.N = N; p.K = K; p.rows_per_frame = rows_per_frame; p.act = act; p.passes = passes;
    p.n_tiles = ceil_div(N, 128);
    p.BN = ceil_div(ceil_div(N, p.n_tiles), 16) * 16;
    p.m_tiles = ceil_div(M, BM);
    p.b_tile_bytes = ceil_div(p.BN * BK * 4, 1024) * 1024;
    const int stage_bytes = (A_TILE_BYTES + p.b_tile_bytes) * (passes == 3 ? 2 : 1);
    const int bar_bytes = (3 * MAX_STAGES + 4) * 8 + 16;
    const int budget = 227 * 1024 - 1024 /*alignment slack*/ - bar_bytes;
    p.stages = std::min(MAX_STAGES, budget / stage_bytes);
    if (p.stages < 2) return ORBIT_ERR_UNSUPPORTED;
    const size_t smem = (size_t)p.stages * stage_bytes + bar_bytes + 1024;

    CUtensorMap map_a, map_bhi, map_blo;
    int rc = make_map(&map_a, A, M, K, BM);
    if (rc) return rc;
    rc = make_map(&map_bhi, w_split, N, K, p.BN);
    if (rc) return rc;
    rc = make_map(&map_blo, w_split + (int64_t)N * K, N, K, p.BN);
    if (rc) return rc;

    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        ORBIT_CUDA(cudaGetDevice(&dev));
        ORBIT_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        ORBIT_CUDA(cudaFuncSetAttribute(pw_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const int grid = std::min(p.m_tiles * p.n_tiles, num_sms);
    pw_tcgen05_kernel<<<grid, NUM_THREADS, smem, st>>>(map_a, map_bhi, map_blo, p);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

}  // namespace orbit
