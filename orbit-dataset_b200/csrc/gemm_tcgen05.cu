// Pointwise (1x1) convolution as a tcgen05 GEMM for sm_100a:  out[M,N] = epi( (A[M,K] * gate) . W[N,K]^T )
//
// Reference op sites: timm conv_pw / conv_pwl / conv_head + BatchNormAct2d inside the extractor invoked at
// model/few_shot_recognisers.py:114-117,143-146 (88% of EfficientNet-B0's MACs); the
// FiLM gamma'/beta' (model/film.py, feature_adapters.py:66-78) arrive folded into `scale`/`shift`.
//
// Numerics: fp32-grade products from THREE fp16 tensor-core products ("FP16x3").  Every fp32 operand is split as
//   x = hi + lo * 2^-11,  hi = fp16(x),  lo = fp16((x - hi) * 2^11)        (22-23 significant bits, like 3xTF32)
// and  a.w ~= a_hi.w_hi + 2^-11 (a_hi.w_lo + a_lo.w_hi).  fp16 has the same 11-bit significand as tf32, and
// tcgen05.mma.kind::f16 covers K = 16 per instruction where kind::tf32 covers 8 AT THE SAME COST PER INSTRUCTION
// (scripts/mma_rate_f16.cu, profiles/r02_mma_rate_f16.txt: 93 clk for every N <= 96 for both kinds, the cost follows
// the operand BYTES): half the tensor-core instructions and half the shared-memory operand traffic of round 1's
// 3xTF32 kernel. The price is fp16's exponent range: |x| must stay below 65504 (conversions saturate); values
// below 2^-14 lose nothing that matters because the scaled lo part carries the residual.
//
// Design (one persistent CTA per SM, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B, zero OOB fill) of the fp32 A tile
//               [128 rows x 64 k] (two 128-byte-wide boxes) and the fp16 weight tiles [BN x 64 k] (hi and lo
//               parts, split once per task) into a multi-stage shared-memory ring, completion on mbarriers.
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.kind::f16 (M=128, N=BN, K=16), accumulators in TMEM.
//   warps 4..   A transform: multiply the landed fp32 tile by the squeeze-excite gate (per frame, per input
//               channel), split it into fp16 hi / lo IN PLACE (32 KB of fp32 become 16 KB hi + 16 KB lo in the
//               K-major SWIZZLE_128B layout the MMA reads), then fence.proxy.async.
//   last 12     epilogue: per k-block promotion of the main accumulator into fp32 registers (see below), then
//               folded BN/FiLM scale-shift, SiLU/ReLU/GELU, residual, swizzled staging slab, TMA store.
// The kernel is HBM-bound by design (A read once, out written once).
#include <cuda.h>
#include <cuda_fp16.h>

#include "gemm_tcgen05.cuh"

namespace orbit {

// w [N,K] fp32 -> out: fp16 hi [N,Kp] | fp16 lo [N,Kp], Kp = K rounded up to 8 (16-byte rows for the TMA), zero padded
__global__ void weight_split_kernel(const float* __restrict__ w, int N, int K, int Kp, __half* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * Kp) return;
    const int n = (int)(i / Kp), k = (int)(i % Kp);
    const float v = k < K ? w[(int64_t)n * K + k] : 0.f;
    const __half hi = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    out[i] = hi;
    out[(int64_t)N * Kp + i] = __float2half_rn(fminf(fmaxf((v - __half2float(hi)) * 2048.0f, -65504.f), 65504.f));   // v - hi is exact in fp32
}

int launch_weight_split(const float* w, int N, int K, float* out, cudaStream_t st) {
    const int Kp = (K + 7) / 8 * 8;
    weight_split_kernel<<<(unsigned)ceil_div64((int64_t)N * Kp, 256), 256, 0, st>>>(w, N, K, Kp, reinterpret_cast<__half*>(out));
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

namespace tc {

constexpr int BM = 128;          // rows per tile (= UMMA M, one TMEM lane per row)
constexpr int BK = 64;           // k per ring stage = one 128-byte swizzle row of fp16 = two 128-byte rows of fp32
constexpr int UMMA_K = 16;       // fp16: 32 bytes per instruction along K
constexpr int A_BOX_BYTES = BM * 32 * 4;      // one fp32 TMA box [128 rows x 32 k] = 16 KB = one fp16 operand tile [128 x 64]
constexpr int A_STAGE_BYTES = 2 * A_BOX_BYTES;
// Warp roles. The hardware arbiter favours higher warp ids, so the epilogue (the role with real ALU work)
// sits last; its first warp id must be a multiple of 4 (a warp reaches TMEM lanes 32*(warp%4)..+31).
constexpr int XF_WARP0 = 4;       // transform warps: XFW = 4 (ungated) or 8 (gated), template parameter
constexpr int NUM_EPI_WARPS = 12, EPI_SPLIT = NUM_EPI_WARPS / 4;    // 4 lane groups x 3 column shares
constexpr int num_threads(int xfw) { return (XF_WARP0 + xfw + NUM_EPI_WARPS) * 32; }   // 640 (<= 102 registers) or 768 (<= 85)
constexpr int MAX_STAGES = 8;
constexpr int NMAIN = 3;          // main (per-k-block) accumulators in flight: the MMA issuer may run this far ahead of the epilogue
constexpr int TMEM_COLS = 512;    // main accumulator x NMAIN + correction accumulator x2, BN (<= 96) fp32 columns each: 480
constexpr int SLAB_BYTES = 32 * 128;       // epilogue staging slab: 32 rows x 32 fp32, SWIZZLE_128B (1 or 2 per warp)
constexpr int L2_PREFETCH_DISTANCE = 6;    // k-blocks (32 KB of A each) requested into L2 ahead of the smem ring

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// All shared-memory traffic of this kernel uses explicit shared-state-space instructions on 32-bit addresses:
// through generic pointers the compiler emitted LD.E/ST.E with 64-bit address arithmetic (ncu, round 1).
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Blocking wait with a hardware suspend-time hint: the warp sleeps inside try_wait until the phase completes
// (or ~20 us pass) instead of burning issue slots in a polling loop. Bounded: a protocol bug must surface
// as a trap (launch failure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (spin > 400000u) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128_u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32u(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// Packed fp32x2 arithmetic (Blackwell): two IEEE round-to-nearest fp32 operations per issue slot. The epilogue is
// issue-bound on the HBM-shaped layers, and every value it touches comes in adjacent-column pairs.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) { f2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ void lds128_f2(uint32_t addr, f2_t& a, f2_t& b) {
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts128_f2(uint32_t addr, f2_t a, f2_t b) {
    asm volatile("st.shared.v2.b64 [%0], {%1,%2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}
// x * sigmoid(x) on a pair: the two MUFU ops per element stay scalar, the three fp32 ops around them are packed
__device__ __forceinline__ f2_t f2_silu(f2_t x) {
    float t0, t1, e0, e1, r0, r1;
    f2_unpack(f2_mul(x, f2_pack(-1.4426950408889634f, -1.4426950408889634f)), t0, t1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
    f2_unpack(f2_add(f2_pack(e0, e1), f2_pack(1.0f, 1.0f)), t0, t1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(t0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(t1));
    return f2_mul(x, f2_pack(r0, r1));
}
__device__ __forceinline__ f2_t f2_relu(f2_t x) { float a, b; f2_unpack(x, a, b); return f2_pack(fmaxf(a, 0.f), fmaxf(b, 0.f)); }

// (x0, x1) fp32 -> packed fp16 pair hi (x0 in the low half = lower address = smaller k) and the scaled residual pair
// lo = fp16((x - hi) * 2^11). x - hi is exact in fp32, and so are the 2^11 scalings, so the residual is formed with two
// packed fp32x2 operations (x * 2^11, then hi * -2^11 + that); saturating conversions keep out-of-range inputs finite.
__device__ __forceinline__ void split_f16x2(f2_t x, uint32_t& hi, uint32_t& lo) {
    float x0, x1, h0, h1, r0, r1;
    f2_unpack(x, x0, x1);
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
    const f2_t r = f2_fma(f2_pack(h0, h1), f2_pack(-2048.0f, -2048.0f), f2_mul(x, f2_pack(2048.0f, 2048.0f)));
    f2_unpack(r, r0, r1);
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ uint32_t pack_f16x2(float x0, float x1) {
    uint32_t hi;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    return hi;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
// cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1, a/b format F16 [7,10)=[10,13)=0, K-major A and B,
// n_dim [17,23) = N>>3, m_dim [24,29) = M>>4
__device__ __forceinline__ uint32_t make_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float* v) {   // no wait: pair with tmem_ld_wait()
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

struct Params {
    const float* scale;
    const float* shift;
    const float* gate;       // [frames, K] or null
    int has_residual;
    int M, N, K, rows_per_frame, act;
    uint32_t rpf_mul, rpf_shift;   // row / rows_per_frame = __umulhi(row, rpf_mul) >> rpf_shift for row < 2^31 (host: fast_div)
    // implicit 3x3 stride-1 pad-1 convolution (CONV3 instances): A = the NHWC activation [B*H*W, Cin] itself; k-block kb covers
    // tap kb / conv_kpt (ky = tap / 3, kx = tap % 3), channels (kb % conv_kpt) * 64 ..; its A tile = the rows m + (ky-1) W + (kx-1)
    int conv_w, conv_h, conv_kpt;
    uint32_t cw_mul, cw_shift, ch_mul, ch_shift;   // fast division by W and by H
    int BN, n_tiles, m_tiles, stages;
    int b_tile_bytes;        // BN * 128 (multiple of 2048)
    int slabs_per_warp;      // 1 or 2 staging slabs per epilogue warp
    int staging_warps;       // epilogue warps that own staging slabs: 12, or 4 * slabs-per-tile with the fixed mapping
    int fixed_slabs;         // tiles of < 3 slabs: slab sl always belongs to column share sl (no rotation): the other shares' staging
                             // memory goes to the operand ring (one or two more stages in flight for the small-N projections)
    float debias;            // kappa: expected truncation loss of a promoted k-block partial, in ulps of that partial
    unsigned* trace;         // dev aid (orbit_debug_set_gemm_trace): per-role clock stamps of CTA 0, [kTraceSteps][16]; else null
};

// Tile -> (m tile, n tile) without a division per tile: the SASS of the straightforward `tile / n_tiles`, `tile % n_tiles`,
// `row / rows_per_frame` had an I2F + MUFU.RCP + fix-up sequence (~150 dependent clocks) for EACH of them at the top of every
// tile in every role (in-kernel trace, round 2: 800 clocks between a slab's TMA store and the next tile's first wait).
struct TileIter {
    int mt, nt, step_m, step_n, n_tiles;
    __device__ __forceinline__ TileIter(int tile0, int stride, int n_tiles_) : n_tiles(n_tiles_) {
        mt = tile0 / n_tiles; nt = tile0 - mt * n_tiles;
        step_m = stride / n_tiles; step_n = stride - step_m * n_tiles;
    }
    __device__ __forceinline__ void next() {
        mt += step_m; nt += step_n;
        if (nt >= n_tiles) { nt -= n_tiles; ++mt; }
    }
};

constexpr int kTraceSteps = 256;   // k-block steps of CTA 0 recorded when Params::trace is set
#if defined(ORBIT_GEMM_TRACE)   // build with ORBIT_NVCC_EXTRA=-DORBIT_GEMM_TRACE (costs registers: not in the shipped library)
__device__ __forceinline__ void trace_stamp(unsigned* trace, uint32_t step, int slot) {
    if (trace && blockIdx.x == 0 && step < (uint32_t)kTraceSteps) trace[step * 16 + slot] = (unsigned)clock64();
}
#else
__device__ __forceinline__ void trace_stamp(unsigned*, uint32_t, int) {}
#endif
constexpr int SS_BYTES = 256;   // per epilogue warp: scale[32] | shift[32] of the slab it is finishing

// Template parameters fix at compile time what would otherwise be decided per element at run time:
//   SPLIT  FP16x3 (hi/lo, fp32-grade) or one plain fp16 product (the `fast` numerics mode, 2^-11 relative per product);
//   GATED / RES  0, 1, or -1 = look at the arguments;  ACT  activation or -1;  XFW  transform warps (4 or 8).
//   NARROW the transform's lane mapping for K <= 32 (one k-block, fp32 box 0 only), see the transform role.
//   CONV3  implicit 3x3 convolution: the producer walks the nine taps with row-shifted TMA boxes, the transform zeroes the rows
//          whose tap falls outside the image (no im2col matrix in HBM).
template <bool SPLIT, int GATED, int ACT, int RES, int XFW, bool NARROW, bool CONV3 = false>
__global__ void __launch_bounds__(num_threads(XFW), 1)
pw_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                  const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_out,
                  const __grid_constant__ CUtensorMap map_res, const Params p) {
    constexpr int NUM_XF_WARPS = XFW, EPI_WARP0 = XF_WARP0 + XFW;
    constexpr int NM = NMAIN;
    const uint32_t MS = (uint32_t)p.BN;                            // TMEM columns per main buffer
    const uint32_t CORR0 = NM * MS;                                // correction accumulators (x2)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const bool gated = GATED < 0 ? p.gate != nullptr : GATED != 0;
    const bool has_res = RES < 0 ? p.has_residual != 0 : RES != 0;
    const int act = ACT < 0 ? p.act : ACT;
    const uint32_t stage_bytes = A_STAGE_BYTES + (uint32_t)p.b_tile_bytes * (SPLIT ? 2 : 1);
    const uint32_t staging = smem;                                   // [NUM_EPI_WARPS][slabs_per_warp][32 rows][128 B]
    const uint32_t sstab = staging + (uint32_t)(p.staging_warps * p.slabs_per_warp * SLAB_BYTES);   // [NUM_EPI_WARPS][SS_BYTES]
    const uint32_t ring = sstab + NUM_EPI_WARPS * SS_BYTES;          // (3 KB: the ring stays 1024-byte aligned)
    const uint32_t bars = ring + (uint32_t)p.stages * stage_bytes;
    auto full = [&](uint32_t s) { return bars + 8u * s; };                              // TMA landed
    auto ready = [&](uint32_t s) { return bars + 8u * (MAX_STAGES + s); };              // transform done
    auto empty = [&](uint32_t s) { return bars + 8u * (2 * MAX_STAGES + s); };          // MMAs that read the slot retired
    auto tmem_empty = [&](uint32_t a) { return bars + 8u * (3 * MAX_STAGES + a); };     // [2] correction accumulator drained
    auto main_full = [&](uint32_t a) { return bars + 8u * (3 * MAX_STAGES + 2 + a); };  // [NMAIN] main accumulator of one k-block complete
    auto main_empty = [&](uint32_t a) { return bars + 8u * (3 * MAX_STAGES + 5 + a); }; // [NMAIN] ... added into the epilogue's registers
    auto res_bar = [&](uint32_t w) { return bars + 8u * (3 * MAX_STAGES + 8 + w); };    // residual slab landed
    const uint32_t tmem_base_slot = bars + 8u * (3 * MAX_STAGES + 8 + NUM_EPI_WARPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_k = ceil_div(p.K, BK);
    const int num_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full(s), 1); mbar_init(ready(s), NUM_XF_WARPS); mbar_init(empty(s), 1); }
        for (int a = 0; a < 2; ++a) mbar_init(tmem_empty(a), NUM_EPI_WARPS);
        for (int a = 0; a < NM; ++a) { mbar_init(main_full(a), 1); mbar_init(main_empty(a), NUM_EPI_WARPS); }
        for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(res_bar(w), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_base_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = lds32u(tmem_base_slot);
    // TMEM columns: main accumulators (hi*hi) at {0,1,2}*BN, correction accumulators (hi*lo + lo*hi, scaled 2^11) at {3,4}*BN.
    // The tensor core adds into its fp32 accumulator with TRUNCATION (measured: -0.45 ulp per accumulation, a
    // systematic bias that grows with K and compounds over the network's ~33 GEMM layers). So the main accumulator
    // only ever holds ONE k-block (4 MMAs): the epilogue warps add it into fp32 registers with round-to-nearest
    // every k-block (triple-buffered against the MMAs), and the 2^-11-scaled correction terms -- whose truncation
    // error is negligible -- accumulate over the whole tile in their own accumulator.

    if (warp < XF_WARP0) {
    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            const uint32_t b_bytes = (uint32_t)p.b_tile_bytes * (SPLIT ? 2 : 1);
            // L2 prefetch cursor: runs L2_PREFETCH_DISTANCE k-blocks ahead of the shared-memory ring, so that HBM
            // latency is covered by requests that cost no shared memory (the ring only has to cover L2 latency).
            int pf_tile = blockIdx.x, pf_kb = 0, pf_tap = 0, pf_kc = 0;
            TileIter pf_ti(blockIdx.x, gridDim.x, p.n_tiles);
            auto prefetch_next = [&]() {
                if (pf_tile >= num_tiles) return;
                int m0 = pf_ti.mt * BM, c0 = pf_kb * BK;
                if (CONV3) {
                    c0 = pf_kc * BK;
                    m0 += (pf_tap / 3 - 1) * p.conv_w + (pf_tap % 3 - 1);
                    if (++pf_kc == p.conv_kpt) { pf_kc = 0; ++pf_tap; }
                }
                tma_prefetch_l2_2d(&map_a, c0, m0);
                if (CONV3 || pf_kb * BK + 32 < p.K) tma_prefetch_l2_2d(&map_a, c0 + 32, m0);
                if (++pf_kb == num_k) { pf_kb = 0; pf_tap = 0; pf_kc = 0; pf_tile += gridDim.x; pf_ti.next(); }
            };
            for (int i = 0; i < L2_PREFETCH_DISTANCE; ++i) prefetch_next();
            uint32_t s = 0, ph = 0, step = 0;
            TileIter ti(blockIdx.x, gridDim.x, p.n_tiles);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ti.next()) {
                const int m0 = ti.mt * BM, n0 = ti.nt * p.BN;
                int tap = 0, kc = 0;
                for (int kb = 0; kb < num_k; ++kb, ++step) {
                    prefetch_next();
                    const bool two = CONV3 || kb * BK + 32 < p.K;  // the second 32-wide fp32 box holds real columns
                    int a_col = kb * BK, a_row = m0;
                    if (CONV3) {                                   // tap-shifted rows of the activation; rows outside [0, M) are zero-filled
                        a_col = kc * BK;
                        a_row = m0 + (tap / 3 - 1) * p.conv_w + (tap % 3 - 1);
                        if (++kc == p.conv_kpt) { kc = 0; ++tap; }
                    }
                    mbar_wait(empty(s), ph ^ 1);
                    trace_stamp(p.trace, step, 0);
#ifdef ORBIT_EXP_SKIP_B      // timing experiment only (wrong results): the weight tiles are fetched during the first trip round the ring only
                    const bool load_b = step < (uint32_t)p.stages;
#else
                    const bool load_b = true;
#endif
#ifdef ORBIT_EXP_SKIP_A
                    const bool load_a = step < (uint32_t)p.stages;
#else
                    const bool load_a = true;
#endif
                    if (!load_a && !load_b) { mbar_arrive(full(s)); }
                    else mbar_expect_tx(full(s), (load_a ? (two ? 2u : 1u) * A_BOX_BYTES : 0u) + (load_b ? b_bytes : 0u));
                    const uint32_t st = ring + s * stage_bytes;
                    if (load_a) {
                        tma_load_2d(st, &map_a, full(s), a_col, a_row);
                        if (two) tma_load_2d(st + A_BOX_BYTES, &map_a, full(s), a_col + 32, a_row);
                    }
                    if (load_b) {
                        tma_load_2d(st + A_STAGE_BYTES, &map_bhi, full(s), kb * BK, n0);
                        if (SPLIT) tma_load_2d(st + A_STAGE_BYTES + p.b_tile_bytes, &map_blo, full(s), kb * BK, n0);
                    }
                    trace_stamp(p.trace, step, 1);
                    if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(BM, p.BN);
            uint32_t it = 0, tcount = 0, s = 0, ph = 0, mb = 0, mph = 0;   // it = global k-block counter; mb/mph = main buffer and its parity
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t acc = tcount & 1;
                if (SPLIT) mbar_wait(tmem_empty(acc), ((tcount >> 1) & 1) ^ 1);
                const uint32_t d_corr = tmem_base + CORR0 + acc * (uint32_t)p.BN;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int ksteps = min(BK / UMMA_K, ceil_div(p.K - kb * BK, UMMA_K));
                    mbar_wait(main_empty(mb), mph ^ 1);
                    trace_stamp(p.trace, it, 4);
                    mbar_wait(ready(s), ph);   // the transform warps saw `full` (A and B landed) before they arrived
                    trace_stamp(p.trace, it, 6);
                    tc_fence_after();
                    const uint32_t d_main = tmem_base + mb * MS;
                    const uint32_t st = ring + s * stage_bytes;
                    const uint64_t a_hi = make_desc_sw128(st);
                    const uint64_t a_lo = make_desc_sw128(st + A_BOX_BYTES);
                    const uint64_t b_hi = make_desc_sw128(st + A_STAGE_BYTES);
                    const uint64_t b_lo = make_desc_sw128(st + A_STAGE_BYTES + p.b_tile_bytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (k < ksteps) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);   // +32 B per k step inside the swizzle row
                            if (SPLIT) {
                                umma_f16(d_corr, a_lo + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                                umma_f16(d_corr, a_hi + adv, b_lo + adv, idesc, 1u);
                            }
                            umma_f16(d_main, a_hi + adv, b_hi + adv, idesc, k ? 1u : 0u);
                        }
                    }
                    umma_commit(empty(s));           // ring slot reusable once these MMAs retire
                    umma_commit(main_full(mb));      // this k-block's main accumulator is complete (and, after the
                                                     // last k-block, the tile's correction accumulator too)
                    trace_stamp(p.trace, it, 7);
                    if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
                    if (++mb == NM) { mb = 0; mph ^= 1; }
                }
            }
        }
    }
    } else if (warp < EPI_WARP0) {
        // ================================ A transform ================================
        // fp32 [128 x 64] (two TMA boxes, SWIZZLE_128B: 16-byte chunk ^= row & 7) -> fp16 hi tile (overlays box 0) and
        // lo tile (overlays box 1), in place. A thread converts 8 consecutive k of one row: two 16-byte source chunks,
        // one 16-byte destination chunk per tile. All lanes that touch a row sit in ONE warp and a batch's loads are
        // issued before its stores, which is what makes the in-place overwrite safe.
        //   K > 32: 8 lanes per row (lane q <-> k = 8q..8q+7: box q/4, source chunks 2(q&3), 2(q&3)+1). The lanes of box 1
        //           fetch their ODD chunk first: a quarter-warp then reads 8 distinct bank groups per LDS.128 (it was a
        //           2-way conflict when both boxes fetched the even chunk: ncu r02a, 48 % of the shared wavefronts).
        //   K <= 32 (one k-block, box 0 only: the 112x112 / 56x56 layers): 4 lanes per row, 8 rows per warp step instead of
        //           leaving half (K = 24, 32) or three quarters (K = 16) of the lanes idle; a quarter-warp holds rows m and
        //           m ^ 5, whose swizzles differ in bits 0 and 2 = conflict-free loads AND stores.
        constexpr int XR = BM / (NUM_XF_WARPS * 4);      // rows per thread in the wide mapping (8 or 4)
        constexpr int XS = NUM_XF_WARPS * 4;             // row stride between them (16 or 32: multiples of the 8-row swizzle period)
        constexpr int RB = 2;                            // rows per load/convert/store batch
        const int t = threadIdx.x - XF_WARP0 * 32;
        constexpr bool narrow = NARROW;
        int q, row0, box;
        if (narrow) {
            const int j = lane >> 2, m = j >> 1;
            q = lane & 3; box = 0;
            row0 = (t >> 5) * 8 + ((j & 1) ? (m ^ 5) : m);
        } else {
            q = t & 7; box = q >> 2;
            row0 = t >> 3;
        }
        constexpr int nrows = narrow ? XR / 2 : XR;      // rows row0 + rstride * i
        constexpr uint32_t rstride = (narrow ? 2u : 1u) * XS * 128u;      // bytes
        const int sw = row0 & 7;
        const bool swap = box != 0;                       // this lane loaded its odd chunk first
        const int ka = 8 * q + 4 * box, kb4 = 8 * q + 4 * (1 - box);      // k (inside the k-block) of the first / second chunk loaded
        const uint32_t src_off = (uint32_t)(box * A_BOX_BYTES + row0 * 128 + (((2 * (q & 3) + box) ^ sw) * 16));   // second chunk: ^ 16
        const uint32_t dst_off = (uint32_t)(row0 * 128 + ((q ^ sw) * 16));
        uint32_t s = 0, ph = 0, xstep = 0;
        TileIter xti(blockIdx.x, gridDim.x, p.n_tiles);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, xti.next()) {
            const float* grow[XR];
            if (gated) {
                const int m0 = xti.mt * BM;
#pragma unroll
                for (int i = 0; i < XR; ++i) {
                    const uint32_t row = (uint32_t)min(m0 + row0 + (int)(rstride >> 7) * i, p.M - 1);
                    grow[i] = p.gate + (int64_t)(p.rpf_mul ? (__umulhi(row, p.rpf_mul) >> p.rpf_shift) : row) * p.K;
                }
            }
            uint32_t pix[XR];                    // CONV3: x | y << 16 of this thread's rows (0xffffffff: row beyond M)
            if (CONV3) {
                const int m0 = xti.mt * BM;
#pragma unroll
                for (int i = 0; i < XR; ++i) {
                    const uint32_t row = (uint32_t)(m0 + row0 + (int)(rstride >> 7) * i);
                    const uint32_t q = p.cw_mul ? (__umulhi(row, p.cw_mul) >> p.cw_shift) : row;            // row / W
                    const uint32_t q2 = p.ch_mul ? (__umulhi(q, p.ch_mul) >> p.ch_shift) : q;               // ... / H
                    pix[i] = row < (uint32_t)p.M ? ((row - q * (uint32_t)p.conv_w) | ((q - q2 * (uint32_t)p.conv_h) << 16)) : 0xffffffffu;
                }
            }
            int xtap = 0, xkc = 0;
            for (int kb = 0; kb < num_k; ++kb) {
                int dy = 0, dx = 0;
                if (CONV3) {
                    dy = xtap / 3 - 1; dx = xtap % 3 - 1;
                    if (++xkc == p.conv_kpt) { xkc = 0; ++xtap; }
                }
                const int krem = p.K - kb * BK;                               // real columns left in this k-block
                const bool active = q * 8 < ceil_div(min(krem, BK), UMMA_K) * UMMA_K;   // the MMAs read this 8-column group
                const bool ga_on = gated && ka < krem, gb_on = gated && kb4 < krem;
                float4 ga[XR], gb[XR];
                if (gated) {                                 // issue the gate loads before blocking on the TMA
#pragma unroll
                    for (int i = 0; i < XR; ++i) {
                        ga[i] = (ga_on && i < nrows) ? ldg4(grow[i] + kb * BK + ka) : make_float4(1.f, 1.f, 1.f, 1.f);
                        gb[i] = (gb_on && i < nrows) ? ldg4(grow[i] + kb * BK + kb4) : make_float4(1.f, 1.f, 1.f, 1.f);
                    }
                }
                mbar_wait(full(s), ph);
                if (t == 0) trace_stamp(p.trace, xstep, 2);
                const uint32_t a = ring + s * stage_bytes;
                if (active) {
#pragma unroll
                    for (int b = 0; b < XR; b += RB) {
                        if (b < nrows) {
                            float4 va[RB], vb[RB];
#pragma unroll
                            for (int i = 0; i < RB; ++i) {
                                va[i] = lds128(a + src_off + (b + i) * rstride);
                                vb[i] = lds128(a + (src_off ^ 16u) + (b + i) * rstride);
                                if (CONV3) {     // the tap of this k-block lies outside the image for this row: zero padding
                                    const uint32_t xy = pix[b + i];
                                    const bool ok = xy != 0xffffffffu && (uint32_t)((int)(xy & 0xffffu) + dx) < (uint32_t)p.conv_w &&
                                                    (uint32_t)((int)(xy >> 16) + dy) < (uint32_t)p.conv_h;
                                    if (!ok) { va[i] = make_float4(0.f, 0.f, 0.f, 0.f); vb[i] = va[i]; }
                                }
                            }
                            __syncwarp(__activemask());      // in-place: every lane of the row has read before any lane writes
#pragma unroll
                            for (int i = 0; i < RB; ++i) {
                                f2_t xa[2] = {f2_pack(va[i].x, va[i].y), f2_pack(va[i].z, va[i].w)};
                                f2_t xb[2] = {f2_pack(vb[i].x, vb[i].y), f2_pack(vb[i].z, vb[i].w)};
                                if (gated) {
                                    const float4 g0 = ga[b + i], g1 = gb[b + i];
                                    xa[0] = f2_mul(xa[0], f2_pack(g0.x, g0.y)); xa[1] = f2_mul(xa[1], f2_pack(g0.z, g0.w));
                                    xb[0] = f2_mul(xb[0], f2_pack(g1.x, g1.y)); xb[1] = f2_mul(xb[1], f2_pack(g1.z, g1.w));
                                    f2_unpack(xa[0], va[i].x, va[i].y); f2_unpack(xa[1], va[i].z, va[i].w);
                                    f2_unpack(xb[0], vb[i].x, vb[i].y); f2_unpack(xb[1], vb[i].z, vb[i].w);
                                }
                                const uint32_t d = a + dst_off + (b + i) * rstride;
                                if (SPLIT) {
                                    uint32_t ha[2], la[2], hb[2], lb[2];
                                    split_f16x2(xa[0], ha[0], la[0]); split_f16x2(xa[1], ha[1], la[1]);
                                    split_f16x2(xb[0], hb[0], lb[0]); split_f16x2(xb[1], hb[1], lb[1]);
                                    sts128_u(d, swap ? hb[0] : ha[0], swap ? hb[1] : ha[1], swap ? ha[0] : hb[0], swap ? ha[1] : hb[1]);
                                    sts128_u(d + A_BOX_BYTES, swap ? lb[0] : la[0], swap ? lb[1] : la[1], swap ? la[0] : lb[0], swap ? la[1] : lb[1]);
                                } else {
                                    const uint32_t ha0 = pack_f16x2(va[i].x, va[i].y), ha1 = pack_f16x2(va[i].z, va[i].w);
                                    const uint32_t hb0 = pack_f16x2(vb[i].x, vb[i].y), hb1 = pack_f16x2(vb[i].z, vb[i].w);
                                    sts128_u(d, swap ? hb0 : ha0, swap ? hb1 : ha1, swap ? ha0 : hb0, swap ? ha1 : hb1);
                                }
                            }
                        }
                    }
                }
                fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (t == 0) trace_stamp(p.trace, xstep, 3);
                if (lane == 0) mbar_arrive(ready(s));
                ++xstep;
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ================================ epilogue ================================
        // warp -> TMEM lane group (warp % 4) x column share; 32 rows x 32 columns slabs go through a swizzled
        // shared-memory staging buffer and leave with one TMA store (coalesced, clipped at the M / N edges);
        // the residual slab arrives the same way.
        const int ew = warp - EPI_WARP0;            // 0..11
        const int lane_grp = warp & 3;              // TMEM lanes 32*lane_grp .. +31 are accessible to this warp
        const int share = ew >> 2;                  // the EPI_SPLIT warps of a lane group share the tile's column slabs
        const int nbuf = p.slabs_per_warp;
        const uint32_t my_staging = staging + (uint32_t)(ew * nbuf * SLAB_BYTES);
        const uint32_t my_ss = sstab + (uint32_t)(ew * SS_BYTES);
        const int sw = lane & 7;
        const int n_slabs = ceil_div(p.BN, 32);     // BN <= 96 -> <= 3 slabs -> one per warp of a lane group
        const uint32_t t_lane = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        uint32_t tcount = 0, res_phase = 0, slab_count = 0, it = 0, mb = 0, mph = 0;
        TileIter eti(blockIdx.x, gridDim.x, p.n_tiles);
        uint32_t rot = 0;                            // tcount % EPI_SPLIT
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount, eti.next(), rot = (rot + 1 == EPI_SPLIT) ? 0u : rot + 1) {
            const uint32_t acc = tcount & 1;
            const int m0 = eti.mt * BM, n0 = eti.nt * p.BN;
            const int row0 = m0 + lane_grp * 32;
            // slab sl belongs to share (sl + tile counter) % EPI_SPLIT: uneven slab counts even out over tiles
            int sl = share - (int)rot;
            if (sl < 0) sl += EPI_SPLIT;
            if (p.fixed_slabs) sl = share;
            const int c0 = sl * 32;
            const bool have = sl < n_slabs;                                  // this warp holds a slab of the tile
            if (!have) {
                // Narrow layers (N <= 64) leave one or two of a lane group's three warps without a slab. They still take part
                // in every accumulator hand-off (wait + arrive, nothing else): every waiter must observe every phase of an
                // mbarrier -- a warp that skipped tiles could fall two phases behind and mistake an older completion of the
                // same parity for its own (tried in round 2: deadlocks under load).
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    mbar_wait(main_full(mb), mph);
                    if (lane == 0) mbar_arrive(main_empty(mb));
                    if (++mb == NM) { mb = 0; mph ^= 1; }
                }
                if (SPLIT && lane == 0) mbar_arrive(tmem_empty(acc));
                continue;
            }
            const bool live = row0 < p.M && n0 + c0 < p.N;                   // ... that has rows / columns to store
            const bool wide = p.BN - c0 > 16;                                // the slab's second 16 columns exist in TMEM
            const uint32_t stage = my_staging + ((nbuf == 2) ? (slab_count & 1u) : 0u) * SLAB_BYTES;
            auto wait_staging_free = [&]() { if (nbuf == 2) tma_store_wait_read1(); else tma_store_wait_read0(); };
            // per-column scale / shift of this slab: one coalesced load per lane now, parked in shared memory after the
            // accumulation (columns beyond N get 0/0: their outputs are exact zeros and the TMA store clips them)
            float my_sc = 0.f, my_sh = 0.f;
            if (live && n0 + c0 + lane < p.N) { my_sc = __ldg(p.scale + n0 + c0 + lane); my_sh = __ldg(p.shift + n0 + c0 + lane); }
            if (has_res && live) {            // TMA-load the residual slab into the staging buffer it will be added in
                if (lane == 0) {
                    wait_staging_free();
                    mbar_expect_tx(res_bar(ew), SLAB_BYTES);
                    tma_load_2d(stage, &map_res, res_bar(ew), n0 + c0, row0);
                }
                __syncwarp();
            }

            // ---- fp32 (round-to-nearest) accumulation of the per-k-block main accumulators ----
            f2_t sum[16];                                  // 32 columns as adjacent pairs
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[j] = 0ull;
            for (int kb = 0; kb < num_k; ++kb, ++it) {
                if (ew == 0 && lane == 0 && kb == 0) trace_stamp(p.trace, it, 15);
                mbar_wait(main_full(mb), mph);
                if (ew == 0 && lane == 0) trace_stamp(p.trace, it, 8);
                tc_fence_after();
                {
                    float u[32];
                    const uint32_t col = t_lane + mb * MS + c0;
                    tmem_ld16_issue(col, u);
                    if (wide) tmem_ld16_issue(col + 16, u + 16);
                    else {
#pragma unroll
                        for (int j = 16; j < 32; ++j) u[j] = 0.f;
                    }
                    tmem_ld_wait();
                    // Every tcgen05.mma result is TRUNCATED to fp32 (measured): the k-block partial u
                    // is short by 0.5 ulp(u) in expectation for its last MMA, plus the earlier ones at their
                    // smaller magnitudes. Adding kappa * ulp(u) * sign(u) back removes the systematic part. For the
                    // default kappa = 1 that is "the next float away from zero", i.e. +1 on the bit pattern: one
                    // 64-bit integer add per column pair (no carry can cross the halves: the low word is never
                    // 0xffffffff) instead of a mask + FMA per element.
                    if (p.debias == 1.0f) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            sum[j] = f2_add(sum[j], f2_pack(u[2 * j], u[2 * j + 1]) + 0x0000000100000001ull);
                    } else {
                        const f2_t debias2 = f2_pack(p.debias * 1.1920929e-07f, p.debias * 1.1920929e-07f);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const f2_t up = f2_pack(u[2 * j], u[2 * j + 1]);
                            const f2_t pow2 = up & 0xff800000ff800000ull;            // sign * 2^exponent of each half
                            sum[j] = f2_add(sum[j], f2_fma(pow2, debias2, up));
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (ew == 0 && lane == 0) trace_stamp(p.trace, it, 9);
                if (lane == 0) mbar_arrive(main_empty(mb));
                if (++mb == NM) { mb = 0; mph ^= 1; }
            }
            if (ew == 0 && lane == 0) trace_stamp(p.trace, it - 1, 10);
            if (SPLIT) {   // tcgen05.commit covers ALL earlier MMAs: the last main_full also completed the correction terms
                {
                    float u[32];
                    const uint32_t col = t_lane + CORR0 + acc * (uint32_t)p.BN + c0;
                    tmem_ld16_issue(col, u);
                    if (wide) tmem_ld16_issue(col + 16, u + 16);
                    else {
#pragma unroll
                        for (int j = 16; j < 32; ++j) u[j] = 0.f;
                    }
                    tmem_ld_wait();
                    const f2_t inv = f2_pack(4.8828125e-4f, 4.8828125e-4f);      // 2^-11: the lo parts were scaled by 2^11
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[j] = f2_fma(f2_pack(u[2 * j], u[2 * j + 1]), inv, sum[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty(acc));
            }
            if (!live) continue;
            if (ew == 0 && lane == 0) trace_stamp(p.trace, it - 1, 11);

            // ---- scale/shift, activation, residual, store ----
            sts32(my_ss + lane * 4, my_sc);
            sts32(my_ss + 128 + lane * 4, my_sh);
            if (has_res) {
                mbar_wait(res_bar(ew), res_phase);
                res_phase ^= 1;
            } else {
                if (lane == 0) wait_staging_free();
            }
            __syncwarp();
            if (ew == 0 && lane == 0) trace_stamp(p.trace, it - 1, 12);
            f2_t sc[2], sh[2], rr[2] = {0ull, 0ull};       // software pipeline: the loads of step q+1 are issued before step q computes
            lds128_f2(my_ss, sc[0], sc[1]);
            lds128_f2(my_ss + 128, sh[0], sh[1]);
            if (has_res) lds128_f2(stage + (uint32_t)((lane * 8 + sw) * 16), rr[0], rr[1]);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                f2_t scn[2] = {0ull, 0ull}, shn[2] = {0ull, 0ull}, rn[2] = {0ull, 0ull};
                if (q < 7) {
                    lds128_f2(my_ss + (q + 1) * 16, scn[0], scn[1]);
                    lds128_f2(my_ss + 128 + (q + 1) * 16, shn[0], shn[1]);
                    if (has_res) lds128_f2(stage + (uint32_t)((lane * 8 + ((q + 1) ^ sw)) * 16), rn[0], rn[1]);
                }
                f2_t o[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    o[h] = f2_fma(sum[2 * q + h], sc[h], sh[h]);
                    if (act == 1) o[h] = f2_silu(o[h]);
                    else if (act == 2) o[h] = f2_relu(o[h]);
                    else if (act == 4) { float x0, x1; f2_unpack(o[h], x0, x1); o[h] = f2_pack(gelu_erf(x0), gelu_erf(x1)); }
                    if (has_res) o[h] = f2_add(o[h], rr[h]);
                    if (act == 16 + 2) o[h] = f2_relu(o[h]);
                }
                sts128_f2(stage + (uint32_t)((lane * 8 + (q ^ sw)) * 16), o[0], o[1]);
#pragma unroll
                for (int h = 0; h < 2; ++h) { sc[h] = scn[h]; sh[h] = shn[h]; rr[h] = rn[h]; }
            }
            if (ew == 0 && lane == 0) trace_stamp(p.trace, it - 1, 13);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) tma_store_2d(&map_out, stage, n0 + c0, row0);
            if (ew == 0 && lane == 0) trace_stamp(p.trace, it - 1, 14);
            ++slab_count;
        }
        if (lane == 0) tma_store_wait_all();
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

// 2-D row-major [rows, cols] tensor with a row pitch of `pitch` elements, box = [box_rows, 128 bytes], SWIZZLE_128B, zero OOB fill
static int make_map(CUtensorMap* map, const void* base, bool f16, int64_t rows, int64_t cols, int64_t pitch, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return ORBIT_ERR_UNSUPPORTED;
    const int esize = f16 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * esize};
    cuuint32_t box[2] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base),
                          dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ORBIT_OK : ORBIT_ERR_UNSUPPORTED;
}

}  // namespace tc

static float g_debias_kappa = 1.0f;   // one ulp of every promoted k-block partial (4 truncating MMAs: 0.5*(1+.75+.5+.25) ulp expected loss)
static unsigned* g_gemm_trace = nullptr;
static int g_narrow = 1;
static int g_fixed_slabs = 0;        // dev A/B switches (orbit_set_global_option): tc_fixed_slabs, tc_double_min_stages
static int g_wide_xf = 1;
static int g_double_min_stages = 3;  // double-buffered epilogue staging only when that still leaves this many ring stages
void set_tcgen05_tuning(int fixed_slabs, int double_min_stages) {
    if (fixed_slabs >= 0) g_fixed_slabs = fixed_slabs;
    if (double_min_stages >= 0) g_double_min_stages = double_min_stages;
}
void set_tcgen05_wide_xf(int v) { g_wide_xf = v; }
int get_tcgen05_wide_xf() { return g_wide_xf; }
void set_tcgen05_narrow(int on) { g_narrow = on; }
int get_tcgen05_narrow() { return g_narrow; }
void set_tcgen05_trace(unsigned* dev_buffer) { g_gemm_trace = dev_buffer; }
void set_tcgen05_debias(float kappa) { g_debias_kappa = kappa; }
float get_tcgen05_debias() { return g_debias_kappa; }

namespace {
struct Conv3Geom { int W, H, Cin; };      // implicit 3x3 stride-1 pad-1 convolution over an NHWC activation
void fast_div(uint32_t d, uint32_t* mul, uint32_t* shift) {      // x / d = umulhi(x, mul) >> shift for x < 2^31 (mul == 0: d == 1)
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    *mul = d <= 1 ? 0u : (uint32_t)(((1ull << (31 + l)) / d) + 1);
    *shift = d <= 1 ? 0u : (31 + l - 32);
}
}  // namespace

static int launch_tcgen05_impl(const float* A, const float* w_split, const float* scale, const float* shift,
                               const float* gate, const float* residual, float* out, int M, int N, int K,
                               int rows_per_frame, int act, int passes, const Conv3Geom* conv, cudaStream_t st) {
    using namespace tc;
    if (K % 4 || N % 4 || (passes != 1 && passes != 3)) return ORBIT_ERR_UNSUPPORTED;
    if (M <= 0) return ORBIT_OK;
    if (!conv && passes == 3 && (act == 0 || act == 1)) {       // the row-streaming kernel covers the small-K / small-N layer shapes
        const int rc = launch_pointwise_stream(A, w_split, scale, shift, gate, residual, out, M, N, K, rows_per_frame, act, st);
        if (rc != ORBIT_ERR_UNSUPPORTED) return rc;
    }
    Params p;
    p.scale = scale; p.shift = shift; p.gate = gate; p.has_residual = residual != nullptr;
    p.M = M; p.N = N; p.K = K; p.rows_per_frame = rows_per_frame; p.act = act;
    fast_div((uint32_t)std::max(rows_per_frame, 1), &p.rpf_mul, &p.rpf_shift);
    p.conv_w = p.conv_h = p.conv_kpt = 0;
    p.cw_mul = p.cw_shift = p.ch_mul = p.ch_shift = 0;
    if (conv) {
        if (conv->Cin % BK || K != 9 * conv->Cin || gate || passes != 3) return ORBIT_ERR_UNSUPPORTED;
        p.conv_w = conv->W; p.conv_h = conv->H; p.conv_kpt = conv->Cin / BK;
        fast_div((uint32_t)conv->W, &p.cw_mul, &p.cw_shift);
        fast_div((uint32_t)conv->H, &p.ch_mul, &p.ch_shift);
    }
    p.debias = passes == 3 ? g_debias_kappa : 0.f;
    p.trace = g_gemm_trace;
    // n-tiles of at most 96 columns (3 store slabs = one per epilogue warp of a lane group); with several n-tiles
    // BN must be a multiple of the 32-column store slab
    const int bn_max = 96;
    p.n_tiles = ceil_div(N, bn_max);
    p.BN = p.n_tiles > 1 ? ceil_div(ceil_div(N, p.n_tiles), 32) * 32 : ceil_div(N, 16) * 16;
    p.m_tiles = ceil_div(M, BM);
    p.b_tile_bytes = p.BN * BK * 2;
    const int stage_bytes = A_STAGE_BYTES + p.b_tile_bytes * (passes == 3 ? 2 : 1);
    const int bar_bytes = (3 * MAX_STAGES + 8 + NUM_EPI_WARPS) * 8 + 16;
    const int budget = 227 * 1024 - 1024 /*alignment slack*/ - bar_bytes - NUM_EPI_WARPS * SS_BYTES;
    const int n_slabs = ceil_div(p.BN, 32);
    p.fixed_slabs = (g_fixed_slabs && n_slabs < EPI_SPLIT) ? 1 : 0;
    p.staging_warps = p.fixed_slabs ? 4 * n_slabs : NUM_EPI_WARPS;
    // double-buffered epilogue staging when that still leaves a deep enough operand ring
    p.slabs_per_warp = (budget - 2 * p.staging_warps * SLAB_BYTES) / stage_bytes >= g_double_min_stages ? 2 : 1;
    const int staging_bytes = p.staging_warps * p.slabs_per_warp * SLAB_BYTES;
    p.stages = std::min(MAX_STAGES, (budget - staging_bytes) / stage_bytes);
    if (p.stages < 2) return ORBIT_ERR_UNSUPPORTED;
    const size_t smem = (size_t)staging_bytes + NUM_EPI_WARPS * SS_BYTES + (size_t)p.stages * stage_bytes + bar_bytes + 1024;

    const int Kp = (K + 7) / 8 * 8;                    // row pitch of the split weights (launch_weight_split)
    const __half* w16 = reinterpret_cast<const __half*>(w_split);
    CUtensorMap map_a, map_bhi, map_blo, map_out, map_res;
    int rc = conv ? make_map(&map_a, A, false, M, conv->Cin, conv->Cin, BM) : make_map(&map_a, A, false, M, K, K, BM);
    if (rc) return rc;
    rc = make_map(&map_bhi, w16, true, N, K, Kp, p.BN);
    if (rc) return rc;
    rc = make_map(&map_blo, w16 + (int64_t)N * Kp, true, N, K, Kp, p.BN);
    if (rc) return rc;
    rc = make_map(&map_out, out, false, M, N, N, 32);
    if (rc) return rc;
    rc = make_map(&map_res, residual ? residual : out, false, M, N, N, 32);
    if (rc) return rc;

    // specialisations for the shapes the backbones use; anything else runs the run-time-dispatch instance
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const Params);
    KernelFn fn = nullptr;
    int xfw = 4;
    const bool g = gate != nullptr, r = residual != nullptr;
    const bool narrow = K <= 32 && g_narrow;
    // every n-tile re-loads and re-splits its A k-blocks: with several n-tiles and a deep K the four transform warps (~1,700 clk
    // per k-block) are slower than the twelve MMAs (1,116 clk): eight transform warps there (g_wide_xf: 0 off, 1 Linear layers, 2 + SiLU)
    // (the implicit 3x3 convolution has K >= 576 and a transform that also masks the out-of-image taps: always eight)
    const int wide_xf = (g_wide_xf && (conv || p.n_tiles >= 2) && K >= 2 * BK) ? g_wide_xf : 0;
    if (conv) {
        if (act == 2 && !r && wide_xf) { fn = pw_tcgen05_kernel<true, 0, 2, 0, 8, false, true>; xfw = 8; }
        else if (act == 18 && r && wide_xf) { fn = pw_tcgen05_kernel<true, 0, 18, 1, 8, false, true>; xfw = 8; }
        else if (act == 2 && !r) fn = pw_tcgen05_kernel<true, 0, 2, 0, 4, false, true>;     // conv3x3 + ReLU (ResNet conv1 of a block, set encoder)
        else if (act == 18 && r) fn = pw_tcgen05_kernel<true, 0, 18, 1, 4, false, true>;    // BasicBlock: relu(bn(conv3x3) + identity)
        else if (act == 0 && !r) fn = pw_tcgen05_kernel<true, 0, 0, 0, 4, false, true>;     // plain conv3x3 (data gradient of the set encoder)
        else if (act == 1 && !r) fn = pw_tcgen05_kernel<true, 0, 1, 0, 4, false, true>;     // conv3x3 + SiLU (EfficientNet-V2 EdgeResidual expand)
        else return ORBIT_ERR_UNSUPPORTED;
    } else if (passes == 3) {
        if (g && act == 0 && !r && narrow) fn = pw_tcgen05_kernel<true, 1, 0, 0, 4, true>;           // first MBConv project (K = 32)
        else if (g && act == 0 && !r) { fn = pw_tcgen05_kernel<true, 1, 0, 0, 8, false>; xfw = 8; }   // MBConv project
        else if (g && act == 0 && r) { fn = pw_tcgen05_kernel<true, 1, 0, 1, 8, false>; xfw = 8; }    // ... + skip
        else if (!g && act == 1 && !r && narrow) fn = pw_tcgen05_kernel<true, 0, 1, 0, 4, true>;    // MBConv expand at 112x112 / 56x56 (K = 16, 24)
        else if (!g && act == 1 && !r && wide_xf == 2) { fn = pw_tcgen05_kernel<true, 0, 1, 0, 8, false>; xfw = 8; }
        else if (!g && act == 0 && !r && wide_xf) { fn = pw_tcgen05_kernel<true, 0, 0, 0, 8, false>; xfw = 8; }
        else if (!g && act == 0 && r && wide_xf) { fn = pw_tcgen05_kernel<true, 0, 0, 1, 8, false>; xfw = 8; }
        else if (!g && act == 4 && !r && wide_xf) { fn = pw_tcgen05_kernel<true, 0, 4, 0, 8, false>; xfw = 8; }
        else if (!g && act == 2 && !r && wide_xf) { fn = pw_tcgen05_kernel<true, 0, 2, 0, 8, false>; xfw = 8; }
        else if (!g && act == 18 && r && wide_xf) { fn = pw_tcgen05_kernel<true, 0, 18, 1, 8, false>; xfw = 8; }
        else if (!g && act == 1 && !r) fn = pw_tcgen05_kernel<true, 0, 1, 0, 4, false>;    // MBConv expand / conv_head (SiLU)
        else if (!g && act == 0 && !r) fn = pw_tcgen05_kernel<true, 0, 0, 0, 4, false>;    // Linear / downsample
        else if (!g && act == 0 && r) fn = pw_tcgen05_kernel<true, 0, 0, 1, 4, false>;     // Linear + residual (ViT), EdgeResidual project
        else if (!g && act == 2 && !r) fn = pw_tcgen05_kernel<true, 0, 2, 0, 4, false>;    // conv + ReLU
        else if (!g && act == 4 && !r) fn = pw_tcgen05_kernel<true, 0, 4, 0, 4, false>;    // Linear + GELU
        else if (!g && act == 18 && r) fn = pw_tcgen05_kernel<true, 0, 18, 1, 4, false>;   // BasicBlock: relu(bn(conv) + identity)
        else if (!g && act == 1 && r) fn = pw_tcgen05_kernel<true, 0, 1, 1, 4, false>;     // ConvBnAct + skip (EfficientNet-V2)
        else fn = pw_tcgen05_kernel<true, -1, -1, -1, 4, false>;
    } else {
        fn = pw_tcgen05_kernel<false, -1, -1, -1, 4, false>;
    }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        ORBIT_CUDA(cudaGetDevice(&dev));
        ORBIT_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    ORBIT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int grid = std::min(p.m_tiles * p.n_tiles, num_sms);
    fn<<<grid, num_threads(xfw), smem, st>>>(map_a, map_bhi, map_blo, map_out, map_res, p);
    ORBIT_RETURN_IF_LAUNCH_FAILED();
    return ORBIT_OK;
}

int launch_pointwise_tcgen05(const float* A, const float* w_split, const float* scale, const float* shift,
                             const float* gate, const float* residual, float* out, int M, int N, int K,
                             int rows_per_frame, int act, int passes, cudaStream_t st) {
    return launch_tcgen05_impl(A, w_split, scale, shift, gate, residual, out, M, N, K, rows_per_frame, act, passes, nullptr, st);
}

int launch_conv3x3_tcgen05(const float* x, const float* w_split, const float* scale, const float* shift, const float* residual,
                           float* out, int B, int H, int W, int Cin, int N, int act, cudaStream_t st) {
    if ((int64_t)B * H * W >= (1ll << 31) || W > 65535 || H > 65535) return ORBIT_ERR_UNSUPPORTED;
    const Conv3Geom g{W, H, Cin};
    return launch_tcgen05_impl(x, w_split, scale, shift, nullptr, residual, out, B * H * W, N, 9 * Cin, H * W, act, 3, &g, st);
}

}  // namespace orbit
