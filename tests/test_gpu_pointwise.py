"""GPU parity of the pointwise-conv GEMM kernels (C ABI orbit_pointwise_conv) against an fp64 reference:
fp32 FFMA tiles (mode 0), tcgen05 FP16x3 (mode 1, must be fp32-grade) and one plain fp16 product (mode 2, the `fast` mode).
Shapes are the EfficientNet-B0 layer shapes (expand / project / head) incl. ragged M, K<64, K % 8 == 4, N not a multiple of 16."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, K, rows_per_frame, act, gated, residual)
    (2 * 112 * 112, 96, 16, 112 * 112, 1, False, False),   # blocks.1.0 expand
    (3 * 56 * 56, 24, 96, 56 * 56, 0, True, False),        # blocks.1.0 project (N=24)
    (2 * 56 * 56, 24, 144, 56 * 56, 0, True, True),        # blocks.1.1 project + residual
    (5 * 28 * 28, 40, 144, 28 * 28, 0, True, False),       # N=40
    (3 * 14 * 14, 480, 80, 14 * 14, 1, False, False),      # expand, 4 n-tiles
    (7 * 14 * 14, 112, 672, 14 * 14, 0, True, True),       # deep K
    (9 * 7 * 7, 1152, 192, 7 * 7, 1, False, False),        # ragged M (441 rows)
    (9 * 7 * 7, 1280, 320, 7 * 7, 1, False, False),        # conv_head
    (130, 16, 32, 65, 0, True, False),                     # blocks.0.0 project, tiny
    (1, 1280, 320, 1, 2, False, False),                    # single row, ReLU
    (300, 64, 40, 100, 1, False, False),                   # K = 40: third k-step half empty
    (300, 48, 36, 100, 0, True, True),                     # K % 8 == 4: padded weight pitch
    (517, 40, 12, 517, 0, False, False),                   # K < 16
    (1000, 200, 132, 250, 2, False, False),                # K = 2 k-blocks + 4
    (148 * 128 * 7 + 77, 24, 96, 3136, 0, True, False),    # 7+ tiles per CTA, one slab per tile: epilogue warps take turns
    (148 * 128 * 4, 40, 144, 784, 0, True, True),          # two slabs per tile, residual, several tiles per CTA
    (148 * 128 * 5, 16, 32, 12544, 0, True, False),        # narrow-K transform mapping, many tiles per CTA
    (148 * 128 * 3 + 5, 96, 16, 12544, 1, False, False),   # K = 16 (two lanes per row), SiLU
]


def _run(mode, A, W, scale, shift, gate, res, rows_per_frame, act):
    from orbit_b200 import lib as L
    lib = L.load()
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, device=A.device)
    wsplit = torch.empty(2 * N * K, device=A.device)
    L.check(lib.orbit_pointwise_conv(L.ptr(A), L.ptr(W), L.ptr(scale), L.ptr(shift), L.ptr(gate), L.ptr(res), L.ptr(out),
                                     M, N, K, rows_per_frame, act, mode, L.ptr(wsplit), L.stream_ptr(A.device)),
            "orbit_pointwise_conv")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("shape", SHAPES, ids=[f"M{s[0]}_N{s[1]}_K{s[2]}" for s in SHAPES])
def test_pointwise_modes(cuda_device, shape):
    M, N, K, rpf, act, gated, residual = shape
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(cuda_device)
    W = (torch.randn(N, K, generator=g) * K ** -0.5).to(cuda_device)
    scale = (1 + 0.1 * torch.randn(N, generator=g)).to(cuda_device)
    shift = (0.1 * torch.randn(N, generator=g)).to(cuda_device)
    frames = (M + rpf - 1) // rpf
    gate = torch.rand(frames, K, generator=g).to(cuda_device) if gated else None
    res = torch.randn(M, N, generator=g).to(cuda_device) if residual else None
    a64 = A.double()
    if gated:
        a64 = a64 * gate.double().repeat_interleave(rpf, dim=0)[:M]
    ref = a64 @ W.double().t() * scale.double() + shift.double()
    if act == 1:
        ref = ref * torch.sigmoid(ref)
    elif act == 2:
        ref = ref.clamp_min(0)
    if residual:
        ref = ref + res.double()
    errs = {}
    for mode in (0, 1, 2):
        out = _run(mode, A, W, scale, shift, gate, res, rpf, act)
        errs[mode] = (out.double() - ref).abs().max().item()
    print(f"M={M} N={N} K={K}: max|err| ffma={errs[0]:.2e} fp16x3={errs[1]:.2e} fp16x1={errs[2]:.2e}")
    mag = max(1.0, ref.abs().max().item())
    assert errs[0] <= 1e-5 * mag
    assert errs[1] <= 1e-5 * mag          # FP16x3 must be fp32-grade
    assert errs[2] <= 5e-3 * mag          # one fp16 product: 11-bit significands


def test_pointwise_rejects_bad_arguments(cuda_device):
    from orbit_b200 import lib as L
    lib = L.load()
    x = torch.zeros(8, 8, device=cuda_device)
    assert lib.orbit_pointwise_conv(None, L.ptr(x), L.ptr(x), L.ptr(x), None, None, L.ptr(x), 8, 8, 8, 1, 0, 0, None, None) == -1
    assert lib.orbit_pointwise_conv(L.ptr(x), L.ptr(x), L.ptr(x), L.ptr(x), None, None, L.ptr(x), 8, 8, 8, 1, 0, 1, None, None) == -1
    assert lib.orbit_pointwise_conv(L.ptr(x), L.ptr(x), L.ptr(x), L.ptr(x), None, None, L.ptr(x), 8, 8, 8, 1, 7, 0, None, None) == -1


@pytest.mark.parametrize("a_scale,w_scale", [(1e-4, 1.0), (1.0, 1e-3), (300.0, 1.0), (1e-3, 30.0)])
def test_fp16x3_dynamic_range(cuda_device, a_scale, w_scale):
    """The fp16 split keeps fp32-grade RELATIVE accuracy over the magnitudes activations and weights take: small values
    go through the 2^11-scaled lo part, so nothing is lost below fp16's normal range (2^-14)."""
    M, N, K = 1024, 96, 672
    g = torch.Generator().manual_seed(7)
    A = (torch.randn(M, K, generator=g) * a_scale).to(cuda_device)
    W = (torch.randn(N, K, generator=g) * K ** -0.5 * w_scale).to(cuda_device)
    one, zero = torch.ones(N, device=cuda_device), torch.zeros(N, device=cuda_device)
    ref = A.double() @ W.double().t()
    out = _run(1, A, W, one, zero, None, None, M, 0)
    rel = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    assert rel <= 2e-6, f"relative error {rel:.2e} at |a|~{a_scale}, |w|~{w_scale}"


def test_fp16x3_is_unbiased(cuda_device):
    """The tensor core truncates every accumulation; the per-k-block promotion + 1-ulp de-bias must leave no systematic
    error (a bias compounds over the network's 33 GEMM layers: DESIGN.md section 4)."""
    M, N, K = 4096, 96, 1152
    g = torch.Generator().manual_seed(11)
    A = torch.rand(M, K, generator=g).to(cuda_device)            # all-positive operands: worst case for truncation bias
    W = (torch.rand(N, K, generator=g) / K).to(cuda_device)
    one, zero = torch.ones(N, device=cuda_device), torch.zeros(N, device=cuda_device)
    ref = A.double() @ W.double().t()
    out = _run(1, A, W, one, zero, None, None, M, 0)
    rel = (out.double() - ref) / ref
    print(f"mean relative error {rel.mean().item():.2e}, rms {rel.pow(2).mean().sqrt().item():.2e}")
    assert abs(rel.mean().item()) <= 1.5e-7      # all-positive worst case: -1.0e-7 measured (kappa = 1 under-corrects it, over-corrects mixed signs by +2e-8)
    assert rel.pow(2).mean().sqrt().item() <= 1.5e-7


@pytest.mark.parametrize("K,N,gated", [(96, 24, True), (32, 16, True), (24, 144, False), (40, 240, False), (16, 96, False)])
def test_streaming_gemm_is_unbiased_and_matches_tcgen05(cuda_device, K, N, gated):
    """The row-streaming kernel (csrc/gemm_stream.cu) of the small-K / small-N layers: no systematic error on all-positive
    operands (hi.hi partial sums are promoted per k-step with round-to-nearest), and agreement with the tcgen05 kernel."""
    from orbit_b200 import lib as L
    lib = L.load()
    M, rpf = 50000, 3136
    g = torch.Generator().manual_seed(K + N)
    A = torch.rand(M, K, generator=g).to(cuda_device)
    W = (torch.rand(N, K, generator=g) / K).to(cuda_device)
    one, zero = torch.ones(N, device=cuda_device), torch.zeros(N, device=cuda_device)
    gate = (0.5 + torch.rand((M + rpf - 1) // rpf, K, generator=g)).to(cuda_device) if gated else None
    a64 = A.double() * gate.double().repeat_interleave(rpf, dim=0)[:M] if gated else A.double()
    ref = a64 @ W.double().t()
    outs = {}
    for stream_on in (1, 0):
        assert lib.orbit_set_global_option(b'tc_stream', stream_on) == 0
        outs[stream_on] = _run(1, A, W, one, zero, gate, None, rpf, 0)
    assert lib.orbit_set_global_option(b'tc_stream', 1) == 0
    rel = (outs[1].double() - ref) / ref
    print(f"K={K} N={N}: streaming mean relative error {rel.mean().item():.2e}, rms {rel.pow(2).mean().sqrt().item():.2e}; "
          f"max |stream - tcgen05| / |ref| = {((outs[1] - outs[0]).double() / ref).abs().max().item():.2e}")
    assert abs(rel.mean().item()) <= 1.5e-7 and rel.pow(2).mean().sqrt().item() <= 1.5e-7
    assert ((outs[1] - outs[0]).double() / ref).abs().max().item() <= 1e-6
