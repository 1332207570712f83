"""GPU parity of the pointwise-conv GEMM kernels (C ABI orbit_pointwise_conv) against an fp64 reference:
fp32 FFMA tiles (mode 0), tcgen05 3xTF32 (mode 1, must be fp32-grade) and tcgen05 1xTF32 (mode 2).
Shapes are the EfficientNet-B0 layer shapes (expand / project / head) incl. ragged M, K<32, N not a multiple of 16."""
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
    print(f"M={M} N={N} K={K}: max|err| ffma={errs[0]:.2e} 3xtf32={errs[1]:.2e} 1xtf32={errs[2]:.2e}")
    mag = max(1.0, ref.abs().max().item())
    assert errs[0] <= 1e-5 * mag
    assert errs[1] <= 1e-5 * mag          # 3xTF32 must be fp32-grade
    assert errs[2] <= 5e-3 * mag          # plain TF32: 10-bit mantissa inputs


def test_pointwise_rejects_bad_arguments(cuda_device):
    from orbit_b200 import lib as L
    lib = L.load()
    x = torch.zeros(8, 8, device=cuda_device)
    assert lib.orbit_pointwise_conv(None, L.ptr(x), L.ptr(x), L.ptr(x), None, None, L.ptr(x), 8, 8, 8, 1, 0, 0, None, None) == -1
    assert lib.orbit_pointwise_conv(L.ptr(x), L.ptr(x), L.ptr(x), L.ptr(x), None, None, L.ptr(x), 8, 8, 8, 1, 0, 1, None, None) == -1
    assert lib.orbit_pointwise_conv(L.ptr(x), L.ptr(x), L.ptr(x), L.ptr(x), None, None, L.ptr(x), 8, 8, 8, 1, 7, 0, None, None) == -1


@pytest.mark.parametrize("shape", [(7 * 14 * 14, 112, 672, 14 * 14, 0, True, True), (9 * 7 * 7, 192, 1152, 7 * 7, 0, True, False),
                                   (3 * 28 * 28, 40, 240, 28 * 28, 0, True, True), (260, 80, 480, 65, 0, True, False)],
                         ids=lambda s: f"M{s[0]}_N{s[1]}_K{s[2]}")
def test_gated_projection_variants_agree(cuda_device, shape):
    """The A/B variants of the gated-projection GEMM (A operand in shared vs tensor memory, three instructions per k-step vs
    the merged [B_hi;B_lo] product) compute the same fp32-grade result: each within 2e-6 relative of fp64."""
    from orbit_b200 import lib as L
    lib = L.load()
    M, N, K, rpf, act, gated, residual = shape
    g = torch.Generator().manual_seed(M * 3 + N)
    A = torch.randn(M, K, generator=g).to(cuda_device)
    W = (torch.randn(N, K, generator=g) * K ** -0.5).to(cuda_device)
    scale = (1 + 0.1 * torch.randn(N, generator=g)).to(cuda_device)
    shift = (0.1 * torch.randn(N, generator=g)).to(cuda_device)
    gate = torch.rand((M + rpf - 1) // rpf, K, generator=g).to(cuda_device)
    res = torch.randn(M, N, generator=g).to(cuda_device) if residual else None
    ref = (A.double() * gate.double().repeat_interleave(rpf, dim=0)[:M]) @ W.double().t() * scale.double() + shift.double()
    if residual:
        ref = ref + res.double()
    outs = {}
    try:
        for atm in (0, 1):
            for merge in (0, 1):
                L.check(lib.orbit_set_global_option(b"tc_a_in_tmem", atm), "set tc_a_in_tmem")
                L.check(lib.orbit_set_global_option(b"tc_merge", merge), "set tc_merge")
                out = _run(1, A, W, scale, shift, gate, res, rpf, act)
                err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
                outs[(atm, merge)] = err
                assert err <= 2e-6, f"a_in_tmem={atm} merge={merge}: relative error {err:.2e}"
    finally:
        lib.orbit_set_global_option(b"tc_a_in_tmem", 1)
        lib.orbit_set_global_option(b"tc_merge", 0)
    print({k: f"{v:.1e}" for k, v in outs.items()})
