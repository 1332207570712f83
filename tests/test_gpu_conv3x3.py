"""GPU parity of the implicit-GEMM 3x3 convolution (C ABI orbit_conv3x3, implicit = 1: the tcgen05 kernel walks the nine taps
with row-shifted TMA boxes of the NHWC activation; no im2col matrix) against torch conv2d in float64 and against the explicit
im2col + GEMM path (implicit = 0). Shapes: the resnet18 BasicBlock convolutions (BASELINE.json configs 1 and 3) and the set
encoder's layers 2-5 (reference model/set_encoders.py:91-105)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [  # (B, H, W, Cin, Cout, act, residual)
    (3, 56, 56, 64, 64, 2, False), (3, 56, 56, 64, 64, 18, True), (2, 28, 28, 128, 128, 18, True), (5, 14, 14, 256, 256, 2, False),
    (9, 7, 7, 512, 512, 18, True), (2, 112, 112, 64, 64, 2, False), (2, 20, 12, 64, 96, 0, False), (1, 5, 3, 128, 40, 2, False),
    (37, 7, 7, 64, 64, 0, False), (3, 21, 21, 64, 64, 1, False),
]


def _run(x, w, scale, shift, res, act, implicit):
    from orbit_b200 import lib as L
    lib = L.load()
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    out = torch.empty(B, H, W, Cout, device=x.device)
    n = lib.orbit_conv3x3_scratch_floats(B, H, W, Cin, Cout, implicit)
    scratch = torch.empty(n, device=x.device)
    L.check(lib.orbit_conv3x3(L.ptr(x), L.ptr(w), L.ptr(scale), L.ptr(shift), L.ptr(res), L.ptr(out), B, H, W, Cin, Cout, act, implicit,
                              L.ptr(scratch), n, L.stream_ptr(x.device)), "orbit_conv3x3")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("B,H,W,Cin,Cout,act,residual", CASES)
def test_conv3x3_implicit_matches_torch(cuda_device, B, H, W, Cin, Cout, act, residual):
    g = torch.Generator().manual_seed(B * H + Cin + Cout)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5
    scale, shift = 1 + 0.1 * torch.randn(Cout, generator=g), 0.1 * torch.randn(Cout, generator=g)
    res = torch.randn(B, Cout, H, W, generator=g) if residual else None
    ref = F.conv2d(x.double(), w.double(), None, 1, 1) * scale.double()[None, :, None, None] + shift.double()[None, :, None, None]
    if act == 2:
        ref = ref.relu()
    elif act == 1:
        ref = ref * torch.sigmoid(ref)
    if residual:
        ref = ref + res.double()
    if act == 18:
        ref = ref.relu()
    xd = x.permute(0, 2, 3, 1).contiguous().to(cuda_device)
    rd = res.permute(0, 2, 3, 1).contiguous().to(cuda_device) if residual else None
    wd, sc, sh = w.to(cuda_device), scale.to(cuda_device), shift.to(cuda_device)
    got = _run(xd, wd, sc, sh, rd, act, 1).permute(0, 3, 1, 2).cpu().double()
    tol = 3e-6 * max(1.0, ref.abs().max().item())
    assert (got - ref).abs().max().item() <= tol, f"implicit conv3x3 differs by {(got - ref).abs().max().item():.2e} (tol {tol:.1e})"
    explicit = _run(xd, wd, sc, sh, rd, act, 0).permute(0, 3, 1, 2).cpu().double()
    assert (explicit - ref).abs().max().item() <= tol
    assert (got - explicit).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,H,W,k,stride,pad,act", [(3, 224, 224, 3, 1, 1, 2), (2, 224, 224, 7, 2, 3, 2), (5, 84, 84, 7, 2, 3, 2),
                                                    (4, 64, 64, 3, 1, 1, 0), (3, 37, 53, 3, 1, 1, 2), (2, 45, 31, 7, 2, 3, 0)])
def test_conv_first_matches_torch(cuda_device, B, H, W, k, stride, pad, act):
    """the direct first convolution on 3-channel NCHW frames (C ABI orbit_conv_first) against torch conv2d in float64"""
    from orbit_b200 import lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(H + W + k)
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(64, 3, k, k, generator=g) * (3 * k * k) ** -0.5
    scale, shift = 1 + 0.1 * torch.randn(64, generator=g), 0.1 * torch.randn(64, generator=g)
    ref = F.conv2d(x.double(), w.double(), None, stride, pad) * scale.double()[None, :, None, None] + shift.double()[None, :, None, None]
    if act == 2:
        ref = ref.relu()
    Ho, Wo = ref.shape[-2:]
    y = torch.full((B, Ho, Wo, 64), float('nan'), device=cuda_device)
    keep = [t.to(cuda_device) for t in (x, w, scale, shift)]
    L.check(lib.orbit_conv_first(*(L.ptr(t) for t in keep), L.ptr(y), B, H, W, k, stride, pad, act, L.stream_ptr(cuda_device)),
            "orbit_conv_first")
    torch.cuda.synchronize()
    got = y.permute(0, 3, 1, 2).cpu().double()
    assert (got - ref).abs().max().item() <= 3e-6 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,H,W,act", [(3, 224, 224, 1), (5, 84, 84, 1), (2, 85, 63, 1), (4, 64, 64, 0), (2, 33, 47, 2), (1, 16, 16, 1)])
def test_stem_matches_torch(cuda_device, B, H, W, act):
    """the EfficientNet stem (C ABI orbit_stem_conv: 3x3 stride 2, TF-SAME, 3 -> 32) against torch conv2d in float64"""
    from orbit_b200 import lib as L
    from oracle.backbones import tf_same_pad
    lib = L.load()
    g = torch.Generator().manual_seed(H + W + act)
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(32, 3, 3, 3, generator=g) * 27 ** -0.5
    scale, shift = 1 + 0.1 * torch.randn(32, generator=g), 0.1 * torch.randn(32, generator=g)
    (pt, pb), (pl, pr) = tf_same_pad(H, 3, 2), tf_same_pad(W, 3, 2)
    ref = F.conv2d(F.pad(x.double(), (pl, pr, pt, pb)), w.double(), None, 2, 0)
    ref = ref * scale.double()[None, :, None, None] + shift.double()[None, :, None, None]
    if act == 1:
        ref = ref * torch.sigmoid(ref)
    elif act == 2:
        ref = ref.relu()
    Ho, Wo = ref.shape[-2:]
    assert (Ho, Wo) == ((H + 1) // 2, (W + 1) // 2)
    y = torch.full((B, Ho, Wo, 32), float('nan'), device=cuda_device)
    keep = [t.to(cuda_device) for t in (x, w, scale, shift)]
    L.check(lib.orbit_stem_conv(*(L.ptr(t) for t in keep), L.ptr(y), B, H, W, act, L.stream_ptr(cuda_device)), "orbit_stem_conv")
    torch.cuda.synchronize()
    got = y.permute(0, 3, 1, 2).cpu().double()
    assert (got - ref).abs().max().item() <= 3e-6 * max(1.0, ref.abs().max().item())
