"""GPU parity of the depthwise conv kernel (C ABI orbit_depthwise_conv) against torch's conv2d with TF-SAME padding,
on the EfficientNet-B0 depthwise shapes incl. odd sizes (84 px pyramid) and the fused SE partial sums."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [  # (B, H, C, k, stride)
    (2, 112, 32, 3, 1), (2, 112, 96, 3, 2), (3, 56, 144, 5, 2), (3, 28, 240, 5, 1), (2, 28, 240, 3, 2),
    (4, 14, 480, 3, 1), (3, 14, 672, 5, 1), (5, 14, 672, 5, 2), (7, 7, 1152, 5, 1), (3, 7, 1152, 3, 1),
    (2, 21, 144, 5, 2), (2, 11, 240, 3, 2), (3, 3, 1152, 5, 1), (1, 42, 96, 3, 2),
    (5, 7, 96, 5, 1), (9, 7, 672, 5, 1), (3, 14, 480, 5, 1),   # staged 5x5 kernel: 32-channel tail chunk, odd frame counts
    (700, 7, 1152, 5, 1), (300, 14, 672, 5, 1),               # ... several double-buffer iterations per block
]


def test_depthwise_staged_and_register_kernels_agree(cuda_device):
    """5x5 stride 1 at 14x14 / 7x7 runs the shared-memory-staged kernel (dw5s_kernel); with the switch off, dw2_kernel: both
    must match torch (the staged path is the default, so CASES above exercise it; this runs the other one)"""
    from orbit_b200 import lib as L
    lib = L.load()
    assert lib.orbit_set_global_option(b'dw5_staged', 0) == 0
    try:
        test_depthwise_matches_torch(cuda_device, 5, 14, 672, 5, 1)
        test_depthwise_matches_torch(cuda_device, 9, 7, 1152, 5, 1)
    finally:
        assert lib.orbit_set_global_option(b'dw5_staged', 1) == 0


@pytest.mark.parametrize("B,H,C,k,stride", CASES)
def test_depthwise_matches_torch(cuda_device, B, H, C, k, stride):
    from orbit_b200 import lib as L
    from oracle.backbones import tf_same_pad
    lib = L.load()
    g = torch.Generator().manual_seed(H * C + k)
    x = torch.randn(B, C, H, H, generator=g)
    w = torch.randn(C, 1, k, k, generator=g) * 0.3
    scale, shift = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    if stride == 1:
        ref = F.conv2d(x, w, None, 1, (k - 1) // 2, 1, C)
    else:
        pt, pb = tf_same_pad(H, k, stride)
        ref = F.conv2d(F.pad(x, (pt, pb, pt, pb)), w, None, stride, 0, 1, C)
    ref = ref * scale[None, :, None, None] + shift[None, :, None, None]
    ref = ref * torch.sigmoid(ref)
    Ho = ref.shape[-1]
    xd = x.permute(0, 2, 3, 1).contiguous().to(cuda_device)
    y = torch.empty(B, Ho, Ho, C, device=cuda_device)
    nparts = lib.orbit_depthwise_partial_floats(B, H, H, C, k, stride)
    partial = torch.full((nparts,), float('nan'), device=cuda_device)
    scratch = torch.empty(k * k * C, device=cuda_device)
    wd, scd, shd = w.to(cuda_device), scale.to(cuda_device), shift.to(cuda_device)   # keep alive: raw pointers cross the ABI
    L.check(lib.orbit_depthwise_conv(L.ptr(xd), L.ptr(wd), L.ptr(scd), L.ptr(shd),
                                     L.ptr(y), L.ptr(partial), L.ptr(scratch), B, H, H, C, k, stride, 1, L.stream_ptr(cuda_device)),
            "orbit_depthwise_conv")
    torch.cuda.synchronize()
    got = y.permute(0, 3, 1, 2).cpu()
    assert (got - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    sums = partial.view(B, -1, C).sum(1).cpu()
    assert (sums - ref.sum((2, 3))).abs().max().item() <= 1e-4 * max(1.0, ref.sum((2, 3)).abs().max().item())
