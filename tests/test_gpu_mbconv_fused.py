"""GPU parity of the fused MBConv front half (C ABI orbit_mbconv_expand_dw: expand 1x1 + bn1 + SiLU -> depthwise kxk + bn2
+ SiLU, the 6x-expanded tensor kept in shared memory) against torch conv2d with TF-SAME padding, and against the unfused
engine path on whole episodes. Reference op sites: timm InvertedResidual conv_pw/bn1/conv_dw/bn2 as restated in
oracle/backbones.py (FiLM site bn2: model/film.py:43-44)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [  # (B, H, Cin, C, k, stride): the three EfficientNet-B0 blocks at 224 px, then odd sizes / the 84 px pyramid
    (2, 112, 16, 96, 3, 2), (2, 56, 24, 144, 3, 1), (3, 56, 24, 144, 5, 2),
    (3, 42, 16, 96, 3, 2), (2, 21, 24, 144, 3, 1), (2, 21, 24, 144, 5, 2), (1, 7, 24, 144, 5, 1), (5, 9, 16, 64, 3, 1),
    (2, 48, 16, 96, 5, 2), (1, 33, 24, 40, 3, 2),
]


@pytest.mark.parametrize("B,H,Cin,C,k,stride", CASES)
def test_fused_expand_depthwise_matches_torch(cuda_device, B, H, Cin, C, k, stride):
    from orbit_b200 import lib as L
    from oracle.backbones import tf_same_pad
    lib = L.load()
    g = torch.Generator().manual_seed(H * C + k + Cin)
    x = torch.randn(B, Cin, H, H, generator=g)
    we = torch.randn(C, Cin, 1, 1, generator=g) * Cin ** -0.5
    s1, h1 = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    wd = torch.randn(C, 1, k, k, generator=g) * 0.3
    s2, h2 = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    e = F.conv2d(x.double(), we.double()) * s1.double()[None, :, None, None] + h1.double()[None, :, None, None]
    e = e * torch.sigmoid(e)
    if stride == 1:
        ref = F.conv2d(e, wd.double(), None, 1, (k - 1) // 2, 1, C)
    else:
        pt, pb = tf_same_pad(H, k, stride)
        ref = F.conv2d(F.pad(e, (pt, pb, pt, pb)), wd.double(), None, stride, 0, 1, C)   # zero padding of the EXPANDED tensor
    ref = ref * s2.double()[None, :, None, None] + h2.double()[None, :, None, None]
    ref = (ref * torch.sigmoid(ref)).float()
    Ho = ref.shape[-1]
    dev = cuda_device
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    y = torch.full((B, Ho, Ho, C), float('nan'), device=dev)
    nparts = lib.orbit_mbconv_partial_floats(B, H, H, C, k, stride)
    partial = torch.full((nparts,), float('nan'), device=dev)
    scratch = torch.empty(k * k * C, device=dev)
    keep = [t.to(dev) for t in (we.reshape(C, Cin).contiguous(), s1, h1, wd, s2, h2)]   # raw pointers cross the ABI
    L.check(lib.orbit_mbconv_expand_dw(L.ptr(xd), *(L.ptr(t) for t in keep), L.ptr(y), L.ptr(partial), L.ptr(scratch),
                                       B, H, H, Cin, C, k, stride, L.stream_ptr(dev)), "orbit_mbconv_expand_dw")
    torch.cuda.synchronize()
    got = y.permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs().max().item()
    print(f"fused expand+dw B={B} H={H} {Cin}->{C} k={k} s={stride}: max|err|={err:.2e} max|ref|={ref.abs().max():.2f}")
    assert err <= 1e-5 * max(1.0, ref.abs().max().item())
    groups = lib.orbit_mbconv_partial_groups(H, H, Cin, C, k, stride)
    sums = partial[:B * groups * C].view(B, groups, C).sum(1).cpu()
    assert (sums - ref.sum((2, 3))).abs().max().item() <= 1e-4 * max(1.0, ref.sum((2, 3)).abs().max().item())


def test_fused_and_unfused_engine_paths_agree(cuda_device, oracle_effnet):
    """The engine option fuse_mbconv only changes WHERE the expanded tensor lives: features agree to fp32 rounding with
    the layer-at-a-time path and with the oracle."""
    import orbit_b200
    m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 2, 256, False, 16)
    m.load_state_dict(oracle_effnet.state_dict(), strict=True)
    m._set_device(cuda_device)
    m._send_to_device()
    m.set_test_mode(True)
    fe = m.feature_extractor
    x = torch.randn(5, 3, 224, 224, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        ref = oracle_effnet.extractor(x)
    assert fe.get_option('fuse_mbconv') == 1          # default: the stride-2 blocks (1.0 and 2.0)
    outs = {}
    for mode in (1, 2, 0):                            # stride-2 blocks / all three 16-24-channel blocks / layer at a time
        fe.set_option('fuse_mbconv', mode)
        outs[mode] = fe(x.to(cuda_device)).cpu()
    fe.set_option('fuse_mbconv', 1)
    scale = max(1.0, ref.abs().max().item())
    print({m: f"{(o - ref).abs().max():.2e}" for m, o in outs.items()}, f"fused(2) vs unfused {(outs[2] - outs[0]).abs().max():.2e}")
    for mode, out in outs.items():
        assert (out - ref).abs().max().item() <= 5e-5 * scale, mode
    assert not torch.equal(outs[2], outs[0])          # the fused path really ran (fp32 summation order differs)
