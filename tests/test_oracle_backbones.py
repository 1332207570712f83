"""The oracle's backbone restatements (oracle/backbones.py) against torchvision's INDEPENDENT implementations.

timm 0.6.12 (reference ``model/feature_extractors.py:39-58``, ``requirements.txt:6``) is not installable offline, so the
backbone arithmetic cannot be pinned against timm itself. torchvision ships the same published architectures written
by other people; with the state dict copied over, the restatement must give the same features:

* ``efficientnet_b0``: torchvision's MBConv stack == the oracle's EfficientNet once the two documented differences are
  neutralised -- BatchNorm eps (timm tf_ variant 1e-3, torchvision 1e-5: overridden on the torchvision side) and the
  TF 'SAME' asymmetric padding of the stride-2 convolutions (switched to symmetric on the oracle side; the TF-SAME
  padding formula itself is checked against the table in SURVEY.md Appendix A).
* ``vit_b_32``: torchvision ``vit_b_32`` with its classification head removed == the oracle's VisionTransformer.
* parameter counts equal timm's published ``num_classes=0`` counts.
"""
import pytest
import torch
import torchvision

from oracle import backbones as B


def _copy_positional(dst: torch.nn.Module, src: torch.nn.Module):
    """Both nets register their layers in forward order: copy tensor i -> tensor i, checking every shape."""
    d, s = dst.state_dict(), src.state_dict()
    assert len(d) == len(s)
    for (kd, vd), (ks, vs) in zip(d.items(), s.items()):
        assert vd.shape == vs.shape, f"{kd} {tuple(vd.shape)} vs {ks} {tuple(vs.shape)}"
        vd.copy_(vs)


def test_tf_same_padding_table():
    # SURVEY.md Appendix A (timm Conv2dSame): (input, kernel, stride) -> (before, after)
    table = {(224, 3, 2): (0, 1), (112, 3, 2): (0, 1), (56, 5, 2): (1, 2), (28, 3, 2): (0, 1), (14, 5, 2): (1, 2),
             (84, 3, 2): (0, 1), (42, 3, 2): (0, 1), (21, 5, 2): (2, 2), (11, 3, 2): (1, 1), (6, 5, 2): (1, 2)}
    for (size, k, s), want in table.items():
        assert B.tf_same_pad(size, k, s) == want, (size, k, s)


@torch.no_grad()
def test_efficientnet_b0_equals_torchvision():
    torch.manual_seed(0)
    tv = torchvision.models.efficientnet_b0(weights=None).eval()
    for m in tv.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eps = 1e-3
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.normal_(1, 0.1)
            m.bias.normal_(0, 0.1)
    tv.classifier = torch.nn.Identity()
    ours = B.EfficientNet().eval()
    assert sum(p.numel() for p in ours.parameters()) == 4_007_548     # timm tf_efficientnet_b0(num_classes=0)
    _copy_positional(ours, tv)
    for m in ours.modules():
        if isinstance(m, B.Conv2dSame):
            m.tf_same = False
    x = torch.randn(3, 3, 96, 96)
    want = tv(x)
    got = ours(x)
    assert got.shape == want.shape == (3, 1280)
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


@torch.no_grad()
def test_efficientnet_tf_same_differs_only_on_stride2():
    """The TF-SAME variant is the symmetric one with the input shifted by the asymmetric pad: on an even-sized input a
    stride-2 'SAME' conv equals a symmetric conv applied to the input rolled by one pixel (interior pixels)."""
    torch.manual_seed(1)
    conv = B.Conv2dSame(8, 8, 3, 2, groups=8, bias=False)
    x = torch.randn(2, 8, 16, 16)
    same = conv(x)
    ref = torch.nn.functional.conv2d(torch.nn.functional.pad(x, (0, 1, 0, 1)), conv.weight, None, 2, 0, 1, 8)
    assert torch.equal(same, ref)
    conv5 = B.Conv2dSame(4, 4, 5, 2, groups=4, bias=False)
    x = torch.randn(1, 4, 14, 14)
    ref5 = torch.nn.functional.conv2d(torch.nn.functional.pad(x, (1, 2, 1, 2)), conv5.weight, None, 2, 0, 1, 4)
    assert torch.equal(conv5(x), ref5)


_VIT_KEYS = (('class_token', 'cls_token'), ('conv_proj.', 'patch_embed.proj.'), ('encoder.pos_embedding', 'pos_embed'),
             ('encoder.layers.encoder_layer_', 'blocks.'), ('.ln_1.', '.norm1.'), ('.ln_2.', '.norm2.'),
             ('.self_attention.in_proj_', '.attn.qkv.'), ('.self_attention.out_proj.', '.attn.proj.'),
             ('.mlp.0.', '.mlp.fc1.'), ('.mlp.3.', '.mlp.fc2.'), ('encoder.ln.', 'norm.'))


@torch.no_grad()
def test_vit_b_32_equals_torchvision():
    torch.manual_seed(0)
    tv = torchvision.models.vit_b_32(weights=None).eval()
    tv.heads = torch.nn.Identity()
    for p in tv.parameters():               # torchvision zero-initialises several tensors: make all of them matter
        p.normal_(0, 0.02)
    for m in tv.modules():
        if isinstance(m, torch.nn.LayerNorm):
            m.weight.add_(1.0)
    remapped = {}
    for k, v in tv.state_dict().items():
        for a, b in _VIT_KEYS:
            k = k.replace(a, b)
        remapped[k] = v
    ours = B.build('vit_b_32').eval()
    assert sum(p.numel() for p in ours.parameters()) == 87_455_232    # timm vit_base_patch32_224(num_classes=0)
    missing, unexpected = ours.load_state_dict(remapped, strict=True)
    assert not missing and not unexpected
    x = torch.randn(2, 3, 224, 224)
    want, got = tv(x), ours(x)
    assert got.shape == want.shape == (2, 768)
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


def _vit_params(d, depth=12, patch=32, tokens=50):
    """Closed form of timm VisionTransformer(num_classes=0): patch conv + cls + pos + depth x (2 LN, qkv, proj, fc1, fc2) + norm."""
    return (3 * patch * patch * d + d) + d + tokens * d + depth * (12 * d * d + 13 * d) + 2 * d


def test_parameter_counts_of_the_other_extractors():
    assert sum(p.numel() for p in B.build('vit_s_32').parameters()) == _vit_params(384)
    assert sum(p.numel() for p in B.build('vit_b_32').parameters()) == _vit_params(768) == 87_455_232
    assert sum(p.numel() for p in B.build('vit_b_32_clip').parameters()) == _vit_params(768) + 2 * 768   # + norm_pre
    assert sum(p.numel() for p in B.build('resnet18').parameters()) == 11_689_512 - 513_000              # torchvision - fc


@pytest.mark.parametrize('name,size', [('efficientnet_b0', 64), ('resnet18', 64)])
@torch.no_grad()
def test_eval_mode_is_batch_independent(name, size):
    """The extractor is frame-wise in eval mode -- the property predict_video and chunked passes rely on."""
    m = B.seeded_init(B.build(name), calib_input=torch.randn(8, 3, size, size, generator=torch.Generator().manual_seed(3)))
    x = torch.randn(4, 3, size, size, generator=torch.Generator().manual_seed(4))
    full = m(x)
    for i in range(4):
        assert (m(x[i:i + 1]) - full[i:i + 1]).abs().max().item() <= 5e-5 * max(1.0, full.abs().max().item())   # MKL blocking varies with the batch size
