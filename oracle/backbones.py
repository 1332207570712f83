"""Backbone restatements (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference builds its extractors with timm==0.6.12 (reference
``model/feature_extractors.py:31-66``); timm is a third-party dependency whose source is not
in /root/reference and is not installable offline, so the *published* architectures are
restated here in plain PyTorch.  Module / parameter names follow timm's state-dict keys
because the reference addresses FiLM tensors by ``<module path>.weight/.bias`` strings
(reference ``model/film.py:68-74``) and ORBIT checkpoints are keyed that way.

* ``EfficientNet``  = timm ``tf_efficientnet_b0(num_classes=0)``: TF "SAME" asymmetric
  padding on stride-2 convs, BN eps 1e-3, SiLU, SE reduce = block-input-channels/4.
* ``VisionTransformer`` = timm ``vit_{small,base}_patch32_224*``(num_classes=0), token pooling.
* ``resnet18`` = torchvision resnet18 with ``fc = Identity`` (BASELINE.json extension; the
  reference at this commit has no resnet18).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- EfficientNet
def tf_same_pad(size: int, k: int, s: int):
    """TF 'SAME' padding (before, after) for one spatial dim."""
    total = max((math.ceil(size / s) - 1) * s + k - size, 0)
    return total // 2, total - total // 2


class Conv2dSame(nn.Conv2d):
    """Conv with TensorFlow 'SAME' padding computed from the input size (``tf_same=False`` gives
    the symmetric padding of the non-tf variants; used only to cross-check against torchvision)."""
    tf_same = True

    def forward(self, x):
        k, s = self.kernel_size[0], self.stride[0]
        if s == 1 or not self.tf_same:
            p = (k - 1) // 2
            return F.conv2d(x, self.weight, self.bias, s, p, 1, self.groups)
        pt, pb = tf_same_pad(x.shape[-2], k, s)
        pl, pr = tf_same_pad(x.shape[-1], k, s)
        x = F.pad(x, (pl, pr, pt, pb))
        return F.conv2d(x, self.weight, self.bias, s, 0, 1, self.groups)


class BatchNormAct2d(nn.BatchNorm2d):
    """BatchNorm2d with a fused activation; the affine weight/bias are the FiLM site."""

    def __init__(self, c, eps=1e-3, act=True):
        super().__init__(c, eps=eps, momentum=0.1)
        self.act = nn.SiLU() if act else nn.Identity()

    def forward(self, x):
        return self.act(super().forward(x))


class SqueezeExcite(nn.Module):
    def __init__(self, c, rd):
        super().__init__()
        self.conv_reduce = nn.Conv2d(c, rd, 1, bias=True)
        self.act1 = nn.SiLU()
        self.conv_expand = nn.Conv2d(rd, c, 1, bias=True)

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        s = self.conv_expand(self.act1(self.conv_reduce(s)))
        return x * torch.sigmoid(s)


class DepthwiseSeparableConv(nn.Module):
    """timm block type of stage 0 (expand ratio 1). NOT a FiLM site (film.py:41-44)."""

    def __init__(self, cin, cout, k, s, eps):
        super().__init__()
        self.conv_dw = Conv2dSame(cin, cin, k, s, groups=cin, bias=False)
        self.bn1 = BatchNormAct2d(cin, eps)
        self.se = SqueezeExcite(cin, max(1, cin // 4))
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = BatchNormAct2d(cout, eps, act=False)
        self.has_skip = (s == 1 and cin == cout)

    def forward(self, x):
        y = self.bn1(self.conv_dw(x))
        y = self.se(y)
        y = self.bn2(self.conv_pw(y))
        return y + x if self.has_skip else y


class InvertedResidual(nn.Module):
    """MBConv. ``bn2`` (after the depthwise conv) is the FiLM site (film.py:43-44)."""

    def __init__(self, cin, cout, k, s, expand, eps):
        super().__init__()
        mid = cin * expand
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = BatchNormAct2d(mid, eps)
        self.conv_dw = Conv2dSame(mid, mid, k, s, groups=mid, bias=False)
        self.bn2 = BatchNormAct2d(mid, eps)
        self.se = SqueezeExcite(mid, max(1, cin // 4))
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = BatchNormAct2d(cout, eps, act=False)
        self.has_skip = (s == 1 and cin == cout)

    def forward(self, x):
        y = self.bn1(self.conv_pw(x))
        y = self.bn2(self.conv_dw(y))
        y = self.se(y)
        y = self.bn3(self.conv_pwl(y))
        return y + x if self.has_skip else y


class ConvBnAct(nn.Module):
    """timm 'cn' block (EfficientNet-V2 stage 0): conv3x3 + bn1(SiLU) (+ skip). ``bn1`` is the FiLM site (film.py:41-42)."""

    def __init__(self, cin, cout, k, s, eps):
        super().__init__()
        self.conv = Conv2dSame(cin, cout, k, s, bias=False)
        self.bn1 = BatchNormAct2d(cout, eps)
        self.has_skip = (s == 1 and cin == cout)

    def forward(self, x):
        y = self.bn1(self.conv(x))
        return y + x if self.has_skip else y


class EdgeResidual(nn.Module):
    """timm 'er' block (fused MBConv): conv_exp kxk + bn1(SiLU) -> conv_pwl 1x1 + bn2 (+ skip). FiLM site: ``bn1``."""

    def __init__(self, cin, cout, k, s, expand, eps):
        super().__init__()
        mid = cin * expand
        self.conv_exp = Conv2dSame(cin, mid, k, s, bias=False)
        self.bn1 = BatchNormAct2d(mid, eps)
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn2 = BatchNormAct2d(cout, eps, act=False)
        self.has_skip = (s == 1 and cin == cout)

    def forward(self, x):
        y = self.bn2(self.conv_pwl(self.bn1(self.conv_exp(x))))
        return y + x if self.has_skip else y


class CondConvResidual(nn.Module):   # imported by film.py:36, never instantiated by the supported extractors
    pass


# (repeats, kernel, stride, out channels, expand) per stage of EfficientNet-B0
EFFNET_B0_STAGES = (
    (1, 3, 1, 16, 1),
    (2, 3, 2, 24, 6),
    (2, 5, 2, 40, 6),
    (3, 3, 2, 80, 6),
    (3, 5, 1, 112, 6),
    (4, 5, 2, 192, 6),
    (1, 3, 1, 320, 6),
)


# (block type, repeats, kernel, stride, out channels, expand) per stage of timm tf_efficientnetv2_s
EFFNET_V2S_STAGES = (
    ('cn', 2, 3, 1, 24, 1),
    ('er', 4, 3, 2, 48, 4),
    ('er', 4, 3, 2, 64, 4),
    ('ir', 6, 3, 2, 128, 4),
    ('ir', 9, 3, 1, 160, 6),
    ('ir', 15, 3, 2, 256, 6),
)


class EfficientNet(nn.Module):
    def __init__(self, eps=1e-3, stem=32, head=1280, stages=EFFNET_B0_STAGES):
        super().__init__()
        self.conv_stem = Conv2dSame(3, stem, 3, 2, bias=False)
        self.bn1 = BatchNormAct2d(stem, eps)
        blocks, cin = [], stem
        for spec in stages:
            kind = spec[0] if isinstance(spec[0], str) else None
            (r, k, s, cout, e) = spec[1:] if kind else spec
            stage = []
            for j in range(r):
                st = s if j == 0 else 1
                if kind == 'cn':
                    stage.append(ConvBnAct(cin, cout, k, st, eps))
                elif kind == 'er':
                    stage.append(EdgeResidual(cin, cout, k, st, e, eps))
                elif kind is None and e == 1:
                    stage.append(DepthwiseSeparableConv(cin, cout, k, st, eps))
                else:
                    stage.append(InvertedResidual(cin, cout, k, st, e, eps))
                cin = cout
            blocks.append(nn.Sequential(*stage))
        self.blocks = nn.Sequential(*blocks)
        self.conv_head = nn.Conv2d(cin, head, 1, bias=False)
        self.bn2 = BatchNormAct2d(head, eps)
        self.output_size = head

    def forward(self, x):
        x = self.bn1(self.conv_stem(x))
        x = self.blocks(x)
        x = self.bn2(self.conv_head(x))
        return x.mean((2, 3))


# ----------------------------------------------------------------------------- ViT
class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        h = self.num_heads
        qkv = self.qkv(x).reshape(B, N, 3, h, C // h).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        a = (q @ k.transpose(-2, -1)) * ((C // h) ** -0.5)
        a = a.softmax(dim=-1)
        return self.proj((a @ v).transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads, eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = Mlp(dim, 4 * dim)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, dim, patch):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, patch, patch, bias=True)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    """timm VisionTransformer(patch 32, 224, num_classes=0, global_pool='token')."""

    def __init__(self, dim=768, depth=12, heads=12, patch=32, img=224, eps=1e-6, pre_norm=False):
        super().__init__()
        self.patch_embed = PatchEmbed(dim, patch)
        n = (img // patch) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, dim))
        self.norm_pre = nn.LayerNorm(dim, eps=eps) if pre_norm else nn.Identity()
        self.blocks = nn.Sequential(*[Block(dim, heads, eps) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=eps)
        self.output_size = dim

    def forward(self, x):
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + self.pos_embed
        x = self.blocks(self.norm_pre(x))
        return self.norm(x)[:, 0]


# ----------------------------------------------------------------------------- factories / init
def resnet18():
    import torchvision
    m = torchvision.models.resnet18(weights=None)
    m.fc = nn.Identity()
    m.output_size = 512
    return m


def build(name: str) -> nn.Module:
    """Extractor strings of reference feature_extractors.py:39-66 (+ resnet18 extension)."""
    if name == 'efficientnet_b0':
        return EfficientNet()
    if name == 'efficientnet_v2_s':
        return EfficientNet(stem=24, stages=EFFNET_V2S_STAGES)
    if name == 'vit_s_32':
        return VisionTransformer(384, 12, 6)
    if name == 'vit_b_32':
        return VisionTransformer(768, 12, 12)
    if name == 'vit_b_32_clip':
        return VisionTransformer(768, 12, 12, eps=1e-5, pre_norm=True)
    if name == 'resnet18':
        return resnet18()
    raise ValueError(f"Invalid feature_extractor_name: {name}")


@torch.no_grad()
def seeded_init(model: nn.Module, seed: int = 1991, calib_input: torch.Tensor = None):
    """Synthetic 'pretrained' weights (there is no network for real checkpoints).

    A Gaussian-initialised deep net sits at the edge of chaos: per-frame deviations are amplified
    layer after layer, a few frames end up with |feature| ~ 10 and |logit| ~ 1e4 (where an absolute
    1e-3 logit tolerance is below fp32 resolution), and the slightest damping collapses all frames
    onto one point instead. So the weights are drawn for *dynamical isometry*: (semi-)orthogonal
    dense / pointwise / full convolutions, depthwise kernels = centre tap + small noise, mildly
    random norm affine parameters; then ONE train-mode calibration pass over ``calib_input``
    [n,3,H,W] (momentum 1 => running stats == batch stats). Features stay O(1) for every frame,
    logits O(100), top-2 gaps >> 1e-3. Returns the model in eval mode with frozen parameters."""
    g = torch.Generator().manual_seed(seed)
    for name, mod in model.named_modules():
        if isinstance(mod, nn.Conv2d):
            w = mod.weight
            if mod.groups > 1 and w.shape[1] == 1:                 # depthwise
                k = w.shape[-1]
                w.copy_(torch.randn(w.shape, generator=g) * (0.2 / k))
                w[:, 0, k // 2, k // 2] += 1.0
            elif '.se.' in name:                                   # squeeze-excite FCs: gates around 0.5
                w.copy_(torch.randn(w.shape, generator=g) * (0.5 * w.shape[1] ** -0.5))
            else:
                flat = torch.empty(w.shape[0], w[0].numel())
                nn.init.orthogonal_(flat, generator=g)
                w.copy_(flat.view_as(w))
            if mod.bias is not None:
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.05)
        elif isinstance(mod, nn.Linear):
            nn.init.orthogonal_(mod.weight, generator=g)
            mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.02)
        elif isinstance(mod, (nn.BatchNorm2d, nn.LayerNorm)):
            mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
            mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
    if isinstance(model, VisionTransformer):
        model.cls_token.copy_(0.02 * torch.randn(model.cls_token.shape, generator=g))
        model.pos_embed.copy_(0.02 * torch.randn(model.pos_embed.shape, generator=g))
        model.norm.weight.mul_(0.3)   # token features O(0.3) -> logits O(100) instead of O(1000)
    bns = [m for m in model.modules() if isinstance(m, nn.BatchNorm2d)]
    if bns and calib_input is not None:
        for m in bns:
            m.momentum = 1.0
        model.train()
        model(calib_input)
        for m in bns:
            m.momentum = 0.1
    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model
