"""GPU parity of the squeeze-excite gate (C ABI orbit_se_gate) against a float64 torch restatement of timm's SqueezeExcite
(mean -> conv_reduce -> SiLU -> conv_expand -> sigmoid), on the EfficientNet-B0 / V2-S gate shapes, frame counts that leave a
ragged last block, multi-chunk weight streams, and the shapes that take the plain kernel (C % 4 != 0, switch off)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [  # (B, C, R, groups, hw)
    (5, 32, 8, 4, 112 * 112), (3, 96, 4, 4, 56 * 56), (7, 144, 6, 2, 56 * 56), (9, 240, 10, 1, 28 * 28),
    (13, 480, 20, 2, 14 * 14), (6, 672, 28, 2, 14 * 14), (11, 1152, 48, 2, 49), (1, 1152, 48, 1, 49),
    (160, 1152, 48, 2, 49), (640, 672, 28, 2, 196), (1600, 1152, 48, 1, 49), (1601, 480, 20, 2, 196), (2000, 240, 10, 3, 784),
    (4, 1536, 64, 1, 49), (3, 2048, 128, 1, 9),       # EfficientNetV2-S sized gates: many chunks per matrix
    (5, 30, 5, 3, 16), (4, 4100, 8, 1, 4),            # C % 4 != 0 / C beyond the ring kernel's thread mapping: plain kernel
]


def _run(cuda_device, B, C, R, groups, hw, seed=0):
    from orbit_b200 import lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(1000 * C + R + B + seed)
    partial = torch.randn(B, groups, C, generator=g) * hw / groups
    w1 = torch.randn(R, C, generator=g) / C ** 0.5
    b1 = 0.1 * torch.randn(R, generator=g)
    w2 = torch.randn(C, R, generator=g) / R ** 0.5
    b2 = 0.1 * torch.randn(C, generator=g)
    mean = partial.double().sum(1) / hw
    hid = mean @ w1.double().t() + b1.double()
    hid = hid * torch.sigmoid(hid)
    ref = torch.sigmoid(hid @ w2.double().t() + b2.double())
    d = [t.to(cuda_device) for t in (partial, w1, b1, w2.t().contiguous(), b2)]
    gate = torch.full((B, C), float('nan'), device=cuda_device)
    L.check(lib.orbit_se_gate(L.ptr(d[0]), groups, hw, L.ptr(d[1]), L.ptr(d[2]), L.ptr(d[3]), L.ptr(d[4]), L.ptr(gate), B, C, R,
                              L.stream_ptr(cuda_device)), "orbit_se_gate")
    torch.cuda.synchronize()
    return gate.cpu(), ref


@pytest.mark.parametrize("B,C,R,groups,hw", CASES)
def test_se_gate_matches_torch(cuda_device, B, C, R, groups, hw):
    got, ref = _run(cuda_device, B, C, R, groups, hw)
    assert torch.isfinite(got).all()
    assert (got.double() - ref).abs().max().item() <= 2e-6       # a gate lies in (0, 1): absolute tolerance, fp32 sums of <= 2048 terms


def test_se_gate_plain_and_ring_kernels_agree(cuda_device):
    """the ring-streamed kernel is the default (CASES above); with the switch off the plain kernel serves the same call"""
    from orbit_b200 import lib as L
    lib = L.load()
    ring, ref = _run(cuda_device, 37, 672, 28, 2, 196)
    assert lib.orbit_set_global_option(b'se_ring', 0) == 0
    try:
        plain, _ = _run(cuda_device, 37, 672, 28, 2, 196)
    finally:
        assert lib.orbit_set_global_option(b'se_ring', 1) == 0
    assert (plain.double() - ref).abs().max().item() <= 2e-6
    assert (plain - ring).abs().max().item() <= 2e-6


def test_se_gate_refuses_bad_arguments(cuda_device):
    from orbit_b200 import lib as L
    lib = L.load()
    t = torch.zeros(64, device=cuda_device)
    assert lib.orbit_se_gate(None, 1, 1, L.ptr(t), L.ptr(t), L.ptr(t), L.ptr(t), L.ptr(t), 1, 4, 4, L.stream_ptr(cuda_device)) != 0
    assert lib.orbit_se_gate(L.ptr(t), 0, 1, L.ptr(t), L.ptr(t), L.ptr(t), L.ptr(t), L.ptr(t), 1, 4, 4, L.stream_ptr(cuda_device)) != 0
