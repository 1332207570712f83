"""GPU parity of the fused head kernels (through the C ABI) against the oracle restatement of
reference model/classifier_heads.py + model/poolers.py. Tolerances: fp32 reductions in a different
order than the CPU => a few ulp of the accumulated magnitude; arg-max must be identical."""
import pytest
import torch

from oracle import parts

pytestmark = pytest.mark.gpu


def _case(seed, ns, nq, L, D, C, labels_offset=0, stride=1):
    g = torch.Generator().manual_seed(seed)
    sf = torch.randn(ns * L, D, generator=g) * 0.7 + 0.3
    qf = torch.randn(nq * L, D, generator=g) * 0.7 + 0.3
    labels = (torch.arange(ns) % C) * stride + labels_offset
    labels = labels[torch.randperm(ns, generator=g)]
    return sf, qf, labels


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
@pytest.mark.parametrize("ns,nq,L,D,C", [(200, 80, 8, 1280, 5), (25, 75, 1, 512, 5), (150, 300, 1, 768, 15),
                                         (7, 3, 3, 1280, 3), (40, 33, 2, 384, 20), (64, 10, 1, 1280, 64)])
def test_proto_head_matches_oracle(cuda_device, metric, ns, nq, L, D, C):
    from orbit_b200 import PrototypicalClassifier
    sf, qf, labels = _case(ns * 7 + D, ns, nq, L, D, C)
    w_ref, b_ref = parts.proto_configure(parts.pool_clips(sf, L), labels, metric)
    ref = parts.proto_predict(parts.pool_clips(qf, L), w_ref, b_ref, 1.7, metric)

    head = PrototypicalClassifier(1.7, metric)
    head.configure(sf.to(cuda_device), labels.to(cuda_device), clip_length=L)
    logits, am = head.predict(qf.to(cuda_device), clip_length=L, want_argmax=True)
    torch.cuda.synchronize()
    assert head.weight.shape == w_ref.shape
    scale = max(1.0, ref.abs().max().item())
    assert (head.weight.cpu() - w_ref).abs().max().item() <= 1e-5
    if metric == 'euclidean':
        assert (head.bias.cpu() - b_ref).abs().max().item() <= 1e-5 * max(1.0, b_ref.abs().max().item())
    assert (logits.cpu() - ref).abs().max().item() <= 2e-6 * scale + 1e-6
    # arg-max: bit-exact class indices (vs the oracle's own arg-max and vs arg-max of our logits)
    assert torch.equal(am.cpu().long(), logits.argmax(dim=1).cpu())
    gap = ref.topk(2, dim=1).values
    decided = (gap[:, 0] - gap[:, 1]) > 1e-4 * scale
    assert torch.equal(am.cpu().long()[decided], ref.argmax(dim=1)[decided])


def test_noncontiguous_labels_and_reset(cuda_device):
    """arbitrary label values -> columns in sorted-label order; reset() clears state and
    predict() before personalise raises AttributeError (classifier_heads.py:197-211)."""
    from orbit_b200 import PrototypicalClassifier
    sf, qf, labels = _case(3, 30, 12, 1, 256, 5, labels_offset=100, stride=7)
    w_ref, b_ref = parts.proto_configure(sf, labels)
    head = PrototypicalClassifier()
    with pytest.raises(AttributeError):
        head.predict(qf.to(cuda_device))
    head.configure(sf.to(cuda_device), labels.to(cuda_device))
    assert torch.equal(head.classes, torch.unique(labels))
    assert (head.predict(qf.to(cuda_device)).cpu() - parts.proto_predict(qf, w_ref, b_ref)).abs().max() < 1e-3
    head.reset()
    assert head.weight is None and head.bias is None
    with pytest.raises(AttributeError):
        head.predict(qf.to(cuda_device))


def test_configure_is_deterministic(cuda_device):
    from orbit_b200 import PrototypicalClassifier
    sf, qf, labels = _case(5, 200, 80, 8, 1280, 5)
    outs = []
    for _ in range(3):
        head = PrototypicalClassifier()
        head.configure(sf.to(cuda_device), labels.to(cuda_device), clip_length=8)
        outs.append((head.weight.detach().clone(), head.bias.detach().clone(),
                     head.predict(qf.to(cuda_device), clip_length=8)))
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0]))


def test_mean_pooler(cuda_device):
    from orbit_b200 import MeanPooler
    x = torch.randn(24 * 8, 1280)
    out = MeanPooler(8)(x.to(cuda_device)).cpu()
    assert (out - parts.pool_clips(x, 8)).abs().max() < 1e-6


def test_cpu_tensor_is_refused():
    """No CPU fallback: a CPU tensor must fail loudly."""
    from orbit_b200 import MeanPooler, OrbitError
    with pytest.raises(OrbitError):
        MeanPooler(2)(torch.randn(4, 8))


def test_edge_cases_empty_query_single_class_max_classes(cuda_device):
    """Empty query set -> [0, C] logits; one class; the 64-class maximum; 65 classes refused (ValueError)."""
    from orbit_b200 import PrototypicalClassifier
    g = torch.Generator().manual_seed(9)
    sf = torch.randn(12, 64, generator=g)
    head = PrototypicalClassifier()
    head.configure(sf.to(cuda_device), torch.zeros(12, dtype=torch.long))
    assert head.weight.shape == (1, 64)
    out = head.predict(torch.empty(0, 64, device=cuda_device))
    assert out.shape == (0, 1)
    lg, am = head.predict(sf.to(cuda_device), want_argmax=True)
    assert lg.shape == (12, 1) and int(am.max()) == 0
    many = torch.randn(130, 64, generator=g)
    labels = torch.arange(130) % 64
    head.configure(many.to(cuda_device), labels.to(cuda_device))
    w_ref, b_ref = parts.proto_configure(many, labels)
    assert (head.weight.cpu() - w_ref).abs().max() <= 1e-5 and head.weight.shape[0] == 64
    with pytest.raises(ValueError):
        head.configure(many.to(cuda_device), (torch.arange(130) % 65).to(cuda_device))
    with pytest.raises(AssertionError):
        head.configure(many.to(cuda_device), labels[:100].to(cuda_device))   # classifier_heads.py:240
