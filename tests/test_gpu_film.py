"""GPU parity of the CNAPs path (adapt_features=True): set encoder, FiLM generator, FiLM-modulated extractor,
Versa head -- against the oracle restatements of model/set_encoders.py, feature_adapters.py, classifier_heads.py."""
import pytest
import torch

from oracle import parts
from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode

pytestmark = pytest.mark.gpu


def _pair(cuda_device, head, size, clip_length=2):
    import orbit_b200
    from orbit_b200.feature_extractors import get_film_parameters
    oracle = OracleRecogniser('efficientnet_b0', True, head, clip_length, 4, 1.0, 1991, calibration_frames(size))
    m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', True, head, clip_length, 4, False, 16)
    m.load_state_dict(oracle.state_dict(), strict=True)
    m._set_device(cuda_device)
    m._send_to_device()
    m.set_test_mode(True)
    # gamma0/beta0 snapshot = the (synthetic) pretrained norm weights, as at reference construction time
    m.film_generator.initial_film_parameters = get_film_parameters(m.film_parameter_names, m.feature_extractor)
    return oracle, m


@pytest.mark.parametrize("size", [84, 224])
def test_set_encoder_matches_oracle(cuda_device, size):
    oracle, m = _pair(cuda_device, 'versa', size)
    x = calibration_frames(size)[:6] + 0.1
    ref = parts.set_encoder_forward(x, oracle.set_encoder_params)
    got = m.set_encoder(x.to(cuda_device)).cpu()
    err = (got - ref).abs().max().item()
    print(f"set encoder @{size}: max|err|={err:.2e} max|ref|={ref.abs().max():.3f}")
    assert got.shape == ref.shape == (6, 64)
    assert err <= 2e-5 * max(1.0, ref.abs().max().item())
    agg = m.set_encoder.aggregate([got[:2].to(cuda_device), got[2:].to(cuda_device)]).cpu()
    assert (agg - parts.task_embedding(ref)).abs().max() <= 2e-5


def test_film_generator_matches_oracle(cuda_device):
    oracle, m = _pair(cuda_device, 'proto', 64)
    z = torch.randn(1, 64, generator=torch.Generator().manual_seed(1))
    ref, l2 = parts.film_generate(z, oracle.film_names, oracle.film_gen_params, oracle.film_initial)
    got = m.film_generator(z.to(cuda_device))
    assert sorted(got) == sorted(ref) and len(got) == 34
    worst = max((got[k].cpu() - ref[k]).abs().max().item() for k in ref)
    print(f"film generator: max|err|={worst:.2e}; l2={float(m.film_generator.regularization_term()):.4f}")
    assert worst <= 2e-6
    assert abs(float(m.film_generator.regularization_term()) - float(l2)) <= 1e-4 * float(l2)
    # FiLM must actually modulate (the synthetic regularisers are O(0.05), not the reference's 1e-3 init)
    assert (ref['bn1.weight'] - oracle.film_initial['bn1.weight']).abs().max() > 1e-3


@pytest.mark.parametrize("head,way,size", [('versa', 5, 96), ('proto', 3, 84)])
def test_cnaps_episode_matches_oracle(cuda_device, head, way, size):
    oracle, m = _pair(cuda_device, head, size)
    spec = EpisodeSpec(way, 2, 3, 2, size)
    ctx, ctx_y, tgt, _ = make_episode(spec, index=1)
    oracle.personalise(ctx, ctx_y)
    ref = oracle.predict(tgt)
    for on_device in (True, False):
        c, t = (ctx.to(cuda_device), tgt.to(cuda_device)) if on_device else (ctx, tgt)
        m.personalise(c, ctx_y.to(cuda_device))
        assert (m.film_dict['bn1.weight'].cpu() - oracle.film_dict['bn1.weight']).abs().max() <= 1e-5
        logits, am = m.predict(t, want_argmax=True)
        err = (logits.cpu() - ref).abs().max().item()
        print(f"{head} + FiLM, clips on device={on_device}: max|dlogit|={err:.2e} max|logit|={ref.abs().max():.1f}")
        assert err <= 1e-3 * max(1.0, ref.abs().max().item() / 100.0)
        assert torch.equal(am.cpu().long(), ref.argmax(1))
        m._reset()
        assert m.film_dict is None and m.classifier.weight is None


@pytest.mark.parametrize("d,counts", [(128, [6, 5, 4, 3]), (512, [3, 1, 4]), (1280, [4, 4, 3])])
def test_mahalanobis_head_matches_oracle(cuda_device, d, counts):
    """Simple-CNAPs head on given features (incl. a one-clip class: the reference's scalar covariance branch)."""
    from orbit_b200.classifier_heads_ext import MahalanobisClassifier
    g = torch.Generator().manual_seed(d)
    labels = torch.cat([torch.full((n,), 3 * c + 1) for c, n in enumerate(counts)])
    labels = labels[torch.randperm(len(labels), generator=g)]
    feats = torch.randn(len(labels), d, generator=g) * 0.5 + 0.3 * torch.randn(1, d, generator=g)
    q = torch.randn(9, d, generator=g) * 0.5
    means, precs = parts.mahalanobis_configure(feats, labels)
    ref = parts.mahalanobis_predict(q, means, precs, 2.0)
    head = MahalanobisClassifier(2.0)
    head.configure(feats.to(cuda_device), labels.to(cuda_device))
    logits, am = head.predict(q.to(cuda_device), want_argmax=True)
    perr = (head.precisions.detach().cpu() - precs).abs().max().item()
    err = (logits.cpu() - ref).abs().max().item()
    print(f"mahalanobis D={d}: max|dP|={perr:.2e} max|dlogit|={err:.2e} max|logit|={ref.abs().max():.2f}")
    assert (head.means.detach().cpu() - means).abs().max() <= 1e-5
    assert perr <= 1e-4
    assert err <= 2e-4 * max(1.0, ref.abs().max().item())
    assert torch.equal(am.cpu().long(), ref.argmax(1))
    head.reset()
    with pytest.raises(AttributeError):
        head.predict(q.to(cuda_device))


def test_simple_cnaps_episode_matches_oracle(cuda_device):
    oracle, m = _pair(cuda_device, 'mahalanobis', 64, clip_length=1)
    spec = EpisodeSpec(3, 3, 2, 1, 64)
    ctx, ctx_y, tgt, _ = make_episode(spec, index=1)
    oracle.personalise(ctx, ctx_y)
    ref = oracle.predict(tgt)
    m.personalise(ctx.to(cuda_device), ctx_y.to(cuda_device))
    logits = m.predict(tgt.to(cuda_device))
    err = (logits.cpu() - ref).abs().max().item()
    print(f"SimpleCNAPs episode: max|dlogit|={err:.2e} max|logit|={ref.abs().max():.2f}")
    assert err <= 1e-3 * max(1.0, ref.abs().max().item())
    assert torch.equal(logits.argmax(1).cpu(), ref.argmax(1))
