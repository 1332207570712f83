"""GPU parity of the EfficientNet-V2-S extractor (reference model/feature_extractors.py:16-19,44-48: timm
tf_efficientnetv2_s_in21k; ConvBnAct / EdgeResidual / InvertedResidual stages, FiLM on ConvBnAct.bn1, EdgeResidual.bn1,
InvertedResidual.bn2 and the root bn1/bn2 per model/film.py:38-46) against the oracle: features, and a CNAPs episode
whose generated FiLM parameters pass through all 84 sites."""
import pytest
import torch

from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode

pytestmark = pytest.mark.gpu


def _pair(cuda_device, adapt, head, size, clip_length=2):
    import orbit_b200
    from orbit_b200.feature_extractors import get_film_parameters
    oracle = OracleRecogniser('efficientnet_v2_s', adapt, head, clip_length, 64, 1.0, 1991, calibration_frames(size))
    m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_v2_s', adapt, head, clip_length, 64, False, 16)
    m.load_state_dict(oracle.state_dict(), strict=True)
    m._set_device(cuda_device)
    m._send_to_device()
    m.set_test_mode(True)
    if adapt:
        m.film_generator.initial_film_parameters = get_film_parameters(m.film_parameter_names, m.feature_extractor)
    return oracle, m


@pytest.mark.parametrize("size,gemm", [(64, 0), (64, 1), (224, 1)])
def test_efficientnet_v2_s_features(cuda_device, size, gemm):
    oracle, m = _pair(cuda_device, False, 'proto', size)
    m.feature_extractor.set_option('gemm', gemm)
    m.feature_extractor.set_option('chunk_frames', 4)
    x = calibration_frames(size)[:6] * 0.9
    with torch.no_grad():
        ref = oracle.extractor(x)
    got = m.feature_extractor(x.to(cuda_device)).cpu()
    err = (got - ref).abs().max().item()
    print(f"efficientnet_v2_s @{size} gemm={gemm}: max|err|={err:.2e} max|ref|={ref.abs().max():.3f}")
    assert got.shape == ref.shape == (6, 1280)
    assert err <= 5e-5 * max(1.0, ref.abs().max().item())


def test_efficientnet_v2_s_cnaps_episode_with_film(cuda_device):
    oracle, m = _pair(cuda_device, True, 'versa', 96)
    assert len(m.film_parameter_names) == 84
    spec = EpisodeSpec(5, 2, 2, 2, 96)
    ctx, ctx_y, tgt, _ = make_episode(spec, index=3)
    oracle.personalise(ctx, ctx_y)
    ref = oracle.predict(tgt)
    m.personalise(ctx.to(cuda_device), ctx_y.to(cuda_device))
    logits, am = m.predict(tgt.to(cuda_device), want_argmax=True)
    err = (logits.cpu() - ref).abs().max().item()
    print(f"CNAPs+efficientnet_v2_s+FiLM: max|dlogit|={err:.2e} max|logit|={ref.abs().max():.1f}")
    assert err <= 1e-3 * max(1.0, ref.abs().max().item() / 100.0)
    assert torch.equal(am.cpu().long(), ref.argmax(1))
