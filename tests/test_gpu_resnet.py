"""GPU parity of the ResNet-18 extension (BASELINE.json configs 1 and 3; torchvision resnet18, fc = Identity) against
the oracle: features at 84/224 px, the config-1 episode (ProtoNets, 84x84, 5-way 5-shot, 1-clip) and a config-3 style
CNAPs episode (Versa head + FiLM on the BasicBlock BatchNorms, variable way)."""
import pytest
import torch

from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode

pytestmark = pytest.mark.gpu


def _pair(cuda_device, adapt, head, size, clip_length=1):
    import orbit_b200
    from orbit_b200.feature_extractors import get_film_parameters
    oracle = OracleRecogniser('resnet18', adapt, head, clip_length, 64, 1.0, 1991, calibration_frames(size))
    m = orbit_b200.SingleStepFewShotRecogniser('resnet18', adapt, head, clip_length, 64, False, 16)
    m.load_state_dict(oracle.state_dict(), strict=True)
    m._set_device(cuda_device)
    m._send_to_device()
    m.set_test_mode(True)
    if adapt:
        m.film_generator.initial_film_parameters = get_film_parameters(m.film_parameter_names, m.feature_extractor)
    return oracle, m


@pytest.mark.parametrize("size,gemm", [(84, 0), (84, 1), (224, 1)])
def test_resnet18_features(cuda_device, size, gemm):
    oracle, m = _pair(cuda_device, False, 'proto', size)
    m.feature_extractor.set_option('gemm', gemm)
    m.feature_extractor.set_option('chunk_frames', 4)
    x = calibration_frames(size)[:6] * 0.9
    with torch.no_grad():
        ref = oracle.extractor(x)
    got = m.feature_extractor(x.to(cuda_device)).cpu()
    err = (got - ref).abs().max().item()
    print(f"resnet18 @{size} gemm={gemm}: max|err|={err:.2e} max|ref|={ref.abs().max():.3f}")
    assert got.shape == ref.shape == (6, 512)
    assert err <= 3e-5 * max(1.0, ref.abs().max().item())


def test_config1_episode_protonet_resnet18_84px(cuda_device):
    oracle, m = _pair(cuda_device, False, 'proto', 84)
    spec = EpisodeSpec(5, 5, 15, 1, 84)               # S1: support [25,1,3,84,84], query [75,1,3,84,84]
    ctx, ctx_y, tgt, _ = make_episode(spec, index=0)
    oracle.personalise(ctx, ctx_y)
    ref = oracle.predict(tgt)
    m.personalise(ctx, ctx_y.to(cuda_device))
    logits, am = m.predict(tgt, want_argmax=True)
    err = (logits.cpu() - ref).abs().max().item()
    print(f"config 1 (ProtoNet+resnet18, 84px): max|dlogit|={err:.2e} max|logit|={ref.abs().max():.1f}")
    assert err <= 1e-3 * max(1.0, ref.abs().max().item() / 100.0)
    assert torch.equal(am.cpu().long(), ref.argmax(1))


@pytest.mark.parametrize("way", [5, 9])
def test_config3_episode_cnaps_resnet18_film(cuda_device, way):
    oracle, m = _pair(cuda_device, True, 'versa', 96)
    assert len(m.film_parameter_names) == 32       # bn1/bn2 of 8 BasicBlocks, weight + bias
    spec = EpisodeSpec(way, 2, 2, 1, 96)
    ctx, ctx_y, tgt, _ = make_episode(spec, index=way)
    oracle.personalise(ctx, ctx_y)
    ref = oracle.predict(tgt)
    m.personalise(ctx.to(cuda_device), ctx_y.to(cuda_device))
    logits, am = m.predict(tgt.to(cuda_device), want_argmax=True)
    err = (logits.cpu() - ref).abs().max().item()
    print(f"config 3 (CNAPs+resnet18+FiLM, {way}-way): max|dlogit|={err:.2e} max|logit|={ref.abs().max():.1f}")
    assert logits.shape == (2 * way, way)
    assert err <= 1e-3 * max(1.0, ref.abs().max().item() / 100.0)
    assert torch.equal(am.cpu().long(), ref.argmax(1))
