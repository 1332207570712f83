"""GPU parity of the native ViT path (timm vit_{small,base}_patch32_224*) against the oracle restatement:
features, a ProtoNets episode, a FiLM (LayerNorm-modulated) episode and the FineTuner configuration (S4 shape, small)."""
import pytest
import torch

from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode

pytestmark = pytest.mark.gpu


def _product(cls, oracle, cuda_device, *args):
    m = cls(*args)
    m.load_state_dict(oracle.state_dict(), strict=True)
    m._set_device(cuda_device)
    m._send_to_device()
    m.set_test_mode(True)
    return m


@pytest.mark.parametrize("name", ["vit_b_32", "vit_s_32", "vit_b_32_clip"])
@pytest.mark.parametrize("gemm", [0, 1])
def test_vit_features_match_oracle(cuda_device, name, gemm):
    import orbit_b200
    oracle = OracleRecogniser(name, False, 'proto', 1, 64)
    m = _product(orbit_b200.SingleStepFewShotRecogniser, oracle, cuda_device, name, False, 'proto', 1, 64, False, 16)
    m.feature_extractor.set_option('gemm', gemm)
    m.feature_extractor.set_option('chunk_frames', 4)
    x = calibration_frames(224)[:7]
    with torch.no_grad():
        ref = oracle.extractor(x)
    got = m.feature_extractor(x.to(cuda_device)).cpu()
    err = (got - ref).abs().max().item()
    print(f"{name} gemm={gemm}: max|err|={err:.2e} max|ref|={ref.abs().max():.3f}")
    assert got.shape == ref.shape
    assert err <= 3e-5 * max(1.0, ref.abs().max().item())


def test_vit_proto_and_film_episodes(cuda_device):
    import orbit_b200
    from orbit_b200.feature_extractors import get_film_parameters
    spec = EpisodeSpec(3, 2, 3, 1, 224)
    ctx, ctx_y, tgt, _ = make_episode(spec, index=1)
    for adapt in (False, True):
        oracle = OracleRecogniser('vit_b_32', adapt, 'proto', 1, 4, calib_input=calibration_frames(224))
        m = _product(orbit_b200.SingleStepFewShotRecogniser, oracle, cuda_device, 'vit_b_32', adapt, 'proto', 1, 4, False, 16)
        if adapt:
            assert len(m.film_parameter_names) == 50      # 25 LayerNorms x (weight, bias), film.py:57-66
            m.film_generator.initial_film_parameters = get_film_parameters(m.film_parameter_names, m.feature_extractor)
        oracle.personalise(ctx, ctx_y)
        ref = oracle.predict(tgt)
        m.personalise(ctx, ctx_y.to(cuda_device))
        logits, am = m.predict(tgt.to(cuda_device), want_argmax=True)
        err = (logits.cpu() - ref).abs().max().item()
        print(f"vit_b_32 adapt_features={adapt}: max|dlogit|={err:.2e} max|logit|={ref.abs().max():.1f}")
        assert err <= 1e-3
        assert torch.equal(am.cpu().long(), ref.argmax(1))


def test_vit_finetuner_config4_shape(cuda_device):
    """BASELINE.json config 4 in miniature: MultiStep + vit_b_32 + linear head, Adam lr 1e-3, 50 steps."""
    import orbit_b200
    spec = EpisodeSpec(4, 3, 2, 1, 224)
    ctx, ctx_y, tgt, _ = make_episode(spec, index=5)
    ctx, ctx_y = ctx[:-1], ctx_y[:-1]      # class counts != N/C (see test_finetune_kernel_matches_torch_optimisers)
    oracle = OracleRecogniser('vit_b_32', False, 'linear', 1, 1024)
    m = _product(orbit_b200.MultiStepFewShotRecogniser, oracle, cuda_device, 'vit_b_32', False, 'linear', 1, 1024, False)
    oracle.personalise_finetune(ctx, ctx_y, num_grad_steps=50, learning_rate=1e-3)
    ref = oracle.predict(tgt)
    m.personalise(ctx, ctx_y, {'num_grad_steps': 50, 'learning_rate': 1e-3, 'optimizer': 'adam', 'loss_fn': None,
                               'extractor_lr_scale': 0.1, 'epsilon': 1e-8, 'weight_decay': 0.0, 'betas': (0.9, 0.999)})
    logits = m.predict(tgt).cpu()
    err = (logits - ref).abs().max().item()
    print(f"FineTuner vit_b_32: max|dlogit|={err:.2e} max|logit|={ref.abs().max():.2f}")
    assert err <= 1e-3 and torch.equal(logits.argmax(1), ref.argmax(1))
