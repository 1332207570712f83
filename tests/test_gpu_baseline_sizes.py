"""The CUDA path against the CPU oracle AT THE SIZES BASELINE.json names, not in miniature:

  S2 (config 2)  ProtoNet + efficientnet_b0, 224 px, 5-way, 200 support + 80 query clips x 8 frames (2,240 frames)
  S3 (config 3)  CNAPs (versa + FiLM) + resnet18, 224 px, 15-way 10-shot (150 support clips), 60 query clips
  S4 (config 4)  FineTuner + vit_b_32, 224 px, 8-way 10-shot, 50 Adam steps lr 1e-3 (utils/args.py:163-178), 40 query clips

The oracle needs 10-40 s of host CPU per case, which is why the other GPU tests use small episodes. The query sets of
S3/S4 are cut to 4-5 clips per class (the support side, the way and the step count are the full configuration)."""
import pytest
import torch

from conftest import assert_logits_match
from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import S2, EpisodeSpec, calibration_frames, make_episode

pytestmark = pytest.mark.gpu


def _load(model, oracle, cuda_device, adapt=False):
    model.load_state_dict(oracle.state_dict(), strict=True)
    model._set_device(cuda_device)
    model._send_to_device()
    model.set_test_mode(True)
    if adapt:
        from orbit_b200.feature_extractors import get_film_parameters
        model.film_generator.initial_film_parameters = get_film_parameters(model.film_parameter_names, model.feature_extractor)
    return model


def test_s2_full_episode_matches_oracle(cuda_device):
    import orbit_b200
    torch.set_num_threads(max(1, torch.get_num_threads()))
    # batch_size only groups clips into passes (eval mode: frames are independent); 8 clips keep the CPU oracle's memory small
    oracle = OracleRecogniser('efficientnet_b0', False, 'proto', S2.clip_length, 8, 1.0, 1991, calibration_frames(224))
    m = _load(orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', S2.clip_length, 256, False, 16),
              oracle, cuda_device)
    ctx, ctx_y, tgt, _ = make_episode(S2, index=0)
    assert ctx.shape == (200, 8, 3, 224, 224) and tgt.shape == (80, 8, 3, 224, 224)
    oracle.personalise(ctx, ctx_y)
    ref = oracle.predict(tgt)
    m.personalise(ctx, ctx_y.to(cuda_device))               # host clips: the path a learner script takes
    logits = m.predict(tgt)
    assert_logits_match(logits, ref, "S2 full episode (2,240 frames)")
    m._reset()
    # `fast` numerics (one fp16 product per GEMM instead of three): report the difference, hold it to 5e-2 of |logit|
    m.feature_extractor.set_option('gemm', 2)
    m.personalise(ctx, ctx_y.to(cuda_device))
    fast = m.predict(tgt).cpu()
    diff = (fast - ref).abs().max().item()
    agree = (fast.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"S2 fast mode: max|dlogit|={diff:.3e} ({diff / ref.abs().max().item():.1e} of max|logit|), arg-max agreement {agree:.3f}")
    assert diff <= 5e-2 * ref.abs().max().item() and agree >= 0.95


def test_s3_cnaps_resnet18_15way_10shot_224(cuda_device):
    import orbit_b200
    oracle = OracleRecogniser('resnet18', True, 'versa', 1, 16, 1.0, 1991, calibration_frames(224))
    m = _load(orbit_b200.SingleStepFewShotRecogniser('resnet18', True, 'versa', 1, 256, False, 16), oracle, cuda_device, adapt=True)
    ctx, ctx_y, tgt, _ = make_episode(EpisodeSpec(15, 10, 4, 1, 224), index=15)
    assert ctx.shape == (150, 1, 3, 224, 224) and len(torch.unique(ctx_y)) == 15
    oracle.personalise(ctx, ctx_y)
    ref = oracle.predict(tgt)
    m.personalise(ctx, ctx_y.to(cuda_device))
    logits = m.predict(tgt)
    assert logits.shape == (60, 15)
    assert_logits_match(logits, ref, "S3 CNAPs + resnet18 + FiLM, 15-way 10-shot @224")


def test_s4_finetuner_vit_b_32_8way_10shot_50_steps(cuda_device):
    import orbit_b200
    oracle = OracleRecogniser('vit_b_32', False, 'linear', 1, 1024)
    m = _load(orbit_b200.MultiStepFewShotRecogniser('vit_b_32', False, 'linear', 1, 1024, False), oracle, cuda_device)
    ctx, ctx_y, tgt, _ = make_episode(EpisodeSpec(8, 10, 5, 1, 224), index=4)
    ctx, ctx_y = ctx[:-1], ctx_y[:-1]    # 79 clips: class counts != N/C (the first bias gradient of a zero-initialised head is
    # exactly 0 for a class holding N/C clips and Adam's g/(|g|+eps) then amplifies rounding noise)
    oracle.personalise_finetune(ctx, ctx_y, num_grad_steps=50, learning_rate=1e-3)
    ref = oracle.predict(tgt)
    m.personalise(ctx, ctx_y, {'num_grad_steps': 50, 'learning_rate': 1e-3, 'optimizer': 'adam', 'loss_fn': None,
                               'extractor_lr_scale': 0.1, 'epsilon': 1e-8, 'weight_decay': 0.0, 'betas': (0.9, 0.999)})
    logits = m.predict(tgt)
    assert logits.shape == (40, 8)
    assert_logits_match(logits, ref, "S4 FineTuner + vit_b_32, 8-way 10-shot, 50 Adam steps")
