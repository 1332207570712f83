"""OpsCounter through real episodes on the device (reference call sites: single-step-learner.py:322,
multi-step-learner.py:169): the counted MACs equal the reference's accounting evaluated by hand for the same shapes."""
import pytest
import torch

from orbit_b200.synthetic import EpisodeSpec, make_episode

pytestmark = pytest.mark.gpu


def _model(cls, cuda_device, *args):
    m = cls(*args)
    m._set_device(cuda_device)
    m._send_to_device()
    m.set_test_mode(True)
    return m


def test_protonet_personalise_macs(cuda_device):
    import orbit_b200
    m = _model(orbit_b200.SingleStepFewShotRecogniser, cuda_device, 'efficientnet_b0', False, 'proto', 2, 4, False, 16)
    oc = orbit_b200.OpsCounter()
    oc.set_base_params(m)
    assert oc.base_params_counter == 4007548 and 'feature extractor: 4.01M' in oc.params_break_down
    spec = EpisodeSpec(3, 2, 1, 2, 64)
    ctx, ctx_y, tgt, _ = make_episode(spec, index=1)
    m.personalise(ctx.to(cuda_device), ctx_y.to(cuda_device), ops_counter=oc)
    n, L, d, c = 6, 2, 1280, 3
    per_frame = m.feature_extractor.count_macs(torch.empty(1, 3, 64, 64))
    want = n * L * per_frame + n * L * d + (c * n + n * d) + 3 * c * d
    assert oc.get_task_macs() == want
    assert oc.get_task_params() == 4007548 + 2 * 4007548        # two traced batches of 4 clips (thop adds params per trace)
    logits = m.predict(tgt.to(cuda_device))
    assert logits.shape == (3, 3)


def test_cnaps_and_finetuner_count_every_stage(cuda_device):
    import orbit_b200
    spec = EpisodeSpec(3, 2, 1, 2, 64)
    ctx, ctx_y, tgt, _ = make_episode(spec, index=2)
    m = _model(orbit_b200.SingleStepFewShotRecogniser, cuda_device, 'efficientnet_b0', True, 'versa', 2, 8, False, 16)
    oc = orbit_b200.OpsCounter()
    m.personalise(ctx.to(cuda_device), ctx_y.to(cuda_device), ops_counter=oc)
    n, L, d, c = 6, 2, 1280, 3
    ext = n * L * m.feature_extractor.count_macs(torch.empty(1, 3, 64, 64))
    enc = n * L * m.set_encoder.count_macs(torch.empty(1, 3, 64, 64))
    gen = m.film_generator.count_macs(None)
    versa = c * (3 * d * d + d + 2)
    assert gen > 0 and enc > 0
    assert oc.get_task_macs() == ext + enc + gen + n * L * d + (c * n + n * d) + versa

    ft = _model(orbit_b200.MultiStepFewShotRecogniser, cuda_device, 'efficientnet_b0', False, 'linear', 2, 8, False, 16)
    oc2 = orbit_b200.OpsCounter()
    args = {'num_grad_steps': 3, 'learning_rate': 0.1, 'optimizer': 'sgd', 'loss_fn': None, 'extractor_lr_scale': 0.1}
    ft.personalise(ctx.to(cuda_device), ctx_y.to(cuda_device), dict(args), ops_counter=oc2)
    # default accounting = the reference's: extractor + pooling re-counted in each of the 3 grad steps
    # (few_shot_recognisers.py:231-246), although the frozen features are computed once here
    assert ft.mac_accounting == 'reference'
    assert oc2.get_task_macs() == 3 * (ext + n * L * d) + 3 * c * n * d
    ft._reset()
    ft.mac_accounting = 'actual'
    oc3 = orbit_b200.OpsCounter()
    ft.personalise(ctx.to(cuda_device), ctx_y.to(cuda_device), dict(args), ops_counter=oc3)
    assert oc3.get_task_macs() == ext + n * L * d + 3 * c * n * d
    before = oc2.get_task_macs()
    ft.predict(tgt.to(cuda_device), ops_counter=oc2)
    nq = 3
    assert oc2.get_task_macs() == before + nq * L * m.feature_extractor.count_macs(torch.empty(1, 3, 64, 64)) + nq * L * d + c * nq * d
