"""The reference learners choose their GPU with `--gpu N` and never call torch.cuda.set_device
(single-step-learner.py:65-66): a model on cuda:1 must work while cuda:0 is the current device. Needs two GPUs
(skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_model_on_non_current_device(oracle_effnet):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import orbit_b200
    from orbit_b200.synthetic import EpisodeSpec, make_episode
    torch.cuda.set_device(0)
    dev = torch.device('cuda:1')
    m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 2, 256, False, 16)
    m.load_state_dict(oracle_effnet.state_dict(), strict=True)
    m._set_device(dev)
    m._send_to_device()
    m.set_test_mode(True)
    ctx, ctx_y, tgt, _ = make_episode(EpisodeSpec(5, 2, 2, 2, 96), index=0)
    oracle_effnet.reset()
    oracle_effnet.personalise(ctx, ctx_y)
    ref = oracle_effnet.predict(tgt)
    for c, t in ((ctx.to(dev), tgt.to(dev)), (ctx, tgt)):          # device-resident and host clips
        m.personalise(c, ctx_y.to(dev))
        logits = m.predict(t)
        assert logits.device == dev and torch.cuda.current_device() == 0
        assert (logits.cpu() - ref).abs().max().item() <= 1e-3
        assert torch.equal(logits.argmax(1).cpu(), ref.argmax(1))
        m._reset()
