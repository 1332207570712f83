"""Full-size (BASELINE.json configs[1], "S2": 200 support + 80 query clips x 8 frames of 224x224 = 2,240 frames per episode)
checks of the CUDA path through size-independent properties, where the CPU oracle would take minutes:
determinism, equivariance under a permutation of the query clips, host- vs device-resident inputs, chunk-size
independence, and the prototype identity  logits == 2 q.mu - mu.mu  (reference classifier_heads.py:94-105,202-263)
evaluated in fp64 from the extractor's own per-frame features."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def s2_model(cuda_device):
    import orbit_b200
    from orbit_b200.synthetic import S2, load_synthetic_checkpoint
    m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', S2.clip_length, 256, False, 16)
    m._set_device(cuda_device)
    m._send_to_device()
    m.set_test_mode(True)
    load_synthetic_checkpoint(m, S2.frame_size)
    return m


@pytest.fixture(scope="module")
def s2_episode():
    from orbit_b200.synthetic import S2, make_episode
    return make_episode(S2, index=11, pin=True)


def _episode_logits(m, ctx, ctx_y, tgt):
    m.personalise(ctx, ctx_y)
    logits, am = m.predict(tgt, want_argmax=True)
    m._reset()
    return logits.clone(), am.clone()


def test_s2_determinism_permutation_and_host_path(cuda_device, s2_model, s2_episode):
    ctx, ctx_y, tgt, tgt_y = s2_episode
    m = s2_model
    cd, cyd, td = ctx.to(cuda_device), ctx_y.to(cuda_device), tgt.to(cuda_device)
    a, am = _episode_logits(m, cd, cyd, td)
    assert a.shape == (80, 5) and torch.isfinite(a).all()
    b, _ = _episode_logits(m, cd, cyd, td)
    assert torch.equal(a, b), "same episode twice must give bit-identical logits"
    # query clips are scored independently: permuting them permutes the rows, bit for bit
    perm = torch.randperm(len(tgt), generator=torch.Generator().manual_seed(5)).to(cuda_device)
    c, cm = _episode_logits(m, cd, cyd, td[perm])
    assert torch.equal(c, a[perm]) and torch.equal(cm, am[perm])
    # support order does not matter beyond fp32 summation order of the class means
    sperm = torch.randperm(len(ctx), generator=torch.Generator().manual_seed(6)).to(cuda_device)
    d, dm = _episode_logits(m, cd[sperm], cyd[sperm], td)
    assert (d - a).abs().max().item() <= 1e-3 and torch.equal(dm, am)
    # pinned host clips (ramped, overlapped H2D passes) == device-resident clips, bit for bit
    e, em = _episode_logits(m, ctx, cyd, tgt)
    assert torch.equal(e, a) and torch.equal(em, am)
    assert m._stager.bytes_copied >= (ctx.numel() + tgt.numel()) * 4
    # the synthetic task is learnable: the prototypes separate the five objects
    assert (am.cpu().long() == tgt_y).float().mean().item() > 0.5


def test_s2_chunk_size_is_invisible(cuda_device, s2_model, s2_episode):
    _, _, tgt, _ = s2_episode
    fe = s2_model.feature_extractor
    frames = tgt.to(cuda_device).flatten(end_dim=1)[:400]
    old = fe.get_option('chunk_frames')
    try:
        fe.set_option('chunk_frames', 400)
        a = fe(frames).clone()
        fe.set_option('chunk_frames', 96)
        b = fe(frames).clone()
    finally:
        fe.set_option('chunk_frames', old)
    assert torch.equal(a, b)


def test_s2_prototype_identity_fp64(cuda_device, s2_model, s2_episode):
    ctx, ctx_y, tgt, _ = s2_episode
    m = s2_model
    L_ = m.clip_length
    cd, td = ctx.to(cuda_device), tgt.to(cuda_device)
    logits, _ = _episode_logits(m, cd, ctx_y.to(cuda_device), td)
    fs = m.feature_extractor(cd.flatten(end_dim=1)).double().cpu()
    fq = m.feature_extractor(td.flatten(end_dim=1)).double().cpu()
    ps, pq = fs.view(-1, L_, fs.shape[1]).mean(1), fq.view(-1, L_, fq.shape[1]).mean(1)
    mu = torch.stack([ps[ctx_y == c].mean(0) for c in torch.unique(ctx_y)])          # sorted unique labels
    ref = 2.0 * pq @ mu.t() - (mu * mu).sum(1)
    err = (logits.double().cpu() - ref).abs().max().item()
    print(f"S2 head identity: max|dlogit|={err:.3e} at max|logit|={ref.abs().max().item():.1f}")
    assert err <= 1e-3
    assert torch.equal(logits.argmax(1).cpu(), ref.argmax(1))
