"""GPU parity of the native EfficientNet-B0 forward (through the C ABI) against the oracle
restatement of timm tf_efficientnet_b0 (oracle/backbones.py), same seeded weights."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _product(oracle, cuda_device, clip_length=2):
    import orbit_b200
    m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', clip_length, 256, False, 16)
    m.load_state_dict(oracle.state_dict(), strict=True)
    m._set_device(cuda_device)
    m._send_to_device()
    m.set_test_mode(True)
    return m


@pytest.mark.parametrize("size,n", [(224, 5), (84, 7), (64, 3), (96, 33)])
@pytest.mark.parametrize("gemm", [0, 1])
def test_efficientnet_features_match_oracle(cuda_device, oracle_effnet, size, n, gemm):
    m = _product(oracle_effnet, cuda_device)
    m.feature_extractor.set_option('gemm', gemm)
    m.feature_extractor.set_option('chunk_frames', 4)   # exercises the multi-chunk path (n > chunk)
    x = torch.randn(n, 3, size, size, generator=torch.Generator().manual_seed(size + n))
    with torch.no_grad():
        ref = oracle_effnet.extractor(x)
    out = m.feature_extractor(x.to(cuda_device)).cpu()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"size={size} n={n} max|err|={err:.3e} max|ref|={scale:.3f}")
    assert out.shape == ref.shape
    assert err <= (2e-5 if gemm == 0 else 5e-5) * max(1.0, scale)


def test_chunking_is_invisible(cuda_device, oracle_effnet):
    """Results must not depend on chunk_frames (eval-mode BN: frames are independent)."""
    m = _product(oracle_effnet, cuda_device)
    x = torch.randn(9, 3, 96, 96, generator=torch.Generator().manual_seed(1)).to(cuda_device)
    m.feature_extractor.set_option('chunk_frames', 16)
    a = m.feature_extractor(x)
    m.feature_extractor.set_option('chunk_frames', 2)
    b = m.feature_extractor(x)
    assert torch.equal(a, b)


def test_episode_logits_and_argmax(cuda_device, oracle_effnet):
    """personalise()+predict() on a synthetic episode: logits within 1e-3 (north-star tolerance, fp32)
    of the oracle and identical class indices; CPU-resident clips exercise the staged H2D path."""
    from orbit_b200.synthetic import EpisodeSpec, make_episode
    spec = EpisodeSpec(way=5, support_clips_per_class=3, query_clips_per_class=4, clip_length=2, frame_size=96)
    ctx, ctx_y, tgt, tgt_y = make_episode(spec, index=3)
    oracle_effnet.reset()
    oracle_effnet.personalise(ctx, ctx_y)
    ref = oracle_effnet.predict(tgt)

    m = _product(oracle_effnet, cuda_device)
    m.stage_copy_frames, m.stage_ramp = 3, (4, 6)   # several copies per pass, several passes per call
    for clips_dev, gemm in ((False, 0), (True, 0), (True, 1), (False, 1)):
        m.feature_extractor.set_option('gemm', gemm)
        c, t = (ctx.to(cuda_device), tgt.to(cuda_device)) if clips_dev else (ctx, tgt)
        m.personalise(c, ctx_y.to(cuda_device))
        logits, am = m.predict(t, want_argmax=True)
        err = (logits.cpu() - ref).abs().max().item()
        print(f"device_clips={clips_dev} gemm={gemm} max|dlogit|={err:.3e} max|logit|={ref.abs().max().item():.2f}")
        assert err <= 1e-3
        assert torch.equal(am.cpu().long(), ref.argmax(dim=1))
        m._reset()
        assert m.classifier.weight is None


def test_device_calibration_matches_oracle(cuda_device):
    """orbit_engine_calibrate == one train-mode pass of the oracle with momentum 1 (batch statistics written
    to running_mean/var); afterwards both sides hold the same checkpoint and give the same features."""
    import orbit_b200
    from oracle import backbones
    from orbit_b200.synthetic import calibration_frames, load_synthetic_checkpoint
    m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 1, 256, False, 16)
    m._set_device(cuda_device)
    m._send_to_device()
    load_synthetic_checkpoint(m, 96, seed=5)
    sd = {k[len('feature_extractor.'):]: v.cpu() for k, v in m.state_dict().items() if k.startswith('feature_extractor.')}
    ref = backbones.build('efficientnet_b0')
    # same weights, but let the ORACLE compute the batch statistics itself
    ref.load_state_dict(sd, strict=True)
    bns = [b for b in ref.modules() if isinstance(b, torch.nn.BatchNorm2d)]
    for b in bns:
        b.reset_running_stats()
        b.momentum = 1.0
    ref.train()
    with torch.no_grad():
        ref(calibration_frames(96))
    ref.eval()
    for (k, b) in [(k, b) for k, b in ref.named_modules() if isinstance(b, torch.nn.BatchNorm2d)]:
        assert (b.running_mean - sd[k + '.running_mean']).abs().max() <= 1e-3 * (1 + sd[k + '.running_mean'].abs().max()), k
        assert ((b.running_var - sd[k + '.running_var']).abs() / (sd[k + '.running_var'] + 1e-3)).max() <= 5e-3, k
    x = torch.randn(4, 3, 96, 96, generator=torch.Generator().manual_seed(0))
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    with torch.no_grad():
        want = ref(x)
    got = m.feature_extractor(x.to(cuda_device)).cpu()
    assert (got - want).abs().max() <= 2e-5 * max(1.0, want.abs().max().item())
    assert want.abs().max() < 20 and want.std() > 1e-3      # well-conditioned synthetic checkpoint


def test_finetune_kernel_matches_torch_optimisers(cuda_device):
    """orbit_linear_finetune on IDENTICAL features vs torch.optim.Adam / SGD driven through the oracle's
    batch loop (batch_size 5 over 13 clips => batches of 5,5,3 re-weighted by batch_len/N)."""
    import orbit_b200
    from orbit_b200.finetune import finetune_linear_head
    from oracle.recogniser import OracleRecogniser
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(13, 256, generator=g) * 0.5 + 0.2
    # 0..C-1 as F.cross_entropy needs; class counts (5,4,2,2) deliberately != N/C: for a class with exactly N/C
    # clips the first bias gradient of the zero-initialised head is exactly 0, and Adam's g/(|g|+eps) then turns
    # torch's ~7e-9 rounding noise into a +-0.04 update -- an ill-conditioned quantity that no other summation
    # order (CPU fp64, or this kernel) reproduces.
    labels = torch.tensor([0] * 5 + [1] * 4 + [2] * 2 + [3] * 2)[torch.randperm(13, generator=g)]
    oracle = OracleRecogniser.__new__(OracleRecogniser)
    oracle.batch_size, oracle.feat_dim, oracle.logit_scale, oracle.clip_length = 5, 256, 1.5, 1
    oracle._features = lambda clips, film=None: clips                                   # features are given
    for opt, lr, steps, wd in (('adam', 0.1, 5, 0.0), ('adam', 1e-3, 50, 0.0), ('adam', 0.01, 20, 0.01), ('sgd', 0.5, 10, 0.0)):
        oracle.personalise_finetune(feats, labels, num_grad_steps=steps, learning_rate=lr, optimizer=opt, momentum=0.9,
                                    weight_decay=wd)
        head = orbit_b200.LinearClassifier(256, 1.5)
        head.init(4)
        head.to(cuda_device)
        finetune_linear_head(head, feats.to(cuda_device), labels, 5, steps, lr, opt,
                             {'epsilon': 1e-8, 'weight_decay': wd, 'betas': (0.9, 0.999), 'momentum': 0.9}, 1.5)
        dw = (head.weight.detach().cpu() - oracle.head[0]).abs().max().item()
        db = (head.bias.detach().cpu() - oracle.head[1]).abs().max().item()
        print(f"{opt} lr={lr} steps={steps} wd={wd}: max|dW|={dw:.2e} max|db|={db:.2e} max|W|={oracle.head[0].abs().max():.3f}")
        assert dw <= 2e-4 * max(1.0, oracle.head[0].abs().max().item()) and db <= 2e-4


@pytest.mark.parametrize("n,d,c,steps", [(80, 768, 8, 50), (300, 1280, 10, 20), (13, 260, 4, 20)])
def test_finetune_grid_and_single_cta_kernels(cuda_device, n, d, c, steps):
    """the cooperative-grid kernel (default; (300, 1280, 10) takes its two-level logit reduction, (13, 260, 4) has D % 16 != 0
    and falls back) and the single-CTA kernel against torch.optim.Adam driven through the oracle's batch loop"""
    import orbit_b200
    from orbit_b200 import lib as L
    from orbit_b200.finetune import finetune_linear_head
    from oracle.recogniser import OracleRecogniser
    lib = L.load()
    g = torch.Generator().manual_seed(n + d)
    feats = torch.randn(n, d, generator=g) * 0.5 + 0.1
    labels = (torch.arange(n) * 7 % (c + 1)).clamp(max=c - 1)       # unbalanced class counts (see the test above)
    oracle = OracleRecogniser.__new__(OracleRecogniser)
    oracle.batch_size, oracle.feat_dim, oracle.logit_scale, oracle.clip_length = 64, d, 1.0, 1
    oracle._features = lambda clips, film=None: clips
    oracle.personalise_finetune(feats, labels, num_grad_steps=steps, learning_rate=1e-3, optimizer='adam')
    got = {}
    for mode in (1, 0):
        assert lib.orbit_set_global_option(b'finetune_grid', mode) == 0
        try:
            head = orbit_b200.LinearClassifier(d, 1.0)
            head.init(c)
            head.to(cuda_device)
            finetune_linear_head(head, feats.to(cuda_device), labels, 64, steps, 1e-3, 'adam', {}, 1.0)
            got[mode] = (head.weight.detach().cpu(), head.bias.detach().cpu())
        finally:
            assert lib.orbit_set_global_option(b'finetune_grid', 1) == 0
        dw = (got[mode][0] - oracle.head[0]).abs().max().item()
        db = (got[mode][1] - oracle.head[1]).abs().max().item()
        assert dw <= 2e-4 * max(1.0, oracle.head[0].abs().max().item()) and db <= 2e-4, (mode, dw, db)
    assert (got[0][0] - got[1][0]).abs().max().item() <= 2e-4


def test_finetuner_matches_oracle(cuda_device):
    """MultiStepFewShotRecogniser.personalise (features from the native extractor, then the device-side Adam loop
    on the linear head) vs the oracle, with the reference's default FineTuner hyper-parameters (Adam, lr 1e-3,
    50 steps; utils/args.py:163-168). Logits within 1e-3, identical arg-max."""
    import orbit_b200
    from oracle.recogniser import OracleRecogniser
    from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode
    spec = EpisodeSpec(4, 3, 3, 1, 64)
    oracle = OracleRecogniser('efficientnet_b0', False, 'linear', 1, 5, 1.0, 1991, calibration_frames(64))
    ctx, ctx_y, tgt, _ = make_episode(spec, index=2)
    ctx, ctx_y = ctx[:-1], ctx_y[:-1]          # unbalanced classes (see test_finetune_kernel_matches_torch_optimisers)
    m = orbit_b200.MultiStepFewShotRecogniser('efficientnet_b0', False, 'linear', 1, 5, False)
    m.load_state_dict(oracle.state_dict(), strict=True)
    m._set_device(cuda_device)
    m._send_to_device()
    for opt, lr, steps in (('adam', 1e-3, 50), ('sgd', 0.05, 10)):
        oracle.personalise_finetune(ctx, ctx_y, num_grad_steps=steps, learning_rate=lr, optimizer=opt, momentum=0.9)
        ref = oracle.predict(tgt)
        args = {'num_grad_steps': steps, 'learning_rate': lr, 'optimizer': opt, 'loss_fn': None,
                'extractor_lr_scale': 0.1, 'epsilon': 1e-8, 'weight_decay': 0.0, 'betas': (0.9, 0.999), 'momentum': 0.9}
        m.personalise(ctx, ctx_y, args)       # CPU clips + CPU labels, as multi-step-learner.py:147 passes them
        logits = m.predict(tgt).cpu()
        err = (logits - ref).abs().max().item()
        print(f"{opt} lr={lr} steps={steps}: max|dlogit|={err:.2e} max|logit|={ref.abs().max():.2f}")
        assert err <= 1e-3 * max(1.0, ref.abs().max().item())
        assert torch.equal(logits.argmax(1), ref.argmax(1))
        m._reset()


def test_ragged_batches_and_frame_history(cuda_device, oracle_effnet):
    """batch_size that does not divide the clip count (3 + 3 + 1 clips), a video turned into causal clips with
    attach_frame_history (data/utils.py:8-28), and an empty video."""
    from orbit_b200 import attach_frame_history
    from orbit_b200.synthetic import EpisodeSpec, make_episode
    m = _product(oracle_effnet, cuda_device)
    m.batch_size = 3
    oracle_effnet.batch_size = 3
    spec = EpisodeSpec(way=7, support_clips_per_class=1, query_clips_per_class=1, clip_length=2, frame_size=64)
    ctx, ctx_y, tgt, _ = make_episode(spec, index=11)
    video = tgt[:, 0]                                   # 7 frames of one "video"
    clips = attach_frame_history(video, 2)
    assert clips.shape == (7, 2, 3, 64, 64)
    try:
        oracle_effnet.reset()
        oracle_effnet.personalise(ctx, ctx_y)
        ref = oracle_effnet.predict(clips)
        m.personalise(ctx, ctx_y.to(cuda_device))
        logits = m.predict(clips.to(cuda_device))
        assert (logits.cpu() - ref).abs().max().item() <= 1e-3
        assert torch.equal(logits.argmax(1).cpu(), ref.argmax(1))
        empty = m.predict(clips[:0].to(cuda_device))
        assert empty.shape == (0, 7)
    finally:
        oracle_effnet.batch_size = 256


@pytest.mark.parametrize("head,adapt", [('proto', False), ('versa', True), ('mahalanobis', False)])
def test_predict_video_dedupes_frame_history_bit_exactly(cuda_device, head, adapt):
    """predict_video(frames) == predict(attach_frame_history(frames, L)) bit for bit, with F instead of
    F*L frames through the extractor (reference test loop single-step-learner.py:327-332, data/utils.py:8-28)."""
    import orbit_b200
    from orbit_b200 import attach_frame_history
    from orbit_b200 import lib as L
    from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode
    m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', adapt, head, 4, 8, False, 16)
    m._set_device(cuda_device); m._send_to_device(); m.set_test_mode(True)
    m.feature_extractor.calibrate_batchnorm(calibration_frames(64).to(cuda_device))
    ctx, ctx_y, _, _ = make_episode(EpisodeSpec(3, 2, 1, 4, 64), index=11)
    m.personalise(ctx.to(cuda_device), ctx_y.to(cuda_device))
    video = calibration_frames(64)[:37].to(cuda_device) * 0.8           # 37 frames: ragged last extractor batch
    before = L.launches()
    ref_logits = m.predict(attach_frame_history(video, 4))
    mid = L.launches()
    logits, am = m.predict_video(video, want_argmax=True)
    after = L.launches()
    assert logits.shape == ref_logits.shape == (37, 3)
    assert torch.equal(logits, ref_logits)
    assert torch.equal(am.long(), ref_logits.argmax(1))
    assert (after - mid) < (mid - before)                                # fewer launches: a quarter of the frames
    # CPU-resident frames take the staged path and give the same bits
    assert torch.equal(m.predict_video(video.cpu()), ref_logits)
    assert m.predict_video(video[:0]).shape[0] == 0
