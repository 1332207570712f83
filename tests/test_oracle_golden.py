"""Pins the oracle restatement (oracle/*.py) against golden vectors produced by the UNMODIFIED
reference code (oracle/make_golden.py, run where /root/reference exists). CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import parts
from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def checksum(*tensors):
    return float(sum(t.double().abs().sum().item() for t in tensors))


@pytest.fixture(scope='module')
def gp():
    return np.load(os.path.join(GOLD, 'parts.npz'))


@pytest.fixture(scope='module')
def gr():
    return np.load(os.path.join(GOLD, 'recogniser.npz'))


def head_case(seed, ns, nq, L, D, C, offset=0, stride=1):
    g = torch.Generator().manual_seed(seed)
    sf = torch.randn(ns * L, D, generator=g) * 0.7 + 0.3
    qf = torch.randn(nq * L, D, generator=g) * 0.7 + 0.3
    labels = ((torch.arange(ns) % C) * stride + offset)[torch.randperm(ns, generator=g)]
    return sf, qf, labels


def close(a, b, tol):
    a = torch.as_tensor(np.asarray(a)) if not isinstance(a, torch.Tensor) else a
    b = torch.as_tensor(np.asarray(b))
    assert a.shape == b.shape
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), f"max err {err}"


@pytest.mark.parametrize('i', range(4))
@pytest.mark.parametrize('name,metric', [('proto', 'euclidean'), ('proto_cosine', 'cosine')])
def test_proto_head_vs_reference(gp, i, name, metric):
    seed, ns, nq, L, D, C, off, st = gp[f'{name}{i}_args'].tolist()
    sf, qf, labels = head_case(seed, ns, nq, L, D, C, off, st)
    assert checksum(sf, qf, labels) == pytest.approx(float(gp[f'{name}{i}_checksum']), rel=1e-12)
    w, b = parts.proto_configure(parts.pool_clips(sf, L), labels, metric)
    close(w, gp[f'{name}{i}_weight'], 1e-6)
    if metric == 'euclidean':
        close(b, gp[f'{name}{i}_bias'], 1e-6)
    logits = parts.proto_predict(parts.pool_clips(qf, L), w, b, 1.7, metric)
    close(logits, gp[f'{name}{i}_logits'], 2e-6)
    assert torch.equal(logits.argmax(1), torch.as_tensor(gp[f'{name}{i}_logits']).argmax(1))


def test_identities_from_survey_8c(gp):
    """(i) logits == s*(2 q.mu - mu.mu); (ii) cosine == normalize(q) @ normalize(mu)^T."""
    sf, qf, labels = head_case(100, 200, 80, 8, 1280, 5)
    f, q = parts.pool_clips(sf, 8), parts.pool_clips(qf, 8)
    mu, classes = parts.class_means(f, labels)
    ref = torch.as_tensor(gp['proto0_logits'])
    assert ((1.7 * (2 * q @ mu.t() - (mu * mu).sum(1))) - ref).abs().max() < 2e-3
    cos = torch.nn.functional.normalize(q, dim=1) @ torch.nn.functional.normalize(mu, dim=1).t()
    assert (1.7 * cos - torch.as_tensor(gp['proto_cosine0_logits'])).abs().max() < 1e-5


def test_versa_mahalanobis_linear_vs_reference(gp):
    sf, qf, labels = head_case(200, 40, 16, 1, 128, 5)
    assert checksum(sf, qf, labels) == pytest.approx(float(gp['versa_checksum']), rel=1e-12)
    w, b = parts.versa_configure(sf, labels, parts.init_versa_params(128, seed=7))
    close(w, gp['versa_weight'], 1e-6)
    close(b, gp['versa_bias'], 1e-6)
    close(parts.linear_predict(qf, w, b, 0.5), gp['versa_logits'], 2e-6)
    sf, qf, labels = head_case(201, 40, 16, 1, 32, 4)
    means, precs = parts.mahalanobis_configure(sf, labels)
    close(means, gp['maha_means'], 1e-6)
    close(parts.mahalanobis_predict(qf, means, precs, 2.0), gp['maha_logits'], 1e-4)
    g = torch.Generator().manual_seed(202)
    w, b = torch.randn(6, 64, generator=g), torch.randn(6, generator=g)
    close(parts.linear_predict(torch.randn(10, 64, generator=g), w, b, 3.0), gp['linear_logits'], 1e-6)


def test_set_encoder_and_film_generator_vs_reference(gp):
    sp = parts.init_set_encoder_params(seed=11)
    g = torch.Generator().manual_seed(203)
    for size in (84, 224):
        clips = torch.randn(3, 2, 3, size, size, generator=g)
        assert checksum(clips) == pytest.approx(float(gp[f'setenc{size}_checksum']), rel=1e-12)
        reps = parts.set_encoder_forward(clips, sp)
        close(reps, gp[f'setenc{size}_reps'], 1e-5)
        close(parts.task_embedding([reps[:2], reps[2:]]), gp[f'setenc{size}_agg'], 1e-5)
    names = ['blocks.1.0.bn2.bias', 'blocks.1.0.bn2.weight', 'bn1.bias', 'bn1.weight']
    sizes = {'blocks.1.0.bn2.bias': 96, 'blocks.1.0.bn2.weight': 96, 'bn1.bias': 32, 'bn1.weight': 32}
    g = torch.Generator().manual_seed(204)
    initial = {n: torch.randn(sizes[n], generator=g) for n in names}
    gen = parts.init_film_generator_params([sizes[n] for n in sorted(names)], seed=13)
    z = torch.randn(1, 64, generator=g)
    film, l2 = parts.film_generate(z, sorted(names), gen, initial)
    for n in names:
        close(film[n], gp['film_' + n], 1e-6)
    assert float(l2) == pytest.approx(float(gp['film_l2']), rel=1e-5)


@pytest.mark.parametrize('L', [1, 3, 8])
def test_attach_frame_history_vs_reference(gp, L):
    from orbit_b200 import attach_frame_history
    fr = torch.arange(7 * 2, dtype=torch.float32).reshape(7, 2, 1, 1)
    assert np.array_equal(parts.attach_frame_history(fr, L).numpy(), gp[f'history{L}'])
    assert np.array_equal(attach_frame_history(fr, L).numpy(), gp[f'history{L}'])   # the product's helper too


CASES = {
    'proto_b0': ('efficientnet_b0', False, 'proto'),
    'cosine_b0': ('efficientnet_b0', False, 'proto_cosine'),
    'cnaps_b0': ('efficientnet_b0', True, 'versa'),
    'protofilm_b0': ('efficientnet_b0', True, 'proto'),
    'simplecnaps_b0': ('efficientnet_b0', True, 'mahalanobis'),
    'proto_vit': ('vit_b_32', False, 'proto'),
}


@pytest.mark.parametrize('tag', list(CASES))
def test_recogniser_vs_reference(gr, tag):
    """Whole personalise()+predict() of the oracle vs the unmodified reference recogniser."""
    extractor, adapt, head = CASES[tag]
    way, sup, qry, L, size = gr[tag + '_spec'].tolist()
    spec = EpisodeSpec(way, sup, qry, L, size)
    oracle = OracleRecogniser(extractor, adapt, head, L, 4, 1.0, 1991, calibration_frames(size))
    ctx, ctx_y, tgt, _ = make_episode(spec, index=1)
    assert checksum(ctx, tgt, ctx_y) == pytest.approx(float(gr[tag + '_checksum']), rel=1e-12)
    oracle.personalise(ctx, ctx_y)
    logits = oracle.predict(tgt)
    ref = torch.as_tensor(gr[tag + '_logits'])
    tol = 2e-4 if head == 'mahalanobis' else 2e-5
    close(logits, ref, tol)
    assert torch.equal(logits.argmax(1), ref.argmax(1))
    if adapt:
        close(oracle.film_dict['bn1.weight'], gr[tag + '_film_bn1_weight'], 1e-5)


def test_finetuner_vs_reference(gr):
    spec = EpisodeSpec(4, 3, 2, 1, 64)
    oracle = OracleRecogniser('efficientnet_b0', False, 'linear', 1, 5, 1.0, 1991, calibration_frames(64))
    ctx, ctx_y, tgt, _ = make_episode(spec, index=2)
    assert checksum(ctx, tgt, ctx_y) == pytest.approx(float(gr['finetune_checksum']), rel=1e-12)
    for recompute in (False, True):
        oracle.personalise_finetune(ctx, ctx_y, num_grad_steps=5, learning_rate=0.1, recompute_features=recompute)
        close(oracle.head[0], gr['finetune_weight'], 1e-4)
        close(oracle.head[1], gr['finetune_bias'], 1e-4)
        close(oracle.predict(tgt), gr['finetune_logits'], 1e-4)
    oracle.personalise_finetune(ctx[:-1], ctx_y[:-1], num_grad_steps=5, learning_rate=0.1)    # class counts 3,3,3,2
    close(oracle.head[0], gr['finetune2_weight'], 1e-5)
    close(oracle.head[1], gr['finetune2_bias'], 1e-5)
    close(oracle.predict(tgt), gr['finetune2_logits'], 1e-5)


def test_finetuner_film_vs_reference(gr):
    """FineTuner + FiLM (few_shot_recognisers.py:196-198,207-246): the oracle restatement vs the unmodified reference."""
    oracle = OracleRecogniser('efficientnet_b0', False, 'linear', 1, 5, 1.0, 1991, calibration_frames(64))
    ctx, ctx_y, tgt, _ = make_episode(EpisodeSpec(4, 3, 2, 1, 64), index=2)
    oracle.personalise_finetune_film(ctx[:-1], ctx_y[:-1], num_grad_steps=3, learning_rate=0.01)
    close(oracle.head[0], gr['finetune_film_weight'], 1e-5)
    close(oracle.head[1], gr['finetune_film_bias'], 1e-5)
    sd = oracle.extractor.state_dict()
    for k in ('bn1.weight', 'blocks.1.0.bn2.weight', 'blocks.3.1.bn2.bias', 'bn2.bias'):
        close(sd[k], gr['finetune_film_' + k], 1e-5)
    close(oracle.predict(tgt), gr['finetune_film_logits'], 1e-4)


def test_efficientnet_v2_s_structure():
    """tf_efficientnetv2_s with num_classes=0: 20,177,488 parameters (timm's published 21.46 M minus the 1280x1000+1000
    classifier), 84 FiLM tensors (2 root + 2 ConvBnAct + 8 EdgeResidual + 30 InvertedResidual sites, weight + bias)."""
    from oracle import backbones, parts
    m = backbones.build('efficientnet_v2_s')
    assert sum(p.numel() for p in m.parameters()) == 20177488
    names = parts.film_parameter_names('efficientnet_v2_s', m)
    assert len(names) == 84 and names[:2] == ['bn1.weight', 'bn1.bias'] and names[-2:] == ['bn2.weight', 'bn2.bias']
    assert 'blocks.0.1.bn1.weight' in names and 'blocks.2.3.bn1.bias' in names and 'blocks.5.14.bn2.weight' in names
    assert not any('.bn3.' in n for n in names)


def test_evaluator_oracle_matches_reference_golden():
    """oracle/evaluator.py vs the reference TestEvaluator's own output (tests/golden/evaluator.npz)."""
    import os
    import numpy as np
    from oracle import evaluator as ev
    from oracle.make_golden import evaluator_case
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'evaluator.npz'))
    users = evaluator_case()
    assert abs(sum(float(np.abs(lg).sum()) for t in users for vs in t for _, lg, _ in vs) - float(g['input_checksum'])) < 1e-3
    dedup = [[[(label, lg[np.unique(np.array(paths), return_index=True)[1]]) for label, lg, paths in videos]
              for videos in tasks] for tasks in users]
    for stat in ('frame_acc', 'frames_to_recognition'):
        got = ev.mean_stats(stat, dedup)
        for level in ('user', 'object', 'task', 'video'):
            assert np.array_equal(np.array(got[level]), g[f'all.{level}.{stat}']), (stat, level)
        for u in range(2):
            got_u = ev.mean_stats(stat, dedup[u:u + 1])
            for level in ('user', 'object', 'task', 'video'):
                assert np.array_equal(np.array(got_u[level]), g[f'current{u}.{level}.{stat}']), (stat, level, u)
