"""CPU-only tests of the host-side mirror of the reference interface (no kernels run)."""
import pytest
import torch

import orbit_b200
from orbit_b200 import OrbitError


def make(classifier='proto', adapt=False):
    return orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', adapt, classifier, 2, 4, False, 16)


def test_constructor_surface_and_errors():
    m = make()
    assert m.clip_length == 2 and m.batch_size == 4 and m.logit_scale == 1.0
    assert m.classifier_name == 'proto' and m.film_parameter_names is None
    assert m.film_generator.regularization_term() == 0
    with pytest.raises(ValueError, match="Classifier bogus not valid"):
        make('bogus')
    with pytest.raises(ValueError, match="Invalid feature_extractor_name"):
        orbit_b200.SingleStepFewShotRecogniser('resnet50', False, 'proto', 1, 1, False, 16)


def test_no_cpu_fallback():
    """The product must fail loudly off-GPU instead of silently computing on the CPU."""
    m = make()
    m.set_test_mode(True)
    clips, labels = torch.zeros(4, 2, 3, 64, 64), torch.tensor([0, 0, 1, 1])
    with pytest.raises(OrbitError, match="no CPU fallback"):
        m.personalise(clips, labels)
    with pytest.raises(OrbitError):
        m.feature_extractor(torch.zeros(1, 3, 64, 64))
    with pytest.raises(OrbitError):
        orbit_b200.PrototypicalClassifier().configure(torch.zeros(4, 8), labels)


def test_predict_before_personalise_raises_attribute_error():
    head = orbit_b200.PrototypicalClassifier()
    with pytest.raises(AttributeError, match="is model personalised"):
        head.predict(torch.zeros(2, 8))


def test_state_dict_roundtrip_and_reset():
    a, b = make(), make()
    with torch.no_grad():
        for p in a.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    b.load_state_dict(a.state_dict(), strict=True)
    assert torch.equal(a.feature_extractor._blob, b.feature_extractor._blob)
    assert all(not p.requires_grad for p in a.feature_extractor.parameters())   # learn_extractor=False freezes
    a._reset()
    assert a.film_dict is None and a.classifier.weight is None


def test_batch_indices_and_frame_history():
    from orbit_b200 import attach_frame_history, get_batch_indices
    assert get_batch_indices(0, 10, 4) == (0, 4)
    assert get_batch_indices(2, 10, 4) == (8, 10)
    fr = torch.arange(5, dtype=torch.float32).reshape(5, 1, 1, 1)
    h = attach_frame_history(fr, 3)
    assert h.shape == (5, 3, 1, 1, 1)
    assert h[:, :, 0, 0, 0].tolist() == [[0, 0, 0], [0, 0, 1], [0, 1, 2], [1, 2, 3], [2, 3, 4]]


def test_class_index_follows_torch_unique_order():
    from orbit_b200.classifier_heads import _class_index
    labels = torch.tensor([107, 100, 114, 100, 107])
    classes, idx = _class_index(labels)
    assert classes.tolist() == torch.unique(labels).tolist() == [100, 107, 114]
    assert idx.tolist() == [1, 0, 2, 0, 1]


def test_multistep_pops_learning_args_like_the_reference():
    m = orbit_b200.MultiStepFewShotRecogniser('efficientnet_b0', False, 'linear', 1, 4, False)
    args = {'num_grad_steps': 2, 'learning_rate': 0.1, 'optimizer': 'adam', 'loss_fn': None, 'extractor_lr_scale': 0.1,
            'epsilon': 1e-8}
    with pytest.raises(OrbitError):     # CPU tensors are refused, but only after the dict was consumed
        m.personalise(torch.zeros(2, 1, 3, 64, 64), torch.tensor([0, 1]), args)
    assert 'num_grad_steps' not in args and 'epsilon' in args   # few_shot_recognisers.py:218-223 mutates the dict


def test_synthetic_episode_shapes():
    from orbit_b200.synthetic import S2, EpisodeSpec, make_episode
    assert S2.way * S2.support_clips_per_class == 200 and S2.way * S2.query_clips_per_class == 80
    ctx, cy, tgt, ty = make_episode(EpisodeSpec(5, 2, 3, 2, 32), index=4)
    assert ctx.shape == (10, 2, 3, 32, 32) and tgt.shape == (15, 2, 3, 32, 32)
    assert sorted(cy.tolist()) == [0, 0, 1, 1, 2, 2, 3, 3, 4, 4]
    ctx2, *_ = make_episode(EpisodeSpec(5, 2, 3, 2, 32), index=4)
    assert torch.equal(ctx, ctx2)


def test_engine_macs_match_published_model_costs():
    """orbit_engine_macs (the analytic stand-in for the thop trace of utils/ops_counter.py:82-88) against the published
    224x224 costs of the same architectures: EfficientNet-B0 0.39 GMACs, EfficientNetV2-S 8.4 GMACs @384 (x (224/384)^2),
    ResNet-18 1.82 GMACs (incl. the 0.5 M fc), ViT-B/32 4.41 GMACs, ViT-S/32 1.15 GMACs (timm model cards)."""
    import torch
    from orbit_b200.feature_extractors import FeatureExtractor
    published = {'efficientnet_b0': 0.39e9, 'efficientnet_v2_s': 8.44e9 * (224 / 384) ** 2, 'resnet18': 1.82e9,
                 'vit_b_32': 4.41e9, 'vit_s_32': 1.15e9}
    for name, want in published.items():
        fe = FeatureExtractor(name)
        got = fe.count_macs(torch.empty(1, 3, 224, 224))
        assert abs(got - want) / want < 0.02, (name, got, want)
        assert fe.count_macs(torch.empty(2, 5, 3, 224, 224)) == 10 * got      # clips x frames


def test_ops_counter_interface_and_head_formulas():
    """OpsCounter mirrors utils/ops_counter.py:10-99; head MACs follow the reference's hand-written counts."""
    import torch
    from orbit_b200 import OpsCounter
    from orbit_b200.ops_counter import clever_format
    from orbit_b200.classifier_heads import HeadClassifier

    class Fake(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(7))

        def count_macs(self, x):
            return 100 * x.shape[0]

    oc = OpsCounter(count_backward=True)
    oc.compute_macs(Fake(), torch.zeros(3, 2))
    assert oc.get_task_macs() == 600 and oc.task_params_counter == 7          # forward + backward multiplier
    oc.add_macs(5)
    assert oc.get_task_macs() == 605
    HeadClassifier._count_class_reps(oc, 10, 4, 3)                            # C*N + N*D
    assert oc.get_task_macs() == 605 + 30 + 40
    oc.task_complete()
    assert oc.get_task_macs() == 0 and oc.get_task_params() == oc.base_params_counter
    with pytest.raises(TypeError):
        oc.compute_macs(torch.nn.Linear(2, 2), torch.zeros(1, 2))
    assert clever_format([4007548, 1234, 12]) == ('4.01M', '1.23K', '12.00B')


def test_host_stager_plan():
    """_HostStager._plan: passes cover the call exactly, ramp first, no tiny tail, and no ramp for a call queued while
    the previous call is still running on the device (predict() after personalise()); deterministic (no timers)."""
    from orbit_b200.few_shot_recognisers import _HostStager
    st = _HostStager.__new__(_HostStager)
    st.chunk_frames, st.ramp, st.copy_frames = 1600, (96, 224, 480), 160
    assert st._plan(1600, behind=False) == [96, 224, 480, 800]
    assert st._plan(640, behind=True) == [640]                 # data will be there before the device gets to it
    assert st._plan(640, behind=False) == [96, 224, 320]
    assert st._plan(830, behind=False) == [96, 224, 510]       # a 30-frame tail is merged into the last pass
    assert st._plan(1630, behind=True) == [816, 814]           # ... or the last two passes are rebalanced when the
    assert st._plan(3230, behind=True) == [1600, 816, 814]     # merged pass would exceed chunk_frames (engine would re-split)
    st.chunk_frames, st.ramp = 2, (4, 6)
    for total in (1, 7, 30):
        for behind in (False, True):
            plan = st._plan(total, behind)
            assert sum(plan) == total and all(n > 0 for n in plan)


def test_training_paths_without_backward_kernels_are_refused():
    """Every entry that needs gradients through the extractor refuses loudly instead of
    returning graph-less logits that would fail later inside loss.backward()."""
    m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', True, 'versa', 1, 4, False, 8)
    clips, labels = torch.zeros(2, 1, 3, 64, 64), torch.tensor([0, 1])
    m.set_test_mode(False)                                   # CNAPs meta-training (single-step-learner.py:196-243) IS implemented
    with pytest.raises(OrbitError):                          # (tests/test_gpu_meta_train.py): only the GPU is missing here
        m.personalise_with_lite(clips, labels)
    with pytest.raises(OrbitError):
        m.personalise(clips, labels)
    m.set_test_mode(True)
    with torch.no_grad(), pytest.raises(OrbitError):         # test mode passes the gate (then fails only for lack of a GPU)
        m.personalise(clips, labels)
    unfrozen = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 1, 4, True, 8)
    unfrozen.set_test_mode(False)
    with pytest.raises(NotImplementedError, match="training the extractor's own weights"):
        unfrozen.personalise(clips, labels)
    args = {'num_grad_steps': 2, 'learning_rate': 0.1, 'optimizer': 'adam', 'loss_fn': None, 'extractor_lr_scale': 0.1}
    ft = orbit_b200.MultiStepFewShotRecogniser('efficientnet_b0', True, 'linear', 1, 4, False)
    with pytest.raises(OrbitError):                          # FineTuner + FiLM IS implemented (tests/test_gpu_train.py): only the GPU is missing here
        ft.personalise(clips, labels, dict(args))
    assert any(p.requires_grad for p in ft.feature_extractor.parameters())      # unfreeze_film (film.py:76-79)
    unfrozen_ft = orbit_b200.MultiStepFewShotRecogniser('efficientnet_b0', False, 'linear', 1, 4, True)
    with pytest.raises(NotImplementedError, match="weight-gradient"):
        unfrozen_ft.personalise(clips, labels, dict(args))
    assert ft.personalise_with_lite(clips, labels) is None   # the reference's own no-op (few_shot_recognisers.py:260-261)


def test_feature_extractor_reprepares_after_weight_updates():
    """prepare() caches BN folds / weight splits; every way of changing weights must invalidate that cache
    (parameters are views bound with .data, whose version counters are their own -- the blob's does not move)."""
    from orbit_b200.feature_extractors import FeatureExtractor
    fe = FeatureExtractor('efficientnet_b0')
    v0 = fe._state_version()
    other = FeatureExtractor('efficientnet_b0', seed=5)
    fe._prepared_key = 'stale'
    fe.load_state_dict(other.state_dict())
    assert fe._prepared_key is None and fe._state_version() != v0
    assert torch.equal(fe._blob, other._blob)                    # the parameters really are views of the blob
    v1 = fe._state_version()
    with torch.no_grad():
        next(fe.parameters()).mul_(2.0)                          # what an optimiser step does
    assert fe._state_version() != v1
    v2 = fe._state_version()
    fe.conv_stem.weight.data.copy_(other.conv_stem.weight)       # invisible alias write: needs mark_dirty()
    assert fe._state_version() == v2
    fe._prepared_key = 'stale'
    fe.mark_dirty()
    assert fe._prepared_key is None


def test_class_index_is_built_without_touching_the_features():
    from orbit_b200.classifier_heads import ClassIndex
    ci = ClassIndex(torch.tensor([7, 3, 7, 9]), device='cpu')
    assert ci.num_classes == 3 and ci.num_clips == 4 and ci.index.tolist() == [1, 0, 1, 2] and ci.index_dev is None
    with pytest.raises(AssertionError, match="another label tensor"):
        ClassIndex.of(torch.tensor([1, 2]), ci, 'cpu')


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm: oracle port of the reference's PyTorch path) prints ONE JSON line with the
    keys the driver reads; it needs no GPU and no liborbit_b200."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "episodes_per_sec" and d["unit"] == "episodes/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("S2")
