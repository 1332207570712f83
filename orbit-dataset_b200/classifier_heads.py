"""Classifier heads behind the reference's ``configure()/predict()/reset()`` interface.

Mirrors reference ``model/classifier_heads.py`` (class names, attributes ``weight``/``bias`` as
``nn.Parameter`` while personalised and ``None`` after ``reset()``, error behaviour) and
``model/poolers.py``; the arithmetic runs in liborbit_b200's fused head kernels.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import lib as L

_METRICS = {'euclidean': 0, 'cosine': 1}


def _class_index(labels: torch.Tensor):
    """Rank of each label among the sorted distinct labels (torch.unique order,
    classifier_heads.py:96,246-248). Done on the host: labels are a few hundred int64s and the class
    COUNT is needed there anyway to size the [C, D] / [Nq, C] outputs."""
    lab = labels.detach().cpu().numpy().astype(np.int64).reshape(-1)
    classes, inverse = np.unique(lab, return_inverse=True)
    return classes, inverse.astype(np.int32)


class ClassIndex:
    """The label bookkeeping of one task, computed BEFORE the support set is enqueued.

    Reading device labels back is a host sync (the reference has several per class: ``unique``, ``nonzero``,
    classifier_heads.py:96-101). Done inside ``configure`` it would block the host until the whole support pass has
    drained, and the query set's H2D copies could not be issued meanwhile; done first, it waits for nothing."""

    def __init__(self, labels: torch.Tensor, device=None):
        self.classes, self.index = _class_index(labels)
        self.num_classes, self.num_clips = len(self.classes), len(self.index)
        device = labels.device if device is None else device
        self.index_dev = torch.from_numpy(self.index).to(device, non_blocking=True) if torch.device(device).type == 'cuda' else None

    @staticmethod
    def of(labels, class_index, device):
        if class_index is None:
            return ClassIndex(labels, device)
        assert class_index.num_clips == labels.numel(), "class index was built for another label tensor"
        if class_index.index_dev is None or class_index.index_dev.device != torch.device(device):
            class_index.index_dev = torch.from_numpy(class_index.index).to(device, non_blocking=True)
        return class_index


class MeanPooler(nn.Module):
    """poolers.py:7-16."""

    def __init__(self, T, dim=1):
        super().__init__()
        self.T = T
        self.dim = dim

    def forward(self, x):
        L.require_cuda(x, "features")
        x = x.contiguous().float()
        n, d = x.shape[0] // self.T, x.shape[-1]
        out = torch.empty(n, d, dtype=torch.float32, device=x.device)
        L.check(L.load().orbit_pool_clips(L.ptr(x), n, self.T, d, L.ptr(out), L.stream_ptr(x.device)), "orbit_pool_clips")
        L.count_launches(1)
        return out


def _head_predict(features, clip_length, weight, bias, metric, logit_scale, want_argmax=False):
    if features.requires_grad and torch.is_grad_enabled() and not want_argmax:
        # meta-training: logits with a graph back to the query features (the head's weights are leaf Parameters)
        from .training import HeadPredictFn
        return HeadPredictFn.apply(features, clip_length, weight.detach(), bias.detach() if bias is not None else None, metric,
                                   logit_scale)
    return _head_predict_raw(features, clip_length, weight, bias, metric, logit_scale, want_argmax)


def _head_predict_raw(features, clip_length, weight, bias, metric, logit_scale, want_argmax=False):
    L.require_cuda(features, "features")
    features = features.detach().contiguous().float()
    n = features.shape[0] // clip_length
    c, d = weight.shape
    if features.shape[1] != d:
        raise ValueError(f"feature dim {features.shape[1]} does not match classifier dim {d}")
    logits = torch.empty(n, c, dtype=torch.float32, device=features.device)
    argmax = torch.empty(n, dtype=torch.int32, device=features.device) if want_argmax else None
    L.check(L.load().orbit_head_predict(L.ptr(features), n, clip_length, d, L.ptr(weight.detach()),
                                        L.ptr(bias.detach() if bias is not None else None), c, metric,
                                        float(logit_scale), L.ptr(logits), L.ptr(argmax), L.stream_ptr(features.device)),
            "orbit_head_predict")
    L.count_launches(1)
    return (logits, argmax) if want_argmax else logits


class LinearClassifier(nn.Module):
    """classifier_heads.py:38-79."""

    def __init__(self, feat_dim, logit_scale: float = 1.0):
        super().__init__()
        self.feat_dim = feat_dim
        self.logit_scale = logit_scale
        self.weight = None
        self.bias = None

    def init(self, num_classes: int):
        self.weight = nn.Parameter(torch.zeros(num_classes, self.feat_dim), requires_grad=True)
        self.bias = nn.Parameter(torch.zeros(num_classes), requires_grad=True)

    def predict(self, features, ops_counter=None, clip_length=1, want_argmax=False):
        if self.weight is None:
            raise AttributeError("Weight and/or bias not set - is model personalised?")
        if ops_counter:   # classifier_heads.py:72-73
            ops_counter.add_macs(self.weight.size(0) * (features.size(0) // clip_length) * self.feat_dim)
        return _head_predict(features, clip_length, self.weight, self.bias, 0, self.logit_scale, want_argmax)

    def reset(self):
        self.weight = None
        self.bias = None


class HeadClassifier(nn.Module):
    def __init__(self, logit_scale: float = 1.0):
        super().__init__()
        self.logit_scale = logit_scale

    @staticmethod
    def _count_class_reps(ops_counter, num_context, feat_dim, num_classes):
        """MACs the reference adds in ``_build_class_reps`` (classifier_heads.py:101-103): per class, selecting its
        rows (N) and mean-pooling them (n_c * D); summed over classes = C*N + N*D."""
        if ops_counter:
            ops_counter.add_macs(num_classes * num_context + num_context * feat_dim)


class PrototypicalClassifier(HeadClassifier):
    """classifier_heads.py:182-263."""

    def __init__(self, logit_scale: float = 1.0, distance_fn: str = 'euclidean'):
        super().__init__(logit_scale)
        self.distance_fn = distance_fn
        self._scratch = None
        self.reset()

    def reset(self):
        self.weight = None
        if self.distance_fn == 'euclidean':
            self.bias = None
        self.classes = None

    def configure(self, context_features, context_labels, ops_counter=None, clip_length=1, class_index=None):
        """``context_features``: clip features [N, D] (reference contract) or, with ``clip_length=L``,
        the FRAME features [N*L, D] -- pooling is then fused into the same launch. ``class_index``: a ``ClassIndex``
        of ``context_labels`` built earlier (keeps this call free of host syncs)."""
        if self.distance_fn not in _METRICS:
            raise ValueError(f"Distance function {self.distance_fn} not valid.")
        L.require_cuda(context_features, "context_features")
        assert context_features.size(0) == context_labels.size(0) * clip_length, \
            "context features and labels are different sizes!"
        lib = L.load()
        feats = context_features.contiguous().float()
        dev = feats.device
        ci = ClassIndex.of(context_labels, class_index, dev)
        classes, idx_dev = ci.classes, ci.index_dev
        c, d, n = ci.num_classes, feats.shape[1], ci.num_clips
        if c > 64:
            raise ValueError("orbit_b200 supports at most 64 classes per task")
        need = lib.orbit_proto_configure_scratch_bytes(64, d)
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != dev:
            self._scratch = torch.zeros(need, dtype=torch.uint8, device=dev)
        weight = torch.empty(c, d, dtype=torch.float32, device=dev)
        euclid = self.distance_fn == 'euclidean'
        bias = torch.empty(c, dtype=torch.float32, device=dev) if euclid else None
        L.check(lib.orbit_proto_configure(L.ptr(feats), L.ptr(idx_dev), n, clip_length, d, c, _METRICS[self.distance_fn],
                                          L.ptr(weight), L.ptr(bias), None, L.ptr(self._scratch), L.stream_ptr(dev)),
                "orbit_proto_configure")
        L.count_launches(1)
        # nn.Parameter wrap as in the reference (classifier_heads.py:261-263): cuts the autograd graph
        self.weight = nn.Parameter(weight)
        if euclid:
            self.bias = nn.Parameter(bias)
        self.classes = torch.from_numpy(classes)
        self._count_class_reps(ops_counter, n, d, c)
        if ops_counter:   # classifier_heads.py:256-259: 2*mu, mu.mu^T, negation -- D MACs each per class
            ops_counter.add_macs(3 * c * d)

    def predict(self, features, ops_counter=None, clip_length=1, want_argmax=False):
        if self.weight is None or (self.distance_fn == 'euclidean' and self.bias is None):
            raise AttributeError("Weight and/or bias not set - is model personalised?")
        if self.distance_fn not in _METRICS:
            raise ValueError(f"Distance function {self.distance_fn} not valid.")
        bias = self.bias if self.distance_fn == 'euclidean' else None
        if ops_counter:   # classifier_heads.py:221-228
            nq, d, c = features.size(0) // clip_length, features.size(1), self.weight.size(0)
            ops_counter.add_macs(nq * d * c if self.distance_fn == 'euclidean' else 2 * nq * d * c + c * d + nq * d)
        return _head_predict(features, clip_length, self.weight, bias, _METRICS[self.distance_fn], self.logit_scale,
                             want_argmax)
