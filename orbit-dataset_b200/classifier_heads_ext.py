"""Versa head (reference model/classifier_heads.py:121-180, model/mlps.py:33-50) and Mahalanobis head
(classifier_heads.py:265-368) on the native kernels."""
import torch
import torch.nn as nn

from . import lib as L
from .classifier_heads import ClassIndex, HeadClassifier, _head_predict


class DenseResidualBlock(nn.Module):
    """mlps.py:33-50: linear1-ELU-linear2-ELU-linear3 (+ identity skip when in_size == out_size)."""

    def __init__(self, in_size, out_size):
        super().__init__()
        self.linear1 = nn.Linear(in_size, out_size)
        self.linear2 = nn.Linear(out_size, out_size)
        self.linear3 = nn.Linear(out_size, out_size)
        self.in_size, self.out_size = in_size, out_size

    def count_macs(self, x):
        return int(x.shape[0]) * (self.in_size * self.out_size + 2 * self.out_size * self.out_size)

    def forward(self, x):
        lib = L.load()
        L.require_cuda(x, "class representations")
        x = x.contiguous().float()
        rows, dev = x.shape[0], x.device
        if self.out_size % 4:   # bias processor (out_size 1): pad the width to 4 so that rows stay 16-byte aligned
            return self._forward_narrow(x)

        def dense(inp, lin, act, skip=None):
            out = torch.empty(rows, lin.out_features, dtype=torch.float32, device=dev)
            L.check(lib.orbit_dense_rows(L.ptr(inp), L.ptr(lin.weight.detach()), L.ptr(lin.bias.detach()), L.ptr(skip),
                                         L.ptr(out), rows, lin.in_features, lin.out_features, act, L.stream_ptr(dev)),
                    "orbit_dense_rows")
            L.count_launches(1)
            return out

        h = dense(x, self.linear1, 3)
        h = dense(h, self.linear2, 3)
        return dense(h, self.linear3, 0, x if self.in_size == self.out_size else None)

    def _forward_narrow(self, x):
        """out_size == 1: linear1 is a D->1 dense layer (native); the two 1->1 'layers' that follow are scalar
        affine maps per class, applied by the same kernel on a zero-padded 4-wide row."""
        lib = L.load()
        rows, dev = x.shape[0], x.device
        out1 = torch.empty(rows, 1, dtype=torch.float32, device=dev)
        L.check(lib.orbit_dense_rows(L.ptr(x), L.ptr(self.linear1.weight.detach()), L.ptr(self.linear1.bias.detach()), None,
                                     L.ptr(out1), rows, self.in_size, 1, 3, L.stream_ptr(dev)), "orbit_dense_rows")
        pad = torch.zeros(rows, 4, dtype=torch.float32, device=dev)
        for lin, act in ((self.linear2, 3), (self.linear3, 0)):
            pad[:, :1] = out1
            w = torch.zeros(1, 4, dtype=torch.float32, device=dev)
            w[:, :1] = lin.weight.detach()
            out1 = torch.empty(rows, 1, dtype=torch.float32, device=dev)
            L.check(lib.orbit_dense_rows(L.ptr(pad), L.ptr(w), L.ptr(lin.bias.detach()), None, L.ptr(out1), rows, 4, 1, act,
                                         L.stream_ptr(dev)), "orbit_dense_rows")
        L.count_launches(3)
        return out1


class VersaClassifier(HeadClassifier):
    """classifier_heads.py:121-180: class means -> two hyper-networks -> rows of a linear layer."""

    def __init__(self, in_size, logit_scale: float = 1.0):
        super().__init__(logit_scale)
        self.weight_processor = DenseResidualBlock(in_size, in_size)
        self.bias_processor = DenseResidualBlock(in_size, 1)
        self._scratch = None
        self.reset()

    def reset(self):
        self.weight = None
        self.bias = None

    def predict(self, target_features, ops_counter=None, clip_length=1, want_argmax=False):
        if self.weight is None or self.bias is None:
            raise AttributeError("Weight and/or bias not set - is model personalised?")
        return _head_predict(target_features, clip_length, self.weight, self.bias, 0, self.logit_scale, want_argmax)

    def configure(self, context_features, context_labels, ops_counter=None, clip_length=1, class_index=None):
        L.require_cuda(context_features, "context_features")
        assert context_features.size(0) == context_labels.size(0) * clip_length, \
            "context features and labels are different sizes!"
        lib = L.load()
        feats = context_features.contiguous().float()
        dev = feats.device
        ci = ClassIndex.of(context_labels, class_index, dev)
        idx_dev = ci.index_dev
        c, d, n = ci.num_classes, feats.shape[1], ci.num_clips
        need = lib.orbit_proto_configure_scratch_bytes(64, d)
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != dev:
            self._scratch = torch.zeros(need, dtype=torch.uint8, device=dev)
        two_mu = torch.empty(c, d, dtype=torch.float32, device=dev)
        mu = torch.empty(c, d, dtype=torch.float32, device=dev)
        # class means via the prototype kernel (cosine mode: no bias output needed)
        L.check(lib.orbit_proto_configure(L.ptr(feats), L.ptr(idx_dev), n, clip_length, d, c, 1, L.ptr(two_mu), None,
                                          L.ptr(mu), L.ptr(self._scratch), L.stream_ptr(dev)), "orbit_proto_configure")
        L.count_launches(1)
        self.weight = nn.Parameter(self.weight_processor(mu))
        self.bias = nn.Parameter(self.bias_processor(mu).reshape(c))
        self._count_class_reps(ops_counter, n, d, c)
        if ops_counter:   # classifier_heads.py:175-177: both hyper-networks traced once per class
            for _ in range(c):
                ops_counter.compute_macs(self.weight_processor, mu[:1])
                ops_counter.compute_macs(self.bias_processor, mu[:1])


class MahalanobisClassifier(HeadClassifier):
    """classifier_heads.py:265-368 (Simple CNAPs): per-class regularised covariances -> precisions; logits are
    negative Mahalanobis distances. Attributes as in the reference: means, precisions, task_mean, task_precision."""

    def __init__(self, logit_scale: float = 1.0):
        super().__init__(logit_scale)
        self.reset()

    def reset(self):
        self.means = None
        self.precisions = None
        self.task_mean = None
        self.task_precision = None

    @staticmethod
    def _pool(features, clip_length):
        if clip_length == 1:
            return features.contiguous().float()
        n, d = features.shape[0] // clip_length, features.shape[1]
        out = torch.empty(n, d, dtype=torch.float32, device=features.device)
        f = features.contiguous().float()
        L.check(L.load().orbit_pool_clips(L.ptr(f), n, clip_length, d, L.ptr(out), L.stream_ptr(f.device)), "orbit_pool_clips")
        L.count_launches(1)
        return out

    def configure(self, context_features, context_labels, ops_counter=None, clip_length=1, class_index=None):
        import ctypes as C
        import numpy as np
        L.require_cuda(context_features, "context_features")
        assert context_features.size(0) == context_labels.size(0) * clip_length, \
            "context features and labels are different sizes!"
        lib = L.load()
        feats = self._pool(context_features, clip_length)
        dev = feats.device
        ci = ClassIndex.of(context_labels, class_index, dev)
        idx = ci.index
        c, (n, d) = ci.num_classes, feats.shape
        order = np.argsort(idx, kind='stable').astype(np.int32)
        counts = np.bincount(idx, minlength=c).astype(np.int32)
        order_dev = torch.from_numpy(order).to(dev)
        mats = torch.empty(c + 1, d, d, dtype=torch.float32, device=dev)     # class precisions, then the task precision
        means = torch.empty(c, d, dtype=torch.float32, device=dev)
        task_mean = torch.empty(d, dtype=torch.float32, device=dev)
        ws = torch.empty(lib.orbit_mahalanobis_configure_workspace_bytes(n, d, c), dtype=torch.uint8, device=dev)
        L.check(lib.orbit_mahalanobis_configure(L.ptr(feats), L.ptr(order_dev), counts.ctypes.data_as(C.c_void_p), n, d, c,
                                                L.ptr(means), L.ptr(mats), L.ptr(task_mean), L.ptr(mats[c]), L.ptr(ws),
                                                L.stream_ptr(dev)), "orbit_mahalanobis_configure")
        L.count_launches(2 * d + 8 * (c + 1))
        self.means = nn.Parameter(means)
        self.task_mean = nn.Parameter(task_mean)
        self.precisions = nn.Parameter(mats[:c])
        self.task_precision = nn.Parameter(mats[c])
        if ops_counter:   # classifier_heads.py:296,308,314-320,363-366
            def cov_macs(rows):
                return rows * d + rows * rows * d + rows * d
            ops_counter.add_macs(cov_macs(n))
            for n_c in counts.tolist():
                ops_counter.add_macs(cov_macs(n_c))
                ops_counter.add_macs(n + n_c * d + 1 + 2 * d * d + (d ** 3) / 3 + d ** 2 - 4 * d / 3)

    def predict(self, target_features, ops_counter=None, clip_length=1, want_argmax=False):
        if self.means is None or self.precisions is None:
            raise AttributeError("Means and/or precisions not set - is model personalised?")
        if target_features.requires_grad and torch.is_grad_enabled() and not want_argmax:
            from .training import MahalanobisPredictFn      # meta-training (Simple CNAPs): logits with a graph to the query features
            return MahalanobisPredictFn.apply(self, target_features, clip_length)
        logits = self._predict_raw(target_features, clip_length)
        if want_argmax:
            return logits, logits.argmax(dim=1).int()
        return logits

    def _predict_raw(self, target_features, clip_length):
        L.require_cuda(target_features, "target_features")
        lib = L.load()
        q = self._pool(target_features.detach(), clip_length)
        nq, d = q.shape
        c = self.means.shape[0]
        logits = torch.empty(nq, c, dtype=torch.float32, device=q.device)
        ws = torch.empty(max(1, lib.orbit_mahalanobis_predict_workspace_bytes(nq, d)), dtype=torch.uint8, device=q.device)
        L.check(lib.orbit_mahalanobis_predict(L.ptr(q), nq, d, L.ptr(self.means.detach()), L.ptr(self.precisions.detach()), c,
                                              float(self.logit_scale), L.ptr(logits), L.ptr(ws), L.stream_ptr(q.device)),
                "orbit_mahalanobis_predict")
        L.count_launches(4 * c + 2)
        return logits
