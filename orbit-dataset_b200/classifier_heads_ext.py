"""Versa head (reference model/classifier_heads.py:121-180, model/mlps.py:33-50) on the native kernels.
The Mahalanobis head (classifier_heads.py:265-368) needs per-class DxD inverses and is not implemented yet."""
import torch
import torch.nn as nn

from . import lib as L
from .classifier_heads import HeadClassifier, _class_index, _head_predict


class DenseResidualBlock(nn.Module):
    """mlps.py:33-50: linear1-ELU-linear2-ELU-linear3 (+ identity skip when in_size == out_size)."""

    def __init__(self, in_size, out_size):
        super().__init__()
        self.linear1 = nn.Linear(in_size, out_size)
        self.linear2 = nn.Linear(out_size, out_size)
        self.linear3 = nn.Linear(out_size, out_size)
        self.in_size, self.out_size = in_size, out_size

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

    def configure(self, context_features, context_labels, ops_counter=None, clip_length=1):
        L.require_cuda(context_features, "context_features")
        assert context_features.size(0) == context_labels.size(0) * clip_length, \
            "context features and labels are different sizes!"
        lib = L.load()
        feats = context_features.contiguous().float()
        dev = feats.device
        classes, idx = _class_index(context_labels)
        c, d, n = len(classes), feats.shape[1], len(idx)
        idx_dev = torch.from_numpy(idx).to(dev, non_blocking=True)
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


class MahalanobisClassifier(HeadClassifier):
    def __init__(self, logit_scale: float = 1.0):
        raise NotImplementedError("the Mahalanobis head (per-class DxD inverses) is not implemented on the B200 path yet")
