"""Device-side FineTuner loop (reference few_shot_recognisers.py:231-246, utils/optim.py:8-32)."""
import torch

from . import lib as L
from .classifier_heads import _class_index


def finetune_linear_head(classifier, clip_features, context_labels, batch_size, num_grad_steps, learning_rate,
                         optimizer, optimizer_kwargs, logit_scale):
    """Runs all ``num_grad_steps`` optimiser steps of the zero-initialised linear head in one kernel.
    ``batch_size`` only changes the fp32 summation order of the reference's gradient accumulation (each batch
    loss is re-weighted by batch_len/N), so it does not enter the computation.
    ``optimizer_kwargs``: the keys the reference reads with getattr (utils/optim.py:16-26)."""
    L.require_cuda(clip_features, "features")
    lib = L.load()
    feats = clip_features.contiguous().float()
    n, d = feats.shape
    classes, idx = _class_index(context_labels)
    c = len(classes)
    if classifier.weight is None or classifier.weight.shape != (c, d):
        raise ValueError("classifier.init(num_classes) must be called with the number of distinct labels")
    if optimizer not in ('adam', 'sgd'):
        raise ValueError(f"optimizer {optimizer} not valid")
    betas = optimizer_kwargs.get('betas', (0.9, 0.999))
    dev = feats.device
    idx_dev = torch.from_numpy(idx).to(dev, non_blocking=True)
    scratch = torch.empty(lib.orbit_linear_finetune_scratch_bytes(n, d, c), dtype=torch.uint8, device=dev)
    w, b = classifier.weight.data, classifier.bias.data
    L.check(lib.orbit_linear_finetune(L.ptr(feats), L.ptr(idx_dev), n, d, c, int(num_grad_steps),
                                      0 if optimizer == 'adam' else 1, float(learning_rate), float(betas[0]),
                                      float(betas[1]), float(optimizer_kwargs.get('epsilon', 1e-8)),
                                      float(optimizer_kwargs.get('weight_decay', 0.0)),
                                      float(optimizer_kwargs.get('momentum', 0.0)), float(logit_scale),
                                      L.ptr(w), L.ptr(b), L.ptr(scratch), L.stream_ptr(dev)), "orbit_linear_finetune")
    L.count_launches(1)
