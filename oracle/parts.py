"""Heads, pooler, set encoder and FiLM generator (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Functional CPU restatements of the reference's ``model/{poolers,classifier_heads,set_encoders,
feature_adapters,mlps}.py``.  Pinned against the real reference modules by
``oracle/make_golden.py`` -> ``tests/golden/parts_*.npz``.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------ pooler (poolers.py:13-16)
def pool_clips(frame_feats: torch.Tensor, clip_length: int) -> torch.Tensor:
    """[N*L, D] -> [N, D]: a clip feature is the mean of its L frame features."""
    d = frame_feats.shape[-1]
    return frame_feats.reshape(-1, clip_length, d).mean(dim=1)


# ------------------------------------------------------------------ class means (classifier_heads.py:94-105)
def class_means(feats: torch.Tensor, labels: torch.Tensor):
    """Per-class mean; classes are the SORTED DISTINCT label values, column j <-> j-th smallest
    label (torch.unique, classifier_heads.py:96; sort at :246-248)."""
    assert feats.shape[0] == labels.shape[0], "context features and labels are different sizes!"
    classes = torch.unique(labels)  # sorted
    mus = [feats[labels == c].mean(dim=0) for c in classes]
    return torch.stack(mus), classes


# ------------------------------------------------------------------ prototypical head
def proto_configure(feats, labels, metric='euclidean'):
    """classifier_heads.py:232-263: W = 2*mu  [C,D];  b = -mu.mu  [C] (euclidean only)."""
    mu, _ = class_means(feats, labels)
    w = 2.0 * mu
    b = -(mu * mu).sum(dim=1) if metric == 'euclidean' else None
    return w, b


def proto_predict(q, w, b, logit_scale=1.0, metric='euclidean'):
    """classifier_heads.py:202-230.  euclidean: s*(q W^T + b) == s*(2 q.mu - mu.mu) (NOT -||q-mu||^2,
    SURVEY F7).  cosine: s*cos(q, W) with torch's eps=1e-8 clamp on each norm."""
    if metric == 'euclidean':
        return logit_scale * (q @ w.t() + b)
    if metric == 'cosine':
        qn = q.norm(dim=1, keepdim=True).clamp_min(1e-8)
        wn = w.norm(dim=1, keepdim=True).clamp_min(1e-8)
        return logit_scale * ((q / qn) @ (w / wn).t())
    raise ValueError(f"Distance function {metric} not valid.")


# ------------------------------------------------------------------ linear head (classifier_heads.py:38-79)
def linear_predict(x, w, b, logit_scale=1.0):
    return logit_scale * (x @ w.t() + b)


# ------------------------------------------------------------------ versa head (classifier_heads.py:121-180, mlps.py:33-50)
def dense_residual_block(x, p, prefix):
    """3-layer ELU MLP with identity skip when in/out widths match (mlps.py:41-50)."""
    h = F.elu(F.linear(x, p[prefix + 'linear1.weight'], p[prefix + 'linear1.bias']))
    h = F.elu(F.linear(h, p[prefix + 'linear2.weight'], p[prefix + 'linear2.bias']))
    h = F.linear(h, p[prefix + 'linear3.weight'], p[prefix + 'linear3.bias'])
    return h + x if h.shape[-1] == x.shape[-1] else h


def versa_configure(feats, labels, p):
    """Class means -> two hyper-nets -> weight rows [C,D] and biases [C]."""
    mu, _ = class_means(feats, labels)
    w = dense_residual_block(mu, p, 'weight_processor.')
    b = dense_residual_block(mu, p, 'bias_processor.').reshape(-1)
    return w, b


def init_versa_params(d, seed=7):
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, (i, o) in {'weight_processor.': (d, d), 'bias_processor.': (d, 1)}.items():
        for k, (fi, fo) in enumerate(((i, o), (o, o), (o, o)), 1):
            p[f'{name}linear{k}.weight'] = torch.randn(fo, fi, generator=g) * (fi ** -0.5) * 0.5
            p[f'{name}linear{k}.bias'] = torch.randn(fo, generator=g) * 0.02
    return p


# ------------------------------------------------------------------ mahalanobis head (classifier_heads.py:265-368)
def _cov(x):
    """classifier_heads.py:349-368: unbiased covariance of rows; single-row special case."""
    if x.shape[0] > 1:
        xc = x - x.mean(dim=0, keepdim=True)
        return xc.t() @ xc / (x.shape[0] - 1)
    xc = x - x.mean(dim=1, keepdim=True)
    return (xc @ xc.t()).squeeze() / (x.shape[1] - 1)


def mahalanobis_configure(feats, labels):
    d = feats.shape[1]
    eye = torch.eye(d, dtype=feats.dtype)
    task_cov = _cov(feats)
    means, precs = [], []
    for c in torch.unique(labels):
        xc = feats[labels == c]
        n = xc.shape[0]
        lam = n / (n + 1)
        means.append(xc.mean(dim=0))
        precs.append(torch.inverse(lam * _cov(xc) + (1 - lam) * task_cov + eye))
    return torch.stack(means), torch.stack(precs)


def mahalanobis_predict(q, means, precs, logit_scale=1.0):
    diff = means[:, None, :] - q[None, :, :]                 # [C, Nq, D]
    return -logit_scale * torch.einsum('cnd,cde,cne->nc', diff, precs, diff)


# ------------------------------------------------------------------ set encoder (set_encoders.py:34-120)
def set_encoder_forward(x, p, eps=1e-5):
    """5 x [conv3x3(pad 1) + BN(eval) + ReLU + maxpool2] + global avg pool -> [frames, 64]."""
    if x.dim() == 5:
        x = x.flatten(end_dim=1)
    for i in range(1, 6):
        pre = f'encoder.layer{i}.'
        x = F.conv2d(x, p[pre + '0.weight'], p[pre + '0.bias'], 1, 1)
        x = F.batch_norm(x, p[pre + '1.running_mean'], p[pre + '1.running_var'],
                         p[pre + '1.weight'], p[pre + '1.bias'], False, 0.0, eps)
        x = F.max_pool2d(F.relu(x), 2, 2)
    return x.mean((2, 3))


def init_set_encoder_params(seed=11, calib_input=None):
    """Seeded weights + BN stats calibrated on ``calib_input`` [n,3,H,W] (activations O(1))."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    x = calib_input if calib_input is not None else torch.randn(16, 3, 84, 84, generator=g)
    x = x[:16]
    cin = 3
    for i in range(1, 6):
        pre = f'encoder.layer{i}.'
        p[pre + '0.weight'] = torch.randn(64, cin, 3, 3, generator=g) * math.sqrt(2.0 / (9 * cin))
        p[pre + '0.bias'] = torch.randn(64, generator=g) * 0.05
        p[pre + '1.weight'] = 1.0 + 0.1 * torch.randn(64, generator=g)
        p[pre + '1.bias'] = 0.1 * torch.randn(64, generator=g)
        y = F.conv2d(x, p[pre + '0.weight'], p[pre + '0.bias'], 1, 1)
        p[pre + '1.running_mean'] = y.mean((0, 2, 3))
        p[pre + '1.running_var'] = y.var((0, 2, 3), unbiased=True)
        p[pre + '1.num_batches_tracked'] = torch.tensor(1)
        y = F.batch_norm(y, p[pre + '1.running_mean'], p[pre + '1.running_var'],
                         p[pre + '1.weight'], p[pre + '1.bias'], False, 0.0, 1e-5)
        x = F.max_pool2d(F.relu(y), 2, 2)
        cin = 64
    return p


def task_embedding(frame_reps):
    """set_encoders.py:61-75 aggregate('mean'): mean over ALL support frames -> [1,64]."""
    if not isinstance(frame_reps, torch.Tensor):
        frame_reps = torch.cat(frame_reps, dim=0)
    return frame_reps.mean(dim=0, keepdim=True)


# ------------------------------------------------------------------ FiLM generator (feature_adapters.py:36-78, mlps.py:52-63)
def film_generate(z, film_names_sorted, gen_params, initial):
    """One DenseBlock (Linear 64->64, LayerNorm, ReLU, Linear 64->size) per FiLM tensor, names in
    sorted order (feature_adapters.py:43-44):
        gamma' = gamma0 * (g(z)*r + 1)   for '...weight'   (:69-71)
        beta'  = beta0  + g(z)*r         for '...bias'     (:72-74)
    Returns (film_dict, l2_term = sum ||r||^2 (:76))."""
    out, l2 = {}, torch.zeros(())
    for i, name in enumerate(film_names_sorted):
        pre = f'generators.{i}.block.'
        h = F.linear(z, gen_params[pre + '0.weight'], gen_params[pre + '0.bias'])
        h = F.layer_norm(h, (h.shape[-1],), gen_params[pre + '1.weight'], gen_params[pre + '1.bias'], 1e-5)
        g = F.linear(F.relu(h), gen_params[pre + '3.weight'], gen_params[pre + '3.bias']).squeeze()
        r = gen_params[f'regularizers.{i}']
        if 'weight' in name:
            out[name] = initial[name] * (g * r + 1.0)
        elif 'bias' in name:
            out[name] = initial[name] + g * r
        l2 = l2 + (r ** 2).sum()
    return out, l2


def init_film_generator_params(sizes_sorted, seed=13, hidden=64, reg_std=0.05):
    """Seeded generator weights. ``reg_std`` is larger than the reference's init (1e-3,
    feature_adapters.py:50) so that the synthetic FiLM perturbation is clearly visible in parity
    tests (a trained generator is not O(1e-3) either)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for i, n in enumerate(sizes_sorted):
        pre = f'generators.{i}.block.'
        p[pre + '0.weight'] = torch.randn(hidden, hidden, generator=g) * hidden ** -0.5
        p[pre + '0.bias'] = torch.randn(hidden, generator=g) * 0.05
        p[pre + '1.weight'] = 1.0 + 0.1 * torch.randn(hidden, generator=g)
        p[pre + '1.bias'] = 0.1 * torch.randn(hidden, generator=g)
        p[pre + '3.weight'] = torch.randn(n, hidden, generator=g) * hidden ** -0.5
        p[pre + '3.bias'] = torch.randn(n, generator=g) * 0.05
        p[f'regularizers.{i}'] = torch.randn(n, generator=g) * reg_std
    return p


# ------------------------------------------------------------------ FiLM site selection (film.py:38-94)
def film_parameter_names(extractor_name: str, model: nn.Module):
    """Names ``<module path>.weight/.bias`` of the norm layers the reference tags:
    EfficientNet: InvertedResidual.bn2 + root bn1/bn2 (film.py:41-47); ViT: every LayerNorm called
    norm/norm1/norm2 (film.py:57-66); resnet18 (extension): bn1/bn2 of each BasicBlock."""
    names = []
    for mname, mod in model.named_modules():
        leaf = mname.split('.')[-1]
        parent = mname.rsplit('.', 1)[0] if '.' in mname else ''
        if 'efficientnet' in extractor_name:
            ptype = type(model.get_submodule(parent)).__name__ if parent else ''
            if isinstance(mod, nn.BatchNorm2d) and (
                    (parent == '' and leaf in ('bn1', 'bn2')) or
                    (leaf == 'bn2' and ptype == 'InvertedResidual') or
                    (leaf == 'bn1' and ptype in ('EdgeResidual', 'ConvBnAct'))):
                names += [mname + '.weight', mname + '.bias']
        elif 'vit' in extractor_name:
            if isinstance(mod, nn.LayerNorm) and leaf in ('norm', 'norm1', 'norm2'):
                names += [mname + '.weight', mname + '.bias']
        elif 'resnet' in extractor_name:
            if isinstance(mod, nn.BatchNorm2d) and parent.startswith('layer') and leaf in ('bn1', 'bn2'):
                names += [mname + '.weight', mname + '.bias']
    return names


# ------------------------------------------------------------------ helpers around the path (data/utils.py:8-28)
def attach_frame_history(frames, history_length):
    """Causal sliding window: clip t = frames[t-L+1 .. t], left-padded with frame 0."""
    if history_length == 1:
        return frames.unsqueeze(1)
    n = frames.shape[0]
    idx = (torch.arange(n)[:, None] + torch.arange(-history_length + 1, 1)[None, :]).clamp_min(0)
    return frames[idx]
