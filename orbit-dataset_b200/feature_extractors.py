"""Feature extractors behind the reference's ``create_feature_extractor`` interface.

Mirrors reference ``model/feature_extractors.py:37-87`` and ``model/film.py:68-94``: same extractor
strings, same (extractor, film_parameter_names) return value, same ``output_size`` attribute, the
same timm state-dict key names.  The module only HOLDS the parameters (as views into one flat fp32
blob that the CUDA engine reads directly); the forward pass is liborbit_b200's native layer plan.
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from . import lib as L

ARCH_IDS = {
    'set_encoder': 100,   # not an extractor string of the reference: the SetEncoder runs through the same engine
    'efficientnet_b0': 0,
    'vit_s_32': 1,
    'vit_b_32': 2,
    'vit_b_32_clip': 3,
    'resnet18': 4,  # BASELINE.json extension; not in the reference at this commit
    'efficientnet_v2_s': 5,
}


class _Node(nn.Module):
    """Anonymous container used to reproduce timm's dotted parameter paths."""


def _engine_table(engine, count_fn, info_fn):
    lib = L.load()
    out = []
    name = C.create_string_buffer(256)
    numel, offset = C.c_int64(), C.c_int64()
    for i in range(count_fn(engine)):
        L.check(info_fn(engine, i, name, 256, C.byref(numel), C.byref(offset)), "param_info")
        out.append((name.value.decode(), numel.value, offset.value))
    return out


class FeatureExtractor(nn.Module):
    """Parameter holder + native forward. ``forward(frames[B,3,H,W]) -> [B, output_size]``."""

    def __init__(self, name: str, seed: int = 0):
        super().__init__()
        if name not in ARCH_IDS:
            raise ValueError(f"Invalid feature_extractor_name: {name}")
        lib = L.load()
        self.extractor_name = name
        handle = C.c_void_p()
        rc = lib.orbit_engine_create(C.byref(handle), ARCH_IDS[name])
        if rc == -2:
            raise NotImplementedError(f"feature extractor '{name}' is not implemented by liborbit_b200 yet")
        L.check(rc, "orbit_engine_create")
        self._engine = handle
        self.output_size = lib.orbit_engine_feat_dim(handle)
        self._table = _engine_table(handle, lib.orbit_engine_num_params, lib.orbit_engine_param_info)
        self._film_table = _engine_table(handle, lib.orbit_engine_num_film, lib.orbit_engine_film_info)
        self._n_floats = lib.orbit_engine_param_floats(handle)
        self._blob = torch.zeros(self._n_floats, dtype=torch.float32)
        self._shapes = {}
        self._derived = None
        self._workspace = None
        self._prepared_key = None
        self._versioned = None
        self._train_derived = self._saved = self._grad_blob = self._train_state = None
        self._build_tree()
        self.reset_parameters(seed)

    # ---- parameter tree ------------------------------------------------------------------------
    def _build_tree(self):
        """Reproduces the reference/timm module tree (dotted names) with parameters that are views of the blob."""
        lib = L.load()
        ndim, dims = C.c_int(), (C.c_int64 * 4)()
        for i, (name, numel, offset) in enumerate(self._table):
            L.check(lib.orbit_engine_param_shape(self._engine, i, C.byref(ndim), dims), "orbit_engine_param_shape")
            shape = tuple(int(dims[d]) for d in range(ndim.value))
            self._shapes[name] = shape
            parent = self
            parts = name.split('.')
            for p in parts[:-1]:
                if not hasattr(parent, p):
                    parent.add_module(p, _Node())
                parent = getattr(parent, p)
            view = self._blob[offset:offset + numel].view(shape)
            if parts[-1] in ('running_mean', 'running_var'):
                parent.register_buffer(parts[-1], view)
                if parts[-1] == 'running_var':
                    parent.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
            else:
                parent.register_parameter(parts[-1], nn.Parameter(view, requires_grad=True))

    def _rebind(self):
        params, bufs = dict(self.named_parameters()), dict(self.named_buffers())
        for name, numel, offset in self._table:
            view = self._blob[offset:offset + numel].view(self._shapes[name])
            (params[name] if name in params else bufs[name]).data = view

    def _apply(self, fn, recurse=True):
        new_blob = fn(self._blob)
        if new_blob.dtype != torch.float32:
            raise TypeError("orbit_b200 feature extractors are fp32 only")
        for mod in self.modules():  # non-blob buffers (num_batches_tracked)
            if mod is not self and 'num_batches_tracked' in mod._buffers:
                mod._buffers['num_batches_tracked'] = fn(mod._buffers['num_batches_tracked'])
        self._blob = new_blob
        self._rebind()
        self._derived = self._workspace = self._prepared_key = self._versioned = None
        self._train_derived = self._saved = self._grad_blob = self._train_state = None
        return self

    @torch.no_grad()
    def reset_parameters(self, seed=0):
        """Seeded stand-in for the reference's ``pretrained=True`` download (there is no network).
        Drawn for dynamical isometry so that a random deep net neither explodes nor collapses:
        (semi-)orthogonal pointwise / dense / stem weights, depthwise = centre tap + small noise,
        small squeeze-excite FCs, near-identity norms. Follow with ``calibrate_batchnorm`` (or load a
        real checkpoint with ``load_state_dict``)."""
        g = torch.Generator().manual_seed(seed)
        for name, numel, offset in self._table:
            dst = self._blob[offset:offset + numel]
            leaf = name.rsplit('.', 1)[-1]
            shape = self._shapes[name]
            if leaf == 'running_var':
                dst.fill_(1.0)
            elif leaf == 'running_mean':
                dst.zero_()
            elif leaf == 'weight' and len(shape) == 1:
                dst.copy_(1.0 + 0.1 * torch.randn(numel, generator=g))
            elif leaf == 'bias':
                dst.copy_((0.1 if name.rsplit('.', 2)[-2].startswith(('bn', 'norm')) else 0.05) *
                          torch.randn(numel, generator=g))
            elif len(shape) == 4 and shape[1] != 1 and shape[2] == 3 and shape[0] > shape[1] * 9:
                # first 3x3 conv of the set encoder (64 filters over 27 inputs): orthonormal columns
                flat = torch.empty(shape[0], numel // shape[0])
                nn.init.orthogonal_(flat, generator=g)
                dst.copy_(flat.flatten())
            elif len(shape) == 4 and shape[1] == 1:            # depthwise
                k = shape[2]
                w = torch.randn(shape, generator=g) * (0.2 / k)
                w[:, 0, k // 2, k // 2] += 1.0
                dst.copy_(w.flatten())
            elif '.se.' in name:
                dst.copy_(torch.randn(numel, generator=g) * (0.5 * shape[1] ** -0.5))
            else:
                flat = torch.empty(shape[0], numel // shape[0])
                nn.init.orthogonal_(flat, generator=g)
                dst.copy_(flat.flatten())

    @torch.no_grad()
    def calibrate_batchnorm(self, frames: torch.Tensor) -> torch.Tensor:
        """Sets every BatchNorm's running statistics to the batch statistics of ``frames`` [B,3,H,W]
        (one native pass in which each layer normalises with this batch, i.e. train-mode BatchNorm with
        momentum 1). Returns the features of that pass."""
        lib = L.load()
        L.require_cuda(frames, "frames")
        frames = frames.contiguous().float()
        n, _, h, w = frames.shape
        old_chunk = self.get_option('chunk_frames')
        self.set_option('chunk_frames', max(n, 2))
        try:
            self._prepared_key = None
            self.prepare(None)
            ws_bytes = lib.orbit_engine_workspace_bytes(self._engine, h, w)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=frames.device)
            feats = torch.empty(n, self.output_size, dtype=torch.float32, device=frames.device)
            L.check(lib.orbit_engine_calibrate(self._engine, L.ptr(self._blob), L.ptr(self._derived), L.ptr(frames), n, h, w,
                                               L.ptr(feats), L.ptr(ws), ws.numel(), L.stream_ptr(frames.device)),
                    "orbit_engine_calibrate")
            L.count_launches(lib.orbit_engine_last_launches(self._engine))
            for mod in self.modules():
                if 'num_batches_tracked' in mod._buffers:
                    mod._buffers['num_batches_tracked'].fill_(1)
        finally:
            self.set_option('chunk_frames', old_chunk)
            self._prepared_key = None
        return feats

    # ---- FiLM ----------------------------------------------------------------------------------
    def film_parameter_names(self):
        """Names in module-registration order, as get_film_parameter_names walks named_modules()
        (film.py:68-74); weight before bias per layer."""
        tagged = {n for n, _, _ in self._film_table}
        return [n for n, _, _ in self._table if n in tagged]

    def film_layout(self):
        """[(name, numel, offset)] of the film blob, in the generator's sorted order."""
        return list(self._film_table)

    # ---- native forward ------------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        L.check(L.load().orbit_engine_set_option(self._engine, key.encode(), int(value)), f"set_option({key})")
        self._workspace = None

    def get_option(self, key: str) -> int:
        v = C.c_int()
        L.check(L.load().orbit_engine_get_option(self._engine, key.encode(), C.byref(v)), f"get_option({key})")
        return v.value

    def count_macs(self, frames) -> int:
        """MACs of one forward over ``frames`` [..., 3, H, W] (any leading dims); see ``ops_counter.py``."""
        h, w = int(frames.shape[-2]), int(frames.shape[-1])
        n = int(frames.numel() // (3 * h * w))
        per_frame = L.load().orbit_engine_macs(self._engine, h, w)
        if per_frame < 0:
            L.check(int(per_frame), "orbit_engine_macs")
        return n * int(per_frame)

    PROFILE_FAMILIES = ('stem_conv', 'depthwise_conv', 'se_gate', 'pointwise_gemm', 'spatial_mean', 'calibration')

    def profile_read(self):
        """{family: dict(ms, launches, bytes, flops)} of the launches timed since the last read
        (needs set_option('profile', 1)); blocks until they have finished."""
        n = len(self.PROFILE_FAMILIES)
        ms, la, by, fl = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_double * n)(), (C.c_double * n)()
        L.check(L.load().orbit_engine_profile_read(self._engine, ms, la, by, fl), "orbit_engine_profile_read")
        return {name: dict(ms=ms[i], launches=la[i], bytes=by[i], flops=fl[i])
                for i, name in enumerate(self.PROFILE_FAMILIES)}

    def prepare(self, film_blob=None):
        """Folds norms (+ FiLM gamma'/beta'), re-lays-out weights. Cheap; re-run when params/film change."""
        lib = L.load()
        L.require_cuda(self._blob, "feature extractor parameters")
        if self._derived is None:
            self._derived = torch.empty(lib.orbit_engine_derived_floats(self._engine), dtype=torch.float32,
                                        device=self._blob.device)
        # Parameters are views of the blob bound with `param.data = view`, which keeps each parameter's OWN version
        # counter: load_state_dict / param.copy_ / optimiser steps bump those, not the blob's. The FiLM blob is written by
        # a kernel through a raw pointer (no version bump at all): the generator stamps a generation id on it.
        key = (self._state_version(), None if film_blob is None else
               (film_blob.data_ptr(), film_blob._version, getattr(film_blob, '_orbit_generation', None)))
        if key == self._prepared_key:
            return
        if film_blob is not None:
            L.require_cuda(film_blob, "film parameters")
            assert film_blob.dtype == torch.float32 and film_blob.numel() == lib.orbit_engine_film_floats(self._engine)
        L.check(lib.orbit_engine_prepare(self._engine, L.ptr(self._blob), L.ptr(film_blob), L.ptr(self._derived),
                                         L.stream_ptr(self._blob.device)), "orbit_engine_prepare")
        L.count_launches(2 + len(self._table) // 4)
        self._prepared_key = key

    def _state_version(self):
        if self._versioned is None:
            self._versioned = [t for t in list(self.parameters()) + list(self.buffers())] + [self._blob]
        return sum(t._version for t in self._versioned)

    def mark_dirty(self):
        """Call after writing parameters through an alias autograd cannot see (``param.data.copy_``, raw pointers)."""
        self._prepared_key = None

    def load_state_dict(self, *args, **kwargs):
        self._prepared_key = None
        return super().load_state_dict(*args, **kwargs)

    def forward(self, frames: torch.Tensor, film_blob=None) -> torch.Tensor:
        if film_blob is not None and film_blob.requires_grad and torch.is_grad_enabled():
            # meta-training: the loss back-propagates through the frozen extractor into the FiLM generator
            from .training import ExtractorFilmFn
            return ExtractorFilmFn.apply(self, frames, film_blob, getattr(film_blob, '_orbit_generation', None))
        lib = L.load()
        L.require_cuda(frames, "frames")
        if frames.dim() != 4 or frames.shape[1] != 3:
            raise ValueError(f"frames must be [B,3,H,W], got {tuple(frames.shape)}")
        frames = frames.contiguous().float()
        self.prepare(film_blob)
        n, _, h, w = frames.shape
        ws_bytes = lib.orbit_engine_workspace_bytes(self._engine, h, w)
        if ws_bytes < 0:
            L.check(int(ws_bytes), "orbit_engine_workspace_bytes")
        if self._workspace is None or self._workspace.numel() < ws_bytes or self._workspace.device != frames.device:
            self._workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=frames.device)
        ws = self._workspace
        feats = torch.empty(n, self.output_size, dtype=torch.float32, device=frames.device)
        L.check(lib.orbit_engine_forward(self._engine, L.ptr(self._blob), L.ptr(self._derived), L.ptr(frames), n, h, w,
                                         L.ptr(feats), L.ptr(ws), ws.numel(),
                                         L.stream_ptr(frames.device)), "orbit_engine_forward")
        L.count_launches(lib.orbit_engine_last_launches(self._engine))
        return feats

    # ---- training through the frozen extractor (FineTuner + FiLM, CNAPs meta-training) ---------------------
    def _param_slots(self):
        """{id(parameter): (blob offset, numel, shape)} for the autograd bridges of training.py"""
        params = dict(self.named_parameters())
        return {id(params[name]): (offset, numel, self._shapes[name]) for name, numel, offset in self._table if name in params}

    def _forward_train_impl(self, frames, film_blob, fresh_arena):
        """One training-mode pass (BatchNorm in eval mode, layer at a time, pre-activations kept): returns the features and
        the state the backward needs. ``fresh_arena``: allocate a new activation arena (several passes may be alive before the
        backward runs, as under autograd) instead of reusing the module's."""
        lib = L.load()
        L.require_cuda(frames, "frames")
        if frames.dim() != 4 or frames.shape[1] != 3:
            raise ValueError(f"frames must be [B,3,H,W], got {tuple(frames.shape)}")
        frames = frames.contiguous().float()
        self.prepare(film_blob)
        n, _, h, w = frames.shape
        per_frame = lib.orbit_engine_train_saved_floats(self._engine, h, w)
        if per_frame < 0:
            raise NotImplementedError(f"training through '{self.extractor_name}' needs backward kernels that do not exist yet "
                                      "(SURVEY.md 8f-3: the MBConv networks and the set encoder are covered)")
        if n > self.get_option('chunk_frames'):
            self.set_option('chunk_frames', n)
        if self._train_derived is None:
            self._train_derived = torch.empty(max(4, lib.orbit_engine_train_derived_floats(self._engine)), dtype=torch.float32,
                                              device=frames.device)
            L.check(lib.orbit_engine_prepare_train(self._engine, L.ptr(self._blob), L.ptr(self._train_derived), L.stream_ptr(frames.device)),
                    "orbit_engine_prepare_train")
        if fresh_arena:
            saved = torch.empty(per_frame * n, dtype=torch.float32, device=frames.device)
        else:
            if self._saved is None or self._saved.numel() < per_frame * n:
                self._saved = torch.empty(per_frame * n, dtype=torch.float32, device=frames.device)
            saved = self._saved
        ws_bytes = lib.orbit_engine_workspace_bytes(self._engine, h, w)
        if self._workspace is None or self._workspace.numel() < ws_bytes:
            self._workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=frames.device)
        feats = torch.empty(n, self.output_size, dtype=torch.float32, device=frames.device)
        L.check(lib.orbit_engine_forward_train(self._engine, L.ptr(self._blob), L.ptr(self._derived), L.ptr(frames), n, h, w, L.ptr(feats),
                                               L.ptr(saved), saved.numel(), L.ptr(self._workspace), self._workspace.numel(),
                                               L.stream_ptr(frames.device)), "orbit_engine_forward_train")
        L.count_launches(lib.orbit_engine_last_launches(self._engine))
        return feats, (saved, frames, (n, h, w))

    def _backward_train_impl(self, dfeats, state, film_blob, grad_blob):
        """ACCUMULATES into ``grad_blob`` (layout of the parameter blob): the FiLM-site BatchNorm weight / bias gradients of
        an MBConv extractor, every parameter's gradient of the set encoder."""
        lib = L.load()
        saved, frames, (n, h, w) = state
        dfeats = dfeats.contiguous().float()
        assert dfeats.shape == (n, self.output_size)
        self.prepare(film_blob)            # no-op unless another task's FiLM parameters were folded in between
        ws_bytes = lib.orbit_engine_workspace_bytes(self._engine, h, w)
        if self._workspace is None or self._workspace.numel() < ws_bytes:
            self._workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dfeats.device)
        L.check(lib.orbit_engine_backward_train(self._engine, L.ptr(self._blob), L.ptr(self._derived), L.ptr(self._train_derived),
                                                L.ptr(saved), L.ptr(frames), L.ptr(dfeats), n, h, w, L.ptr(grad_blob),
                                                L.ptr(self._workspace), self._workspace.numel(), L.stream_ptr(dfeats.device)),
                "orbit_engine_backward_train")
        L.count_launches(lib.orbit_engine_last_launches(self._engine))

    def forward_train(self, frames: torch.Tensor) -> torch.Tensor:
        """Forward pass that keeps the pre-activations the backward needs (BatchNorm in eval mode, layer at a time);
        same features as ``forward``. One pass: ``len(frames) <= chunk_frames``. Follow with ``backward_train``."""
        feats, self._train_state = self._forward_train_impl(frames, None, fresh_arena=False)
        return feats

    def backward_train(self, dfeats: torch.Tensor):
        """Back-propagates ``dfeats`` [B, output_size] of the last ``forward_train`` and ACCUMULATES the gradients of the
        FiLM parameters into ``param.grad`` (views of one gradient blob laid out like the parameters)."""
        if self._grad_blob is None:
            self._grad_blob = torch.zeros(self._n_floats, dtype=torch.float32, device=dfeats.device)
            tagged = {name for name, _, _ in self._film_table}
            params = dict(self.named_parameters())
            for name, numel, offset in self._table:
                if name in tagged:
                    params[name].grad = self._grad_blob[offset:offset + numel].view(self._shapes[name])
        self._backward_train_impl(dfeats, self._train_state, None, self._grad_blob)

    def zero_film_grads(self):
        if self._grad_blob is not None:
            self._grad_blob.zero_()

    def __del__(self):
        try:
            if getattr(self, '_engine', None):
                L.load().orbit_engine_destroy(self._engine)
                self._engine = None
        except Exception:
            pass


def freeze_extractor(feature_extractor):
    """feature_extractors.py:81-87."""
    for param in feature_extractor.parameters():
        param.requires_grad = False


def create_feature_extractor(feature_extractor_name: str, pretrained: bool, with_film: bool = False,
                             learn_extractor: bool = True):
    """Same contract as reference feature_extractors.py:37-79. ``pretrained=True`` cannot download
    (no network): weights come from $ORBIT_B200_PRETRAINED/<name>.pth when that file exists, else a
    seeded initialisation; callers normally ``load_state_dict`` a checkpoint afterwards."""
    feature_extractor = FeatureExtractor(feature_extractor_name)
    root = os.environ.get('ORBIT_B200_PRETRAINED')
    if pretrained and root and os.path.exists(os.path.join(root, feature_extractor_name + '.pth')):
        feature_extractor.load_state_dict(torch.load(os.path.join(root, feature_extractor_name + '.pth')))
    if not learn_extractor:
        freeze_extractor(feature_extractor)
    film_param_names = feature_extractor.film_parameter_names() if with_film else None
    return feature_extractor, film_param_names


# ---- reference model/film.py helpers, same names/semantics -----------------------------------------
def unfreeze_film(film_parameter_names, feature_extractor):
    for name, param in feature_extractor.named_parameters():
        if name in film_parameter_names:
            param.requires_grad = True


def get_film_parameters(film_parameter_names, feature_extractor):
    film_params = {}
    if film_parameter_names is not None:
        for name, param in feature_extractor.named_parameters():
            if name in film_parameter_names:
                film_params[name] = param.detach().clone()
    return film_params


def get_film_parameter_sizes(film_parameter_names, feature_extractor):
    return {name: len(param) for name, param in feature_extractor.named_parameters() if name in film_parameter_names}
