"""Synthetic ORBIT-like data of the BASELINE.json shapes. Shared by tests and bench.

Frames are seeded fp32 NCHW tensors that look like normalised pixels (reference
data/datasets.py:428-431): unit Gaussian noise plus the fixed low-frequency "appearance" of one of
16 synthetic object classes, so that prototypes separate and arg-max parity is meaningful.
Labels are class-balanced and permuted (seed = 1991 + episode index; reference default seed
utils/args.py:99)."""
from dataclasses import dataclass

import torch

NUM_OBJECTS = 16
_BANK_SEED = 1991


@dataclass
class EpisodeSpec:
    way: int = 5
    support_clips_per_class: int = 40   # S2: 5 videos x 8 clips
    query_clips_per_class: int = 16     # S2: 2 videos x 8 clips
    clip_length: int = 8
    frame_size: int = 224


S2 = EpisodeSpec()                      # ProtoNet + efficientnet_b0, 224, 5-way 5-shot, 8 clips x 8 frames
S1 = EpisodeSpec(5, 5, 15, 1, 84)       # config 1 shape (84x84, 1-clip)


def object_bank(size: int):
    """Appearance [NUM_OBJECTS,3,size,size] of the synthetic objects: colour offset + smooth pattern."""
    g = torch.Generator().manual_seed(_BANK_SEED)
    colour = torch.randn(NUM_OBJECTS, 3, 1, 1, generator=g) * 0.5
    ramp = torch.linspace(-1, 1, size)
    fx = torch.rand(NUM_OBJECTS, generator=g) * 4 + 1
    fy = torch.rand(NUM_OBJECTS, generator=g) * 4 + 1
    pat = torch.stack([torch.sin(fx[c] * ramp)[None, :] * torch.cos(fy[c] * ramp)[:, None] for c in range(NUM_OBJECTS)])
    return colour + 1.0 * pat[:, None]


def calibration_frames(size: int, per_object: int = None, seed: int = 7):
    """Frames of the episode distribution (every object equally often) used to calibrate the synthetic
    weights' BatchNorm statistics."""
    if per_object is None:
        fs = -(-size // 32)
        per_object = max(2, -(-64 // (fs * fs)))
    g = torch.Generator().manual_seed(seed)
    n = per_object * NUM_OBJECTS
    x = torch.randn(n, 3, size, size, generator=g)
    return x + object_bank(size)[torch.arange(n) % NUM_OBJECTS]


def make_episode(spec: EpisodeSpec, index: int = 0, seed: int = 1991, pin: bool = False):
    """Returns (context_clips [Ns,L,3,H,W], context_labels [Ns] int64, target_clips [Nq,L,3,H,W],
    target_labels [Nq])."""
    g = torch.Generator().manual_seed(seed + index)
    ns, nq = spec.way * spec.support_clips_per_class, spec.way * spec.query_clips_per_class
    shape = (spec.clip_length, 3, spec.frame_size, spec.frame_size)
    ctx = torch.empty((ns,) + shape, dtype=torch.float32, pin_memory=pin)
    tgt = torch.empty((nq,) + shape, dtype=torch.float32, pin_memory=pin)
    ctx.normal_(generator=g)
    tgt.normal_(generator=g)
    objects = torch.randperm(NUM_OBJECTS, generator=g)[:spec.way]      # which objects this task is about
    ctx_labels = torch.arange(spec.way).repeat_interleave(spec.support_clips_per_class)
    tgt_labels = torch.arange(spec.way).repeat_interleave(spec.query_clips_per_class)
    ctx_labels = ctx_labels[torch.randperm(ns, generator=g)]
    tgt_labels = tgt_labels[torch.randperm(nq, generator=g)]
    bank = object_bank(spec.frame_size)[objects]
    ctx += bank[ctx_labels][:, None]
    tgt += bank[tgt_labels][:, None]
    return ctx, ctx_labels, tgt, tgt_labels


def make_episode_on_device(spec: EpisodeSpec, index: int, device, seed: int = 1991):
    """Same construction as ``make_episode`` with the frame noise drawn by the DEVICE generator (Philox, seeded by the
    episode index only): a fixed list of episodes can be materialised on whichever GPU an episode is dealt to, and is
    bit-identical there -- what the sharded evaluation's exactness check needs. Labels / object choice
    come from the CPU generator (a few hundred integers)."""
    g = torch.Generator(device=device).manual_seed(seed + index)
    cg = torch.Generator().manual_seed(seed + index)
    ns, nq = spec.way * spec.support_clips_per_class, spec.way * spec.query_clips_per_class
    shape = (spec.clip_length, 3, spec.frame_size, spec.frame_size)
    ctx = torch.randn((ns,) + shape, generator=g, device=device)
    tgt = torch.randn((nq,) + shape, generator=g, device=device)
    objects = torch.randperm(NUM_OBJECTS, generator=cg)[:spec.way]
    ctx_labels = torch.arange(spec.way).repeat_interleave(spec.support_clips_per_class)[torch.randperm(ns, generator=cg)]
    tgt_labels = torch.arange(spec.way).repeat_interleave(spec.query_clips_per_class)[torch.randperm(nq, generator=cg)]
    bank = object_bank(spec.frame_size)[objects].to(device)
    ctx_labels, tgt_labels = ctx_labels.to(device), tgt_labels.to(device)
    ctx += bank[ctx_labels][:, None]
    tgt += bank[tgt_labels][:, None]
    return ctx, ctx_labels, tgt, tgt_labels


def load_synthetic_checkpoint(model, frame_size: int = 224, seed: int = 1991):
    """Gives a recogniser that is already on its CUDA device a synthetic 'pretrained' extractor: seeded
    isometric weights (FeatureExtractor.reset_parameters) + BatchNorm statistics calibrated on device over
    ``calibration_frames(frame_size)``. Deterministic, so every rank of a multi-GPU run gets the same weights."""
    fe = model.feature_extractor
    fe.reset_parameters(seed)          # CPU RNG, written into the (device-resident) parameter blob
    fe.calibrate_batchnorm(calibration_frames(frame_size).to(fe._blob.device))
    if getattr(model, 'adapt_features', False) and hasattr(model, 'film_generator'):
        from .feature_extractors import get_film_parameters
        model.film_generator.initial_film_parameters = get_film_parameters(model.film_parameter_names, fe)
    return model
