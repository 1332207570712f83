"""Synthetic episodes of the BASELINE.json shapes: seeded fp32 NCHW frames that look like
normalised pixels (reference data/datasets.py:428-431) and class-balanced labels. Shared by tests and bench."""
from dataclasses import dataclass

import torch


@dataclass
class EpisodeSpec:
    way: int = 5
    support_clips_per_class: int = 40   # S2: 5 videos x 8 clips
    query_clips_per_class: int = 16     # S2: 2 videos x 8 clips
    clip_length: int = 8
    frame_size: int = 224


S2 = EpisodeSpec()                                     # ProtoNet + efficientnet_b0, 224, 5-way 5-shot 8 clips x 8 frames
S1 = EpisodeSpec(5, 5, 15, 1, 84)                      # config 1 shape (84x84, 1-clip)
TINY = EpisodeSpec(5, 2, 2, 2, 64)                     # unit tests


def make_episode(spec: EpisodeSpec, index: int = 0, seed: int = 1991, pin: bool = False):
    """Returns (context_clips [Ns,L,3,H,W], context_labels [Ns] int64, target_clips [Nq,L,3,H,W],
    target_labels [Nq]). Seed = 1991 + episode index (reference default seed, utils/args.py:99)."""
    g = torch.Generator().manual_seed(seed + index)
    ns, nq = spec.way * spec.support_clips_per_class, spec.way * spec.query_clips_per_class
    shape = (spec.clip_length, 3, spec.frame_size, spec.frame_size)
    ctx = torch.empty((ns,) + shape, dtype=torch.float32, pin_memory=pin)
    tgt = torch.empty((nq,) + shape, dtype=torch.float32, pin_memory=pin)
    ctx.normal_(generator=g)
    tgt.normal_(generator=g)
    ctx_labels = torch.arange(spec.way).repeat_interleave(spec.support_clips_per_class)
    tgt_labels = torch.arange(spec.way).repeat_interleave(spec.query_clips_per_class)
    ctx_labels = ctx_labels[torch.randperm(ns, generator=g)]
    tgt_labels = tgt_labels[torch.randperm(nq, generator=g)]
    # give every class a distinct low-frequency signature so that prototypes separate (random-init nets
    # otherwise map i.i.d. noise frames to nearly identical features and arg-max ties are meaningless)
    sig = torch.randn(spec.way, 3, 1, 1, generator=g) * 0.75
    ramp = torch.linspace(-1, 1, spec.frame_size)
    pat = torch.stack([torch.sin((c + 1) * 1.7 * ramp)[None, :] * torch.cos((c + 1) * 1.1 * ramp)[:, None]
                       for c in range(spec.way)])[:, None]          # [way,1,H,W]
    ctx += (sig[ctx_labels] + pat[ctx_labels])[:, None]
    tgt += (sig[tgt_labels] + pat[tgt_labels])[:, None]
    return ctx, ctx_labels, tgt, tgt_labels
