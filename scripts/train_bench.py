"""Dev helper: one CNAPs meta-training step (reference single-step-learner.py:196-243) on a B200, frozen EfficientNet-B0.
LITE step = personalise_with_lite + predict_a_batch + loss.backward(); plain step = personalise + predict + loss.backward()."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import orbit_b200
from orbit_b200.synthetic import EpisodeSpec, load_synthetic_checkpoint, make_episode

dev = torch.device('cuda:0')
way, shot, query, L, size, H = 5, 40, 16, 1, 224, 16          # 200 context clips, 80 target clips, clip_length 1 (reference default)
m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', True, 'versa', L, 256, False, H, 1.0)
m._set_device(dev); m._send_to_device()
load_synthetic_checkpoint(m, size)
m.set_test_mode(False)
ctx, ctx_y, tgt, tgt_y = make_episode(EpisodeSpec(way, shot, query, L, size), index=0, pin=True)
ctx_d, tgt_d, ctx_yd, tgt_yd = ctx.to(dev), tgt.to(dev), ctx_y.to(dev), tgt_y.to(dev)
params = list(m.set_encoder.parameters()) + list(m.film_generator.parameters())
opt = torch.optim.Adam(params, lr=1e-4)

def lite_step(c, t):
    m._clear_caches()
    np.random.seed(0)
    m.personalise_with_lite(c, ctx_yd)
    logits = m.predict_a_batch(t)
    loss = len(ctx_yd) / (H * 16) * F.cross_entropy(logits, tgt_yd) + 0.001 * m.film_generator.regularization_term()
    loss.backward()
    m._reset()
    return loss

def plain_step(c, t):
    m.personalise(c, ctx_yd)
    logits = m.predict(t)
    loss = F.cross_entropy(logits, tgt_yd) / 16 + 0.001 * m.film_generator.regularization_term()
    loss.backward()
    m._reset()
    return loss

def test_step(c, t):
    with torch.no_grad():
        m.set_test_mode(True)
        m.personalise(c, ctx_yd); lg = m.predict(t); m._reset()
        m.set_test_mode(False)
    return lg

for name, fn in (('inference episode (no graph)', test_step), ('LITE training step (H = 16)', lite_step), ('plain training step (all 200 context clips with grad)', plain_step)):
    for where, c, t in (('device clips', ctx_d, tgt_d), ('pinned host clips', ctx, tgt)):
        for _ in range(2):
            fn(c, t); opt.zero_grad()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            out = fn(c, t); opt.step(); opt.zero_grad()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / n * 1e3
        print(f"{name:55s} {where:18s} {ms:8.1f} ms  (peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)", flush=True)
