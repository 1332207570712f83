"""Dev helper: kernel-time table (torch.profiler, CUPTI) of one CNAPs + LITE meta-training step at 224 px."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
import orbit_b200
from orbit_b200.synthetic import EpisodeSpec, load_synthetic_checkpoint, make_episode
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda:0')
H = 16
m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', True, 'versa', 1, 256, False, H, 1.0)
m._set_device(dev); m._send_to_device(); load_synthetic_checkpoint(m, 224); m.set_test_mode(False)
ctx, ctx_y, tgt, tgt_y = make_episode(EpisodeSpec(5, 40, 16, 1, 224), index=0, pin=True)
cyd, tyd = ctx_y.to(dev), tgt_y.to(dev)
def step():
    m._clear_caches(); np.random.seed(0)
    m.personalise_with_lite(ctx, cyd)
    loss = len(cyd) / (H * 16) * F.cross_entropy(m.predict_a_batch(tgt), tyd) + 0.001 * m.film_generator.regularization_term()
    loss.backward(); m._reset()
    for p in m.parameters(): p.grad = None
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
