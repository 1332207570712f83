"""Dev helper: torch.profiler (CUPTI) kernel table of one bench.py episode of a configuration (CFG = S1 | S2 | S3 | S4)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from orbit_b200.synthetic import make_episode
name = os.environ.get('CFG', 'S4')
from orbit_b200 import lib as L
for kv in filter(None, os.environ.get('OPTS', '').split(',')):
    k, v = kv.split('=')
    assert L.load().orbit_set_global_option(k.encode(), int(v)) == 0, kv
dev = torch.device('cuda:0')
model = bench.build_model(name, dev, 1, 1600)
spec = bench.config_spec(name)
c, cy, t, ty = make_episode(spec, index=0)
c, t = c.to(dev), t.to(dev)
cy = cy if name == 'S4' else cy.to(dev)
for _ in range(3): bench.run_episode(name, model, c, cy, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): bench.run_episode(name, model, c, cy, t)
e1.record(); torch.cuda.synchronize()
print(f"{name}: {e0.elapsed_time(e1) / 3:.2f} ms per episode")
from torch.profiler import profile, ProfilerActivity
if os.environ.get('NOPROF'): sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    bench.run_episode(name, model, c, cy, t); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=24, max_name_column_width=90))
