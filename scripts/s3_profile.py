"""Dev helper: where does a CNAPs + resnet18 (config 3) episode spend its time? Per kernel family, per engine."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import orbit_b200
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode
from orbit_b200.feature_extractors import get_film_parameters
dev = torch.device('cuda:0')
m = orbit_b200.SingleStepFewShotRecogniser('resnet18', True, 'versa', 1, 256, False, 16)
m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
fe = m.feature_extractor
fe.reset_parameters(1991); fe.calibrate_batchnorm(calibration_frames(224).to(dev))
m.film_generator.initial_film_parameters = get_film_parameters(m.film_parameter_names, fe)
c, cy, t, ty = [x.to(dev) for x in make_episode(EpisodeSpec(12, 10, 20, 1, 224), index=3)]
def step():
    m.personalise(c, cy); lg = m.predict(t); m._reset(); return lg
for _ in range(2): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize(); print(f"episode (120 support + 240 query frames): {(time.perf_counter()-t0)/5*1e3:.1f} ms")
engines = {'extractor': fe}
se = getattr(m, 'set_encoder', None)
for name in ('encoder', 'engine', '_engine_module', 'pre_pooling_fn'):
    if se is not None and hasattr(se, name): print('set encoder attr', name, type(getattr(se, name)))
for k, e in list(engines.items()):
    e.set_option('profile', 1)
torch.cuda.synchronize(); step(); torch.cuda.synchronize()
for k, e in engines.items():
    prof = e.profile_read(); e.set_option('profile', 0)
    tot = sum(p['ms'] for p in prof.values())
    print(k, f"total {tot:.2f} ms", {f: (round(p['ms'], 2), p['launches'], round(p['flops'] / max(p['ms'], 1e-9) / 1e9, 1)) for f, p in prof.items() if p['launches']})
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
