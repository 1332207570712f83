"""Dev helper: GPU timeline of ONE e2e S2 episode from pinned host clips (CUPTI through torch.profiler): H2D copies, backbone passes
(first stem_kernel .. spatial_mean_kernel), head kernels, and the idle time between passes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import orbit_b200
from orbit_b200.synthetic import S2, load_synthetic_checkpoint, make_episode
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda:0')
m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 8, 256, False, 16)
m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
load_synthetic_checkpoint(m, 224)
m.feature_extractor.set_option('chunk_frames', 1600)
if os.environ.get('RAMP'):
    m.stage_ramp = tuple(int(v) for v in os.environ['RAMP'].split(','))
c, cy, t, ty = make_episode(S2, index=0, pin=True)
cyd = cy.to(dev)
def step():
    m.personalise(c, cyd); lg = m.predict(t); m._reset(); return lg.cpu()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
copies = [(e.time_range.start - t0, e.time_range.end - t0) for e in ev if 'Memcpy HtoD' in e.name]
kern = [(e.name, e.time_range.start - t0, e.time_range.end - t0) for e in ev if 'Memcpy' not in e.name and 'Memset' not in e.name]
print(f"H2D: {len(copies)} copies, first starts {copies[0][0] / 1e3:.2f} ms, last ends {copies[-1][1] / 1e3:.2f} ms, busy {sum(b - a for a, b in copies) / 1e3:.2f} ms")
passes, cur = [], None
for name, a, b in kern:
    if 'stem_kernel' in name: cur = [a, b, 0.0, 0]
    if cur is not None:
        cur[1] = b; cur[2] += b - a; cur[3] += 1
        if 'spatial_mean' in name: passes.append(cur); cur = None
prev_end = 0.0
for i, (a, b, busy, n) in enumerate(passes):
    landed = max((ce for cs, ce in copies if ce <= a), default=0.0)
    print(f"pass {i}: {a / 1e3:7.2f} -> {b / 1e3:7.2f} ms ({(b - a) / 1e3:5.2f} ms, kernels {busy / 1e3:5.2f} ms, {n} launches); idle before it {(a - prev_end) / 1e3:5.2f} ms; last copy finished before its start at {landed / 1e3:7.2f} ms")
    prev_end = b
print(f"last kernel ends {kern[-1][2] / 1e3:.2f} ms; kernels after the last pass: {[k[0][:40] for k in kern if k[1] > passes[-1][1]]}")
