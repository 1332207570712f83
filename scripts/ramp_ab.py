"""Dev helper: e2e S2 episode time for several host-clip ramps, pinned and pageable sources, on one box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, orbit_b200
from orbit_b200.synthetic import S2, load_synthetic_checkpoint, make_episode
dev = torch.device('cuda:0')
m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 8, 256, False, 16)
m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
load_synthetic_checkpoint(m, 224)
m.feature_extractor.set_option('chunk_frames', 1600)
cp, cy, tp, ty = make_episode(S2, index=0, pin=True)
c, t = cp.clone(), tp.clone()          # pageable copies
assert not c.is_pinned()
cyd = cy.to(dev)
def run(cc, tt, n=6):
    def step():
        m.personalise(cc, cyd); lg = m.predict(tt); m._reset(); return lg.cpu()
    for _ in range(2): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): step()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for rep in range(2):
    for ramp in ((160, 320, 480), (96, 192, 320, 416), (128, 256, 384), (96, 192, 320, 416, 576)):
        m.stage_ramp = ramp
        print(ramp, f"pinned {run(cp, tp):.2f} ms  pageable {run(c, t):.2f} ms", flush=True)
