"""Dev helper: where does the host-input (e2e) path lose time? H2D bandwidth, slice size, ramp, chunk size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import orbit_b200
from orbit_b200.synthetic import S2, load_synthetic_checkpoint, make_episode

dev = torch.device('cuda:0')
m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 8, 256, False, 16)
m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
load_synthetic_checkpoint(m, 224)
c, cy, t, ty = make_episode(S2, index=0, pin=True)
cyd = cy.to(dev)
buf = torch.empty_like(c, device=dev)
for _ in range(2):
    buf.copy_(c, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); buf.copy_(c, non_blocking=True); torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"H2D pinned: {c.numel()*4/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms for support set)", flush=True)
cd, td = c.to(dev), t.to(dev)
del buf

def run(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def step_host():
    m.personalise(c, cyd); lg = m.predict(t); m._reset(); return lg.cpu()
def step_dev():
    m.personalise(cd, cyd); lg = m.predict(td); m._reset(); return lg.cpu()

fe = m.feature_extractor
for chunk in (() if os.environ.get('RAMP_ONLY') else (640, 1600)):
    fe.set_option('chunk_frames', chunk)
    print(f"device-resident, chunk {chunk}: {run(step_dev):.1f} ms", flush=True)
fe.set_option('chunk_frames', 1600)
RAMPS = ((160, (96, 224, 480)), (160, (160, 320, 480)), (160, (64, 160, 288, 448)), (160, (128, 288, 544)), (160, (96, 192, 320, 416)),
         (160, (160, 352, 512)), (160, (96, 224, 480)), (160, (160, 320, 480)))
if os.environ.get('RAMP_ONLY'):
    fe.set_option('chunk_frames', 1600)
for cf, ramp in RAMPS:
    m.stage_copy_frames, m.stage_ramp = cf, ramp
    step_host()
    print(f"host clips, copy {cf} ramp {ramp}: {run(step_host, 6):.1f} ms", flush=True)
if os.environ.get('RAMP_ONLY'):
    sys.exit(0)
# CPU-side cost of enqueueing one device-resident episode (no sync inside)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    m.personalise(cd, cyd); lg = m.predict(td); m._reset()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"enqueue time per episode {1e3*(t1-t0)/3:.1f} ms, total {1e3*(t2-t0)/3:.1f} ms")
# CPU-side enqueue cost of ONE backbone pass on an idle device (queue empty => nothing blocks)
for nfr in (16, 96, 640):
    x = torch.randn(nfr, 3, 224, 224, device=dev)
    fe(x); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fe(x); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t0))
    print(f"one pass of {nfr} frames: enqueue {1e3*min(t[0] for t in ts):.2f} ms, total {1e3*min(t[1] for t in ts):.2f} ms")
