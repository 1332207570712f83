"""Dev helper: fixed cost of a backbone pass -- per-family CUDA-event profile of the EfficientNet-B0 forward at several pass sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import orbit_b200
dev = torch.device('cuda:0')
if os.environ.get('SEF'):
    from orbit_b200 import lib as L
    assert L.load().orbit_set_global_option(b'se_frames', int(os.environ['SEF'])) == 0
m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 8, 256, False, 16)
m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
fe = m.feature_extractor
fe.set_option('chunk_frames', 1600)
xs = {n: torch.randn(n, 3, 224, 224, device=dev) for n in (96, 224, 480, 800, 1600)}
for n, x in xs.items():
    for _ in range(3): fe(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fe(x)
    e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / 5
    fe.set_option('profile', 1)
    fe(x); torch.cuda.synchronize()
    prof = fe.profile_read()
    fe.set_option('profile', 0)
    fam = {k: round(v['ms'], 3) for k, v in prof.items() if v['launches']}
    s = sum(fam.values())
    print(f"{n:5d} frames: {total:7.3f} ms ({total / n * 1000:6.2f} us/frame); sum of kernel times {s:7.3f} ms; families {fam}", flush=True)
