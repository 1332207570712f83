"""Dev helper: per-launch kernel durations (torch.profiler / CUPTI) of one EfficientNet-B0 backbone pass at a small and a large
pass size, launch by launch: which launches do not shrink with the frame count (the fixed cost of a pass)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import orbit_b200
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda:0')
m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 8, 256, False, 16)
m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
fe = m.feature_extractor
fe.set_option('chunk_frames', 1600)
small, large = int(os.environ.get('SMALL', 160)), int(os.environ.get('LARGE', 1600))
res = {}
for n in (small, large):
    x = torch.randn(n, 3, 224, 224, device=dev)
    for _ in range(3): fe(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fe(x); torch.cuda.synchronize()
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    res[n] = [(e.name, e.time_range.start, e.time_range.end) for e in ev]
    t0, t1 = res[n][0][1], res[n][-1][2]
    busy = sum(b - a for _, a, b in res[n])
    print(f"{n} frames: {len(ev)} launches, span {(t1 - t0) / 1e3:.3f} ms, kernel time {busy / 1e3:.3f} ms, gaps {(t1 - t0 - busy) / 1e3:.3f} ms")
a, b = res[small], res[large]
assert len(a) == len(b)
rows = []
for (na, sa, ea), (nb, sb, eb) in zip(a, b):
    da, db = ea - sa, eb - sb
    rows.append((da - db * small / large, da, db, na[:70]))
print(f"launch: t({small}) us, t({large}) us, excess over linear scaling us")
tot = 0.0
for exc, da, db, name in rows:
    tot += exc
    print(f"{da:8.1f} {db:9.1f} {exc:8.1f}  {name}")
print("total excess", tot)
