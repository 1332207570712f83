"""Dev helper: the EfficientNet stem (orbit_stem_conv) on 224x224 frames; sweeps the groups-per-block switch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L
dev = torch.device('cuda:0')
lib = L.load()
for B in (1600, 160):
    x = torch.randn(B, 3, 224, 224, device=dev)
    w, sc, sh = torch.randn(32, 3, 3, 3, device=dev) * 0.2, torch.ones(32, device=dev), torch.zeros(32, device=dev)
    y = torch.empty(B, 112, 112, 32, device=dev)
    for groups in (1, 2, 4, 8, 16):
        assert lib.orbit_set_global_option(b'stem_groups', groups) == 0
        ts = []
        for it in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.orbit_stem_conv(L.ptr(x), L.ptr(w), L.ptr(sc), L.ptr(sh), L.ptr(y), B, 224, 224, 1, L.stream_ptr(dev)), "stem")
            e1.record(); torch.cuda.synchronize()
            if it >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts)[len(ts) // 2]
        print(f"groups/block {groups:2d}: {t:8.1f} us / {B} frames, {(x.numel() + y.numel()) * 4 / t / 1e6:.2f} TB/s", flush=True)
