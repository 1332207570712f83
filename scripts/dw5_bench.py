"""Dev helper: the 5x5 stride-1 depthwise layers of EfficientNet-B0 (orbit_depthwise_conv), us per launch and algorithmic TB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L
dev = torch.device('cuda:0')
lib = L.load()
B = int(os.environ.get('FRAMES', 1600))
for H, C in ((28, 240), (14, 480), (14, 672), (7, 1152)):
    x = torch.randn(B, H, H, C, device=dev); x2 = torch.randn(B, H, H, C, device=dev)
    w = torch.randn(C, 1, 5, 5, device=dev) * 0.2; sc = torch.ones(C, device=dev); sh = torch.zeros(C, device=dev)
    y = torch.empty(B, H, H, C, device=dev)
    partial = torch.empty(lib.orbit_depthwise_partial_floats(B, H, H, C, 5, 1), device=dev)
    scratch = torch.empty(25 * C, device=dev)
    ts = []
    for it in range(10):
        src = x if it % 2 else x2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.orbit_depthwise_conv(L.ptr(src), L.ptr(w), L.ptr(sc), L.ptr(sh), L.ptr(y), L.ptr(partial), L.ptr(scratch), B, H, H, C, 5, 1, 1,
                                         L.stream_ptr(dev)), "dw")
        e1.record(); torch.cuda.synchronize()
        if it >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(f"{H}x{H}x{C}: {t:7.1f} us / {B} frames, {2 * x.numel() * 4 / t / 1e6:.2f} TB/s", flush=True)
