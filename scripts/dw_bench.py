"""Dev helper: per-layer timing of the depthwise kernel (C ABI) on the EfficientNet-B0 shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
lib = L.load(); dev = torch.device('cuda:0')
if os.environ.get('DW5S'): assert lib.orbit_set_global_option(b'dw5_staged', int(os.environ['DW5S'])) == 0
layers = [(112, 32, 3, 1), (112, 96, 3, 2), (56, 144, 3, 1), (56, 144, 5, 2), (28, 240, 5, 1), (28, 240, 3, 2), (14, 480, 3, 1),
          (14, 480, 5, 1), (14, 672, 5, 1), (14, 672, 5, 2), (7, 1152, 5, 1), (7, 1152, 3, 1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot = 0
for H, C, k, s in layers:
    Ho = (H + s - 1) // s
    x = torch.randn(B, H, H, C, device=dev); w = torch.randn(C, 1, k, k, device=dev)
    sc = torch.ones(C, device=dev); sh = torch.zeros(C, device=dev)
    y = torch.empty(B, Ho, Ho, C, device=dev)
    partial = torch.empty(lib.orbit_depthwise_partial_floats(B, H, H, C, k, s), device=dev)
    scratch = torch.empty(k * k * C, device=dev)
    ts = []
    for it in range(4):
        if not os.environ.get("NOFLUSH"): flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.orbit_depthwise_conv(L.ptr(x), L.ptr(w), L.ptr(sc), L.ptr(sh), L.ptr(y), L.ptr(partial), L.ptr(scratch), B, H, H, C, k, s, 1,
                                      L.stream_ptr(dev))
        e1.record(); torch.cuda.synchronize(); assert rc == 0
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:]); bytes_ = 4.0 * B * C * (H * H + Ho * Ho)
    print(f"H={H:4d} C={C:5d} k={k} s={s}: {t*1e3:8.1f} us  {bytes_/t/1e6:7.0f} GB/s", flush=True)
