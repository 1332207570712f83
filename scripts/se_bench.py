"""Dev helper: per-launch time of the squeeze-excite gate (orbit_se_gate) on the EfficientNet-B0 gate shapes, ring-streamed
kernel against the plain one (orbit_set_global_option "se_ring")."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L

dev = torch.device('cuda:0')
lib = L.load()
SHAPES = [(32, 8, 4), (96, 4, 4), (144, 6, 2), (240, 10, 1), (480, 20, 2), (672, 28, 2), (1152, 48, 2)]
flush = torch.empty(64 * 1024 * 1024, device=dev)
for B in (1600, 640, 160):
    for C, R, groups in SHAPES:
        partial = torch.randn(B, groups, C, device=dev)
        w1, b1, w2t, b2 = torch.randn(R, C, device=dev), torch.randn(R, device=dev), torch.randn(R, C, device=dev), torch.randn(C, device=dev)
        gate = torch.empty(B, C, device=dev)
        out = []
        for ring in (0, 1):
            lib.orbit_set_global_option(b'se_ring', ring)
            ts = []
            for it in range(12):
                flush.zero_()                      # weights and sums come from HBM / a cold L2, as inside a pass
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                L.check(lib.orbit_se_gate(L.ptr(partial), groups, 49, L.ptr(w1), L.ptr(b1), L.ptr(w2t), L.ptr(b2), L.ptr(gate), B, C, R,
                                          L.stream_ptr(dev)), "se")
                e1.record(); torch.cuda.synchronize()
                if it >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
            out.append(sorted(ts)[len(ts) // 2])
        print(f"B={B:5d} C={C:5d} R={R:3d}  plain {out[0]:7.1f} us   ring {out[1]:7.1f} us", flush=True)
