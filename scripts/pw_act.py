"""Dev helper: one GEMM shape timed with act = 0 / 1 / 2 (how much of an expand layer is the SiLU epilogue?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L
lib = L.load(); dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (B, hw, K, N) in ((640, 12544, 16, 96), (640, 3136, 24, 144), (640, 784, 40, 240), (640, 49, 192, 1152)):
    M = B * hw
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * K ** -0.5
    sc = torch.ones(N, device=dev); sh = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev); ws = torch.empty(2 * N * K, device=dev)
    line = f"M={M} K={K} N={N}:"
    for act in (0, 1, 2):
        ts = []
        for it in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.orbit_pointwise_conv(L.ptr(A), L.ptr(W), L.ptr(sc), L.ptr(sh), None, None, L.ptr(out), M, N, K, hw, act, 1, L.ptr(ws), L.stream_ptr(dev))
            e1.record(); torch.cuda.synchronize(); assert rc == 0
            ts.append(e0.elapsed_time(e1))
        line += f"  act{act}: {min(ts[1:])*1e3:7.1f} us"
    print(line, flush=True)
