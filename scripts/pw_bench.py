"""Dev helper: per-layer timing of the pointwise-conv GEMM (C ABI) on the EfficientNet-B0 layer shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
modes = [int(m) for m in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2").split(',')]
lib = L.load()
if os.environ.get('NARROW'): assert lib.orbit_set_global_option(b'tc_narrow', int(os.environ['NARROW'])) == 0
if os.environ.get('STREAM'): assert lib.orbit_set_global_option(b'tc_stream', int(os.environ['STREAM'])) == 0
if os.environ.get('FIXED'): assert lib.orbit_set_global_option(b'tc_fixed_slabs', int(os.environ['FIXED'])) == 0
if os.environ.get('DMIN'): assert lib.orbit_set_global_option(b'tc_double_min_stages', int(os.environ['DMIN'])) == 0
dev = torch.device('cuda:0')
# (HW, K, N, act, gated, residual, name)
layers = [(112*112, 32, 16, 0, 1, 0, 'b0 proj'), (112*112, 16, 96, 1, 0, 0, 'b1.0 exp'), (56*56, 96, 24, 0, 1, 0, 'b1.0 proj'),
          (56*56, 24, 144, 1, 0, 0, 'b1.1 exp'), (56*56, 144, 24, 0, 1, 1, 'b1.1 proj'), (28*28, 144, 40, 0, 1, 0, 'b2.0 proj'),
          (28*28, 40, 240, 1, 0, 0, 'b2.1 exp'), (28*28, 240, 40, 0, 1, 1, 'b2.1 proj'), (14*14, 240, 80, 0, 1, 0, 'b3.0 proj'),
          (14*14, 80, 480, 1, 0, 0, 'b3.1 exp'), (14*14, 480, 80, 0, 1, 1, 'b3.1 proj'), (14*14, 480, 112, 0, 1, 0, 'b4.0 proj'),
          (14*14, 112, 672, 1, 0, 0, 'b4.1 exp'), (14*14, 672, 112, 0, 1, 1, 'b4.1 proj'), (7*7, 672, 192, 0, 1, 0, 'b5.0 proj'),
          (7*7, 192, 1152, 1, 0, 0, 'b5.1 exp'), (7*7, 1152, 192, 0, 1, 1, 'b5.1 proj'), (7*7, 1152, 320, 0, 1, 0, 'b6.0 proj'),
          (7*7, 320, 1280, 1, 0, 0, 'head')]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"B={B} frames; columns: layer M K N | per mode: us, GB/s (algorithmic), TFLOP/s")
for hw, K, N, act, gated, resid, name in layers:
    M = B * hw
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * K ** -0.5
    sc = torch.ones(N, device=dev); sh = torch.zeros(N, device=dev)
    gate = torch.rand(B, K, device=dev) if gated else None
    res = torch.randn(M, N, device=dev) if resid else None
    out = torch.empty(M, N, device=dev); ws = torch.empty(2 * N * K, device=dev)
    bytes_ = 4.0 * (M * K + M * N * (2 if resid else 1) + N * K)
    line = f"{name:10s} M={M:8d} K={K:5d} N={N:5d} |"
    for mode in modes:
        ts = []
        for it in range(4):
            if not os.environ.get("NOFLUSH"): flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.orbit_pointwise_conv(L.ptr(A), L.ptr(W), L.ptr(sc), L.ptr(sh), L.ptr(gate), L.ptr(res), L.ptr(out), M, N, K,
                                          hw, act, mode, L.ptr(ws), L.stream_ptr(dev))
            e1.record(); torch.cuda.synchronize()
            assert rc == 0, rc
            ts.append(e0.elapsed_time(e1))
        t = min(ts[1:])
        line += f" m{mode}: {t*1e3:8.1f}us {bytes_/t/1e6:7.0f}GB/s {2.0*M*N*K/t/1e9:6.1f}TF |"
    print(line, flush=True)
