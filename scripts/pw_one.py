"""Dev helper (ncu target): a few launches of ONE pointwise GEMM shape.  usage: pw_one.py B HW K N act gated resid [mode]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L
B, hw, K, N, act, gated, resid = [int(a) for a in sys.argv[1:8]]
mode = int(sys.argv[8]) if len(sys.argv) > 8 else 1
lib = L.load(); dev = torch.device('cuda:0')
M = B * hw
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * K ** -0.5
sc = torch.ones(N, device=dev); sh = torch.zeros(N, device=dev)
gate = torch.rand(B, K, device=dev) if gated else None
res = torch.randn(M, N, device=dev) if resid else None
out = torch.empty(M, N, device=dev); ws = torch.empty(2 * N * K, device=dev)
for it in range(3):
    rc = lib.orbit_pointwise_conv(L.ptr(A), L.ptr(W), L.ptr(sc), L.ptr(sh), L.ptr(gate), L.ptr(res), L.ptr(out), M, N, K,
                                  hw, act, mode, L.ptr(ws), L.stream_ptr(dev))
    assert rc == 0, rc
torch.cuda.synchronize()
