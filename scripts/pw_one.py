"""Dev helper for ncu: runs orbit_pointwise_conv on one shape a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L
M, N, K, rpf, act, gated, resid, mode = [int(x) for x in sys.argv[1:9]]
lib = L.load(); dev = torch.device('cuda:0')
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * K ** -0.5
sc = torch.ones(N, device=dev); sh = torch.zeros(N, device=dev)
gate = torch.rand((M + rpf - 1) // rpf, K, device=dev) if gated else None
res = torch.randn(M, N, device=dev) if resid else None
out = torch.empty(M, N, device=dev); ws = torch.empty(2 * N * K, device=dev)
for _ in range(3):
    rc = lib.orbit_pointwise_conv(L.ptr(A), L.ptr(W), L.ptr(sc), L.ptr(sh), L.ptr(gate), L.ptr(res), L.ptr(out), M, N, K, rpf, act, mode,
                                  L.ptr(ws), L.stream_ptr(dev))
    assert rc == 0
torch.cuda.synchronize()
