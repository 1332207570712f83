"""Dev helper (ncu target): a few launches of ONE depthwise shape.  usage: dw_one.py B H C k s"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L
B, H, C, k, s = [int(a) for a in sys.argv[1:6]]
lib = L.load(); dev = torch.device('cuda:0')
Ho = (H + s - 1) // s
x = torch.randn(B, H, H, C, device=dev); w = torch.randn(C, 1, k, k, device=dev)
sc = torch.ones(C, device=dev); sh = torch.zeros(C, device=dev)
y = torch.empty(B, Ho, Ho, C, device=dev)
partial = torch.empty(lib.orbit_depthwise_partial_floats(B, H, H, C, k, s), device=dev)
scratch = torch.empty(k * k * C, device=dev)
for it in range(3):
    rc = lib.orbit_depthwise_conv(L.ptr(x), L.ptr(w), L.ptr(sc), L.ptr(sh), L.ptr(y), L.ptr(partial), L.ptr(scratch), B, H, H, C, k, s, 1,
                                  L.stream_ptr(dev))
    assert rc == 0
torch.cuda.synchronize()
