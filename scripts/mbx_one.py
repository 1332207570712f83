"""Dev helper (ncu target): a few launches of the fused expand+depthwise kernel.  usage: mbx_one.py B H Cin C k s"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L
B, H, Cin, C, k, s = [int(a) for a in sys.argv[1:7]]
lib = L.load(); dev = torch.device('cuda:0')
Ho = (H + s - 1) // s
x = torch.randn(B, H, H, Cin, device=dev); we = torch.randn(C, Cin, device=dev) * Cin ** -0.5
one, zero = torch.ones(C, device=dev), torch.zeros(C, device=dev)
wd = torch.randn(C, 1, k, k, device=dev) * 0.3
y = torch.empty(B, Ho, Ho, C, device=dev)
part = torch.empty(lib.orbit_mbconv_partial_floats(B, H, H, C, k, s), device=dev)
scr = torch.empty(k * k * C, device=dev)
for _ in range(3):
    assert lib.orbit_mbconv_expand_dw(L.ptr(x), L.ptr(we), L.ptr(one), L.ptr(zero), L.ptr(wd), L.ptr(one), L.ptr(zero), L.ptr(y),
                                      L.ptr(part), L.ptr(scr), B, H, H, Cin, C, k, s, L.stream_ptr(dev)) == 0
torch.cuda.synchronize()
