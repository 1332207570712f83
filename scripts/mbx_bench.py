"""Dev helper: the fused expand+depthwise kernel (orbit_mbconv_expand_dw) against the unfused pair (tcgen05 expand GEMM +
depthwise kernel) on the three EfficientNet-B0 blocks with 16 / 24 input channels.  usage: mbx_bench.py [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = L.load(); dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn):
    ts = []
    for _ in range(4):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:]) * 1e3
for name, H, Cin, C, k, s in [('b1.0', 112, 16, 96, 3, 2), ('b1.1', 56, 24, 144, 3, 1), ('b2.0', 56, 24, 144, 5, 2)]:
    Ho = (H + s - 1) // s
    x = torch.randn(B, H, H, Cin, device=dev); we = torch.randn(C, Cin, device=dev) * Cin ** -0.5
    one, zero = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    wd = torch.randn(C, 1, k, k, device=dev) * 0.3
    y = torch.empty(B, Ho, Ho, C, device=dev); e = torch.empty(B * H * H, C, device=dev)
    part = torch.empty(max(lib.orbit_mbconv_partial_floats(B, H, H, C, k, s), lib.orbit_depthwise_partial_floats(B, H, H, C, k, s)), device=dev)
    scr = torch.empty(k * k * C, device=dev); ws = torch.empty(2 * C * Cin, device=dev)
    st = L.stream_ptr(dev)
    def fused():
        assert lib.orbit_mbconv_expand_dw(L.ptr(x), L.ptr(we), L.ptr(one), L.ptr(zero), L.ptr(wd), L.ptr(one), L.ptr(zero), L.ptr(y),
                                          L.ptr(part), L.ptr(scr), B, H, H, Cin, C, k, s, L.stream_ptr(dev)) == 0
    def gemm():
        assert lib.orbit_pointwise_conv(L.ptr(x), L.ptr(we), L.ptr(one), L.ptr(zero), None, None, L.ptr(e), B * H * H, C, Cin, H * H, 1, 1,
                                        L.ptr(ws), L.stream_ptr(dev)) == 0
    def dw():
        assert lib.orbit_depthwise_conv(L.ptr(e), L.ptr(wd), L.ptr(one), L.ptr(zero), L.ptr(y), L.ptr(part), L.ptr(scr), B, H, H, C, k, s, 1,
                                        L.stream_ptr(dev)) == 0
    tf, tg, td = timeit(fused), timeit(gemm), timeit(dw)
    byts = 4.0 * B * (H * H * Cin + Ho * Ho * C)
    print(f"{name} B={B}: fused {tf:8.1f} us ({byts / tf / 1e3:6.0f} GB/s of in+out) | expand GEMM {tg:8.1f} + depthwise {td:8.1f} = {tg + td:8.1f} us", flush=True)
