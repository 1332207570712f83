"""Dev helper: per-role timeline of CTA 0 of ONE tcgen05 GEMM launch (orbit_debug_set_gemm_trace).
usage: gemm_trace.py B HW K N act gated resid"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L
B, hw, K, N, act, gated, resid = [int(a) for a in sys.argv[1:8]]
lib = L.load(); dev = torch.device('cuda:0')

M = B * hw
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * K ** -0.5
sc = torch.ones(N, device=dev); sh = torch.zeros(N, device=dev)
gate = torch.rand(B, K, device=dev) if gated else None
res = torch.randn(M, N, device=dev) if resid else None
out = torch.empty(M, N, device=dev); ws = torch.empty(2 * N * K, device=dev)
trace = torch.zeros(256, 16, dtype=torch.int32, device=dev)
def run():
    rc = lib.orbit_pointwise_conv(L.ptr(A), L.ptr(W), L.ptr(sc), L.ptr(sh), L.ptr(gate), L.ptr(res), L.ptr(out), M, N, K, hw, act, 1,
                                  L.ptr(ws), L.stream_ptr(dev)); assert rc == 0
run(); torch.cuda.synchronize()
lib.orbit_debug_set_gemm_trace(L.ptr(trace)); run(); torch.cuda.synchronize(); lib.orbit_debug_set_gemm_trace(None)
t = trace.cpu().numpy().astype('uint32').astype('int64')
t0 = int(t[0, 0])
names = ['P:empty_ok', 'P:tma_issued', 'X:full_ok', 'X:ready', 'M:main_empty_ok', '(unused)', 'M:ready_ok', 'M:committed', 'E:main_full_ok', 'E:arrived']
num_k = (K + 63) // 64
print(f"M={M} K={K} N={N} num_k={num_k}; clocks relative to the first producer stamp; one row per k-block step of CTA 0")
print('step ' + ' '.join(f'{n:>15s}' for n in names))
for s in range(32, 32 + 3 * num_k + 4):
    if s >= 256: break
    print(f'{s:4d} ' + ' '.join(f'{(int(t[s, i]) - t0) & 0xffffffff:15d}' for i in range(10)))
import numpy as np
sel = slice(32, 200)
for i, n in enumerate(names):
    d = np.diff((t[sel, i] - t0) & 0xffffffff)
    print(f"{n:16s} mean period {d.mean():8.1f} clk")
lat = lambda a, b: float((((t[sel, b] - t[sel, a]) & 0xffffffff).astype('int64')).mean())
if os.environ.get('EPI'):
    # epilogue tail of warp EPI_WARP0 (stamps exist only for the tiles whose slab it owns): slots 8 (main_full seen), 9 (arrived),
    # 10 (k loop done), 11 (correction added), 12 (staging free + scale/shift parked), 13 (slab written), 14 (TMA store issued)
    rows = [r for r in range(32, 200) if t[r, 14] != 0 and t[r, 8] != 0]
    for a, b, nm in ((8, 9, 'main_full -> arrived'), (9, 10, 'arrived -> k loop done'), (10, 11, 'correction ld + fma'), (11, 12, 'staging wait + table'),
                     (12, 13, 'scale/act/sts loop'), (13, 14, 'fence + TMA store issue')):
        d = [((int(t[r, b]) - int(t[r, a])) & 0xffffffff) for r in rows]
        print(f"E {nm:28s} {np.mean(d):8.1f} clk  (n={len(d)})")
    d = [((int(t[r, 8]) - int(t[r, 15])) & 0xffffffff) for r in rows if t[r, 15] != 0]
    print(f"E loop top -> main_full seen        {np.mean(d):8.1f} clk  (n={len(d)})")
    d = [((int(t[r + 1, 15]) - int(t[r, 14])) & 0xffffffff) for r in rows if r + 1 in rows and t[r + 1, 15] != 0]
    print(f"E store issued -> next loop top     {np.mean(d):8.1f} clk  (n={len(d)})")
    d = np.diff([int(t[r, 14]) for r in rows]) & 0xffffffff
    print(f"E slab-to-slab period of warp 0   {np.mean(d):8.1f} clk")
print(f"TMA issue -> X full_ok   {lat(1, 2):8.1f}\nX full_ok -> X ready      {lat(2, 3):8.1f}\nX ready -> M ready_ok     {lat(3, 6):8.1f}\n"
      f"M main_empty_ok->ready_ok {lat(4, 6):8.1f}\nM ready_ok -> committed   {lat(6, 7):8.1f}\nM committed -> E full_ok  {lat(7, 8):8.1f}\n"
      f"E full_ok -> E arrived    {lat(8, 9):8.1f}\nP empty_ok -> tma_issued  {lat(0, 1):8.1f}")
