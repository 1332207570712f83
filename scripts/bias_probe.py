"""Dev probe: signed relative error of the pointwise GEMM modes vs fp64 (is the tensor-core accumulation biased?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orbit_b200 import lib as L
lib = L.load(); dev = torch.device('cuda:0')
kappa = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib.orbit_set_global_option(b'tc_debias_x1000', kappa); print('kappa x1000 =', kappa)
M, N = 4096, 64
for K in (32, 128, 512, 1152):
    for kind in ('positive', 'mixed'):
        g = torch.Generator().manual_seed(K)
        A = torch.rand(M, K, generator=g) + 0.5
        W = torch.rand(N, K, generator=g) + 0.5
        if kind == 'mixed':
            A = A * torch.sign(torch.randn(M, K, generator=g)); 
        A, W = A.to(dev), W.to(dev)
        sc = torch.ones(N, device=dev); sh = torch.zeros(N, device=dev)
        ref = A.double() @ W.double().t()
        line = f"K={K:5d} {kind:8s}"
        for mode in (0, 1):
            out = torch.empty(M, N, device=dev); ws = torch.empty(2 * N * K, device=dev)
            rc = lib.orbit_pointwise_conv(L.ptr(A), L.ptr(W), L.ptr(sc), L.ptr(sh), None, None, L.ptr(out), M, N, K, M, 0, mode, L.ptr(ws), L.stream_ptr(dev))
            torch.cuda.synchronize(); assert rc == 0
            d = (out.double() - ref)
            scale = ref.abs().mean()
            line += f" | m{mode}: mean {float(d.mean()/scale):+.2e} toward0 {float((d*torch.sign(ref)).mean()/scale):+.2e} rms {float(d.pow(2).mean().sqrt()/scale):.2e}"
        print(line, flush=True)
