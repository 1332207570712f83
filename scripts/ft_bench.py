"""Dev helper: FineTuner inner loop (orbit_linear_finetune), cooperative-grid kernel against the single-CTA one."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, orbit_b200
from orbit_b200.finetune import finetune_linear_head
from orbit_b200 import lib as L
dev = torch.device('cuda:0'); lib = L.load()
for (n, d, c) in ((80, 768, 8), (80, 1280, 8), (600, 1280, 12)):
    feats = torch.randn(n, d, device=dev) * 0.5; labels = torch.arange(n) % c
    for mode in (0, 1):
        lib.orbit_set_global_option(b'finetune_grid', mode)
        ts = []
        for it in range(6):
            head = orbit_b200.LinearClassifier(d, 1.0); head.init(c); head.to(dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            finetune_linear_head(head, feats, labels, 1024, 50, 1e-3, 'adam', {}, 1.0)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(n, d, c, 'grid' if mode else 'single', f"{sorted(ts)[2]*1e3:.0f} us")
