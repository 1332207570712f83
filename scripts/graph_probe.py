"""Dev helper: how much of a backbone pass is launch gaps? Eager launches vs one CUDA graph of the same pass."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import orbit_b200
from orbit_b200.synthetic import load_synthetic_checkpoint
dev = torch.device('cuda:0')
m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 8, 256, False, 16)
m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
load_synthetic_checkpoint(m, 224)
fe = m.feature_extractor
fe.set_option('chunk_frames', 1600)
for n in (96, 640, 1600):
    x = torch.randn(n, 3, 224, 224, device=dev)
    for _ in range(3): y = fe(x)
    torch.cuda.synchronize()
    def timeit(fn, reps=5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    eager = timeit(lambda: fe(x))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fe(x)
    torch.cuda.current_stream().wait_stream(s)
    try:
        with torch.cuda.graph(g):
            yg = fe(x)
        graph = timeit(g.replay)
        ok = torch.equal(yg, y)
        print(f"{n:5d} frames: eager {eager:7.3f} ms, graph {graph:7.3f} ms ({eager - graph:+.3f} ms), same result: {ok}", flush=True)
    except Exception as e:
        print(f"{n} frames: eager {eager:.3f} ms; graph capture failed: {type(e).__name__}: {str(e)[:200]}", flush=True)
