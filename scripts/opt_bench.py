"""Dev helper: A/B of process-wide kernel switches (orbit_set_global_option) on the whole EfficientNet-B0 forward.
usage: opt_bench.py "key=value,key=value" "key=value" ...   (each argument = one configuration; "-" = defaults)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import orbit_b200
from orbit_b200 import lib as L

dev = torch.device('cuda:0')
lib = L.load()
m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 8, 256, False, 16)
m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
fe = m.feature_extractor
frames = int(os.environ.get('FRAMES', 1600))
fe.set_option('chunk_frames', frames)
x = torch.randn(frames, 3, 224, 224, device=dev)
y = torch.randn(frames, 3, 224, 224, device=dev)      # alternate inputs: nothing stays in L2
for rep in range(2):
    for cfg in sys.argv[1:] or ['-']:
        if cfg != '-':
            for kv in cfg.split(','):
                k, v = kv.split('=')
                if k.startswith('engine.'):
                    fe.set_option(k[7:], int(v))
                else:
                    assert lib.orbit_set_global_option(k.encode(), int(v)) == 0, kv
        for _ in range(2):
            fe(x); fe(y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fe(x); fe(y)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 6
        print(f"{cfg:40s} {ms:8.3f} ms / {frames} frames -> S2 episode (2240 frames) {ms / frames * 2240:7.2f} ms", flush=True)
