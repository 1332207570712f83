"""Dev helper: times the native EfficientNet-B0 forward for several chunk sizes (device-resident frames)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import orbit_b200

dev = torch.device('cuda:0')
gemm = int(sys.argv[1]) if len(sys.argv) > 1 else 0
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 8, 256, False, 16)
m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
fe = m.feature_extractor
fe.set_option('gemm', gemm)
x = torch.randn(frames, 3, 224, 224, device=dev)
for chunk in [int(c) for c in (sys.argv[3] if len(sys.argv) > 3 else "4,8,16,32,64,128").split(',')]:
    fe.set_option('chunk_frames', chunk)
    for _ in range(2):
        fe(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        fe(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"gemm={gemm} chunk={chunk:4d}: {ms:8.2f} ms / {frames} frames = {ms/frames*1000:7.1f} us/frame "
          f"-> {frames/ms*1000:9.0f} frames/s  (S2 episode 2240 frames: {ms/frames*2240:7.1f} ms)", flush=True)
