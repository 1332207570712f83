"""Validates the CPU arm's extrapolation (bench.py times a 240-frame sample of the S2 episode and scales by frame count):
times ONE full 2,240-frame S2 episode with the same oracle port on the same host cores and prints both."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import cpu_reference_episode_seconds
from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import S2, calibration_frames, make_episode
threads = os.cpu_count()
torch.set_num_threads(threads)
scaled_s, sample_s, _, info = cpu_reference_episode_seconds('S2', None, 0)
oracle = OracleRecogniser('efficientnet_b0', False, 'proto', S2.clip_length, 16, 1.0, 1991, calibration_frames(224))
ctx, ctx_y, tgt, _ = make_episode(S2, index=0)
t0 = time.perf_counter()
oracle.personalise(ctx, ctx_y)
logits = oracle.predict(tgt)
full_s = time.perf_counter() - t0
print(json.dumps({"host_threads": threads, "sample": info["sample"], "scaled_episode_s": scaled_s, "full_episode_s": full_s,
                  "full_over_scaled": full_s / scaled_s, "episodes_per_sec_full": 1.0 / full_s}))
