import sys; sys.path.insert(0,'/root/repo')
import torch, orbit_b200
from orbit_b200 import lib as L
from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode
dev = torch.device("cuda:0")
oracle = OracleRecogniser('efficientnet_b0', False, 'proto', clip_length=2, batch_size=256, calib_input=calibration_frames(96))
ctx, ctx_y, tgt, _ = make_episode(EpisodeSpec(5, 2, 2, 2, 96), index=0)
oracle.personalise(ctx, ctx_y); ref = oracle.predict(tgt)
for opts in ({}, {'tc_stream':0}, {'mbconv_stream':0}, {'dw5_staged':0}, {'tc_stream':0,'mbconv_stream':0,'dw5_staged':0}):
    for k in ('tc_stream','mbconv_stream','dw5_staged'): L.load().orbit_set_global_option(k.encode(), opts.get(k,1))
    model = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 2, 256, False, 16)
    model.load_state_dict(oracle.state_dict(), strict=True); model._set_device(dev); model._send_to_device(); model.set_test_mode(True)
    model.personalise(ctx.to(dev), ctx_y.to(dev)); logits = model.predict(tgt.to(dev))
    print(opts, 'err', (logits.cpu()-ref).abs().max().item(), 'max|logit|', ref.abs().max().item())
