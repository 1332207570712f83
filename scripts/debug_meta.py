import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import orbit_b200
from oracle import parts
from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode
dev = torch.device('cuda:0')
spec = EpisodeSpec(4, 3, 2, 2, 64)
calib = calibration_frames(64)
oracle = OracleRecogniser('efficientnet_b0', True, 'versa', 2, 4, 1.0, 1991, calib)
m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', True, 'versa', 2, 4, False, 1, 1.0)
m.load_state_dict(oracle.state_dict(), strict=True)
m._set_device(dev); m._send_to_device()
ctx, ctx_y, tgt, tgt_y = make_episode(spec, index=3)
m.set_test_mode(True)
with torch.no_grad():
    oracle.personalise(ctx, ctx_y)
    lo = oracle.predict(tgt)
    zo = oracle._task_embedding(ctx)
    for host in (True, False):
        c, t = (ctx, tgt) if host else (ctx.to(dev), tgt.to(dev))
        m.personalise(c, ctx_y.to(dev))
        z0 = m._get_task_embedding_in_batches(c).clone()
        print('host' if host else 'device')
        print(' z vs oracle', (z0.cpu() - zo).abs().max().item(), zo.abs().max().item())
        for k in ('bn1.weight', 'blocks.3.1.bn2.bias'):
            print(' film', k, (m.film_dict[k].cpu() - oracle.film_dict[k]).abs().max().item())
        feats0 = m._get_features_in_batches(c, m.film_dict).clone()
        fo = oracle._features(ctx, oracle.film_dict)
        print(' ctx feats vs oracle', (feats0.cpu() - fo).abs().max().item(), fo.abs().max().item())
        print(' head W', (m.classifier.weight.cpu() - oracle.head[0]).abs().max().item(), oracle.head[0].abs().max().item())
        l0 = m.predict(t).clone()
        print(' logits vs oracle', (l0.cpu() - lo).abs().max().item(), lo.abs().max().item())
        m._reset()
