import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import orbit_b200
from orbit_b200 import lib as L
from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode
gr = np.load('/root/repo/tests/golden/recogniser.npz')
dev = torch.device('cuda:0')
weights = OracleRecogniser('efficientnet_b0', False, 'linear', 1, 5, 1.0, 1991, calibration_frames(64)).state_dict()
for stream in (1, 0):
    L.load().orbit_set_global_option(b'tc_stream', stream)
    m = orbit_b200.MultiStepFewShotRecogniser('efficientnet_b0', False, 'linear', 1, 5, False, 1.0)
    m.load_state_dict(weights, strict=True); m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
    ctx, ctx_y, tgt, _ = make_episode(EpisodeSpec(4, 3, 2, 1, 64), index=2)
    args = {'num_grad_steps': 5, 'learning_rate': 0.1, 'optimizer': 'adam', 'loss_fn': None, 'extractor_lr_scale': 0.1,
            'epsilon': 1e-8, 'weight_decay': 0.0, 'betas': (0.9, 0.999), 'momentum': 0.0}
    m.personalise(ctx[:-1], ctx_y[:-1], dict(args))
    dw = (m.classifier.weight.detach().cpu() - torch.as_tensor(gr['finetune2_weight'])).abs()
    db = (m.classifier.bias.detach().cpu() - torch.as_tensor(gr['finetune2_bias'])).abs()
    lg = m.predict(tgt).cpu()
    print('stream', stream, 'dW max', dw.max().item(), 'n>1e-4', int((dw > 1e-4).sum()), 'of', dw.numel(), 'dB', db.max().item(),
          'dlogit', (lg - torch.as_tensor(gr['finetune2_logits'])).abs().max().item(), 'max|logit|', float(np.abs(gr['finetune2_logits']).max()))
    if stream: 
        idx = (dw > 1e-4).nonzero()
        print(idx[:10].tolist(), m.classifier.weight.detach().cpu()[dw > 1e-4][:10], torch.as_tensor(gr['finetune2_weight'])[dw > 1e-4][:10])
