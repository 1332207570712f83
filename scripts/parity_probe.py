"""Dev probe: episode logits error vs the oracle for the current library (ORBIT_B200_LIB selects a variant)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, orbit_b200
from oracle.recogniser import OracleRecogniser
from orbit_b200.synthetic import EpisodeSpec, calibration_frames, make_episode
dev = torch.device('cuda:0')
from orbit_b200 import lib as L
kappa = int(sys.argv[1]) if len(sys.argv) > 1 else 0
L.load().orbit_set_global_option(b'tc_debias_x1000', kappa); print('kappa x1000 =', kappa)
for size, spec in ((96, EpisodeSpec(5, 3, 4, 2, 96)), (224, EpisodeSpec(5, 4, 4, 2, 224))):
  if True:
    oracle = OracleRecogniser('efficientnet_b0', False, 'proto', 2, 256, calib_input=calibration_frames(size))
    m = orbit_b200.SingleStepFewShotRecogniser('efficientnet_b0', False, 'proto', 2, 256, False, 16)
    m.load_state_dict(oracle.state_dict(), strict=True); m._set_device(dev); m._send_to_device(); m.set_test_mode(True)
    for idx in (0, 3):
        ctx, cy, tgt, ty = make_episode(spec, index=idx)
        oracle.reset(); oracle.personalise(ctx, cy); ref = oracle.predict(tgt)
        ref64 = None
        for gemm in (1,):
            m.feature_extractor.set_option('gemm', gemm)
            m.personalise(ctx.to(dev), cy.to(dev)); lg = m.predict(tgt.to(dev)).cpu(); m._reset()
            f = m.feature_extractor(tgt.flatten(end_dim=1).to(dev)).cpu()
            fo = oracle.extractor(tgt.flatten(end_dim=1))
            rel = ((f - fo) / fo.abs().clamp_min(1e-3))
            print(f"size={size} ep={idx} gemm={gemm}: max|dlogit|={(lg-ref).abs().max():.2e} |logit|max={ref.abs().max():.1f} "
                  f"feat err max={(f-fo).abs().max():.2e} mean signed rel={rel.mean():+.2e}", flush=True)
