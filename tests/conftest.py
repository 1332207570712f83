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
