"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/orbit_b200.h declares, error codes behave, and argument validation happens BEFORE any GPU work."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'orbit_b200.h')


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(orbit_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ('orbit_proto_configure', 'orbit_head_predict', 'orbit_pool_clips', 'orbit_engine_forward',
                 'orbit_engine_prepare', 'orbit_pointwise_conv', 'orbit_engine_calibrate'):
        assert must in names
    assert len(names) >= 25


def test_library_exports_every_declared_symbol():
    from orbit_b200 import lib as L
    lib = L.load()
    raw = C.CDLL(L.LIB_PATH)
    for name in declared_functions():
        assert hasattr(raw, name), f"{name} declared in orbit_b200.h but not exported"
        assert name in L._SIGNATURES, f"{name} has no ctypes signature in orbit_b200/lib.py"
    assert lib.orbit_abi_version() == 1


def test_error_strings_and_argument_validation_without_gpu():
    from orbit_b200 import lib as L
    lib = L.load()
    assert lib.orbit_error_string(0) == b"ok"
    assert b"invalid argument" in lib.orbit_error_string(-1)
    assert b"workspace" in lib.orbit_error_string(-3)
    # null pointers / bad sizes are refused before anything is launched
    assert lib.orbit_pool_clips(None, 1, 1, 4, None, None) == -1
    assert lib.orbit_pool_clips(None, 0, 1, 4, None, None) == 0          # empty input is legal
    assert lib.orbit_proto_configure(None, None, 1, 1, 4, 1, 0, None, None, None, None, None) == -1
    assert lib.orbit_head_predict(None, 1, 1, 4, None, None, 1, 0, 1.0, None, None, None) == -1
    assert lib.orbit_engine_forward(None, None, None, None, 1, 8, 8, None, None, 0, None) == -1
    assert lib.orbit_proto_configure_scratch_bytes(5, 1280) == 64 * 4 + 5 * 10 * 4


def test_engine_plan_metadata_matches_timm_efficientnet_b0():
    """Parameter names/sizes of the native plan = timm tf_efficientnet_b0 state_dict (no GPU needed)."""
    import torch
    from orbit_b200.feature_extractors import FeatureExtractor
    from oracle import backbones, parts
    fe = FeatureExtractor('efficientnet_b0')
    ref = backbones.build('efficientnet_b0')
    ref_sd, sd = ref.state_dict(), fe.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())
    for k in sd:
        assert sd[k].shape == ref_sd[k].shape, k
    assert sum(p.numel() for p in fe.parameters()) == 4007548
    assert fe.output_size == 1280
    # FiLM sites (film.py:38-74): 17 norm layers -> 34 tensors, 20,480 values; generator order = sorted names
    names = fe.film_parameter_names()
    assert names == parts.film_parameter_names('efficientnet_b0', ref)
    assert len(names) == 34 and sum(n for _, n, _ in fe.film_layout()) == 20480
    assert [n for n, _, _ in fe.film_layout()] == sorted(names)
    # load_state_dict writes straight into the flat blob the kernels read
    backbones.seeded_init(ref, 3)
    fe.load_state_dict(ref.state_dict(), strict=True)
    name, numel, off = fe._table[10]
    assert torch.equal(fe._blob[off:off + numel], ref.state_dict()[name].flatten())
    with pytest.raises(ValueError):
        FeatureExtractor('resnet9000')


def test_shipped_library_is_not_an_experiment_build():
    """a library compiled with a timing-experiment flag (ORBIT_EXP_*: wrong results by design) must never pass for the product"""
    from orbit_b200 import lib as L
    assert L.load().orbit_experiment_build() == 0


def test_build_is_stale_when_the_nvcc_flags_differ(monkeypatch):
    """build.py must not mistake a library built with other flags (ORBIT_NVCC_EXTRA: trace / experiment builds) for the shipped
    one: the flags are stamped next to the .so and compared, not only the file times"""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("orbit_build", os.path.join(root, "orbit-dataset_b200", "build.py"))
    build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build)
    monkeypatch.delenv("ORBIT_NVCC_EXTRA", raising=False)
    monkeypatch.delenv("ORBIT_LINK_LIBCUDA", raising=False)
    build.build_library()                       # up to date after this (no-op when the tree was just built)
    assert not build._stale()
    monkeypatch.setenv("ORBIT_NVCC_EXTRA", "-DORBIT_GEMM_TRACE")
    assert build._stale()
    monkeypatch.delenv("ORBIT_NVCC_EXTRA")
    assert not build._stale()
