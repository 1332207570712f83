"""ctypes binding of liborbit_b200.so (the C ABI declared in include/orbit_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised."""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORBIT_B200_LIB", os.path.join(HERE, "liborbit_b200.so"))   # override: dev A/B builds

_i, _i64, _f, _p = C.c_int, C.c_int64, C.c_float, C.c_void_p
_SIGNATURES = {
    "orbit_abi_version": (_i, []),
    "orbit_error_string": (C.c_char_p, [_i]),
    "orbit_device_check": (_i, []),
    "orbit_set_global_option": (_i, [C.c_char_p, _i]),
    "orbit_get_global_option": (_i, [C.c_char_p, C.POINTER(_i)]),
    "orbit_debug_set_gemm_trace": (_i, [_p]),
    "orbit_pool_clips": (_i, [_p, _i, _i, _i, _p, _p]),
    "orbit_pool_history": (_i, [_p, _i, _i, _i, _p, _p]),
    "orbit_proto_configure_scratch_bytes": (_i64, [_i, _i]),
    "orbit_proto_configure": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "orbit_head_predict": (_i, [_p, _i, _i, _i, _p, _p, _i, _i, _f, _p, _p, _p]),
    "orbit_film_table_entry_bytes": (_i, []),
    "orbit_film_generate": (_i, [_p, _p, _i, _i, _p, _i, _p, _p]),
    "orbit_dense_rows": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "orbit_mahalanobis_configure_workspace_bytes": (_i64, [_i, _i, _i]),
    "orbit_mahalanobis_configure": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "orbit_mahalanobis_predict_workspace_bytes": (_i64, [_i, _i]),
    "orbit_mahalanobis_predict": (_i, [_p, _i, _i, _p, _p, _i, _f, _p, _p, _p]),
    "orbit_linear_finetune_scratch_bytes": (_i64, [_i, _i, _i]),
    "orbit_linear_finetune": (_i, [_p, _p, _i, _i, _i, _i, _i, _f, _f, _f, _f, _f, _f, _f, _p, _p, _p, _p]),
    "orbit_pointwise_conv": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "orbit_depthwise_partial_floats": (_i64, [_i, _i, _i, _i, _i, _i]),
    "orbit_depthwise_conv": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "orbit_mbconv_partial_floats": (_i64, [_i, _i, _i, _i, _i, _i]),
    "orbit_mbconv_partial_groups": (_i, [_i, _i, _i, _i, _i, _i]),
    "orbit_mbconv_expand_dw": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "orbit_engine_create": (_i, [C.POINTER(_p), _i]),
    "orbit_engine_destroy": (None, [_p]),
    "orbit_engine_feat_dim": (_i, [_p]),
    "orbit_engine_num_params": (_i, [_p]),
    "orbit_engine_param_info": (_i, [_p, _i, C.c_char_p, _i, C.POINTER(_i64), C.POINTER(_i64)]),
    "orbit_engine_param_shape": (_i, [_p, _i, C.POINTER(_i), C.POINTER(_i64)]),
    "orbit_engine_param_floats": (_i64, [_p]),
    "orbit_engine_num_film": (_i, [_p]),
    "orbit_engine_film_info": (_i, [_p, _i, C.c_char_p, _i, C.POINTER(_i64), C.POINTER(_i64)]),
    "orbit_engine_film_floats": (_i64, [_p]),
    "orbit_engine_derived_floats": (_i64, [_p]),
    "orbit_engine_prepare": (_i, [_p, _p, _p, _p, _p]),
    "orbit_engine_set_option": (_i, [_p, C.c_char_p, _i]),
    "orbit_engine_get_option": (_i, [_p, C.c_char_p, C.POINTER(_i)]),
    "orbit_engine_workspace_bytes": (_i64, [_p, _i, _i]),
    "orbit_engine_macs": (_i64, [_p, _i, _i]),
    "orbit_video_stats": (_i, [_p, _p, _i, _p, _p, _p, _i, _p, _p, _p]),
    "orbit_se_gate": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "orbit_experiment_build": (_i, []),
    "orbit_stem_conv": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "orbit_conv_first": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "orbit_conv3x3_scratch_floats": (_i64, [_i, _i, _i, _i, _i, _i]),
    "orbit_conv3x3": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i64, _p]),
    "orbit_engine_forward": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _i64, _p]),
    "orbit_engine_calibrate": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _i64, _p]),
    "orbit_engine_train_saved_floats": (_i64, [_p, _i, _i]),
    "orbit_engine_train_derived_floats": (_i64, [_p]),
    "orbit_engine_prepare_train": (_i, [_p, _p, _p, _p]),
    "orbit_engine_forward_train": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _i64, _p, _i64, _p]),
    "orbit_engine_backward_train": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _i64, _p]),
    "orbit_engine_film_grad": (_i, [_p, _p, _p, _p]),
    "orbit_film_generate_backward": (_i, [_p, _p, _i, _p, _i, _p, _p, _p, _p, _p]),
    "orbit_head_predict_backward": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _f, _p, _p]),
    "orbit_mahalanobis_predict_backward": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _p, _p]),
    "orbit_linear_ce_scratch_floats": (_i64, [_i, _i, _i]),
    "orbit_linear_ce_backward": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p, _p, _p, _p, _p]),
    "orbit_engine_profile_read": (_i, [_p, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]),
    "orbit_engine_last_launches": (_i64, [_p]),
}

_lib = None
_call_device = [None]   # device of the stream handed to the NEXT C call (set by stream_ptr, consumed by the call wrapper)


def _guarded(fn):
    """The kernels, cuTensorMapEncode and cudaFuncSetAttribute run in the CURRENT CUDA context, but the reference learners
    pick `cuda:N` through `--gpu N` without ever calling torch.cuda.set_device. Every C call takes the stream of the
    tensors' device (stream_ptr is evaluated as its last argument); the wrapper makes that device current for the call."""
    def call(*args):
        dev, _call_device[0] = _call_device[0], None
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args)
        with torch.cuda.device(dev):
            return fn(*args)
    call.__name__ = getattr(fn, '__name__', 'orbit_call')
    return call


class _Lib:
    """Attribute access returns the device-guarded ctypes functions."""

    def __init__(self, cdll, names):
        self._cdll = cdll
        for name in names:
            fn = getattr(cdll, name)
            setattr(self, name, _guarded(fn) if _SIGNATURES[name][1] and _SIGNATURES[name][1][-1] is _p else fn)


class OrbitError(RuntimeError):
    pass


def load():
    """Loads the shared library (once). Raises loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OrbitError(f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
                             "(orbit_b200 has no CPU/PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if lib.orbit_abi_version() != 1:
            raise OrbitError("liborbit_b200.so ABI version mismatch; rebuild")
        if lib.orbit_experiment_build() and not os.environ.get("ORBIT_ALLOW_EXPERIMENT_BUILD"):
            raise OrbitError("liborbit_b200.so was compiled with a timing-experiment flag (ORBIT_EXP_*: wrong results by design); "
                             "rebuild with `python __graft_entry__.py build`, or set ORBIT_ALLOW_EXPERIMENT_BUILD=1 to time it")
        _lib = _Lib(lib, list(_SIGNATURES))
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().orbit_error_string(int(rc)).decode()
        raise OrbitError(f"{what} failed with code {rc}: {msg}")


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise OrbitError(f"{name} must be a CUDA tensor: orbit_b200 runs only on an sm_100 GPU (no CPU fallback)")


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on ``device``; also tells the call wrapper which device to make current."""
    if device is not None:
        device = torch.device(device)
        _call_device[0] = device if device.type == 'cuda' else None
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_launches = 0  # kernels enqueued through this binding (bench.py reports it as gpu_launches)


def count_launches(n):
    global _launches
    _launches += int(n)


def launches():
    return _launches
