"""Builds liborbit_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so must travel with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liborbit_b200.so")
SOURCES = ["api.cu", "head.cu", "convnet.cu", "engine.cu", "gemm_tcgen05.cu", "gemm_stream.cu", "finetune.cu", "adapters.cu", "vit.cu", "mahalanobis.cu",
           "evaluator.cu", "train.cu", "train_setenc.cu", "mbconv_stream.cu", "conv_first.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


STAMP = LIB + ".flags"      # the nvcc flags the library on disk was built with (an experiment build must not pass for the shipped one)


def _flags():
    return " ".join(NVCC_FLAGS + os.environ.get("ORBIT_NVCC_EXTRA", "").split() + (["-lcuda"] if os.environ.get("ORBIT_LINK_LIBCUDA") else []))


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(STAMP) or open(STAMP).read() != _flags():
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "orbit_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("ORBIT_NVCC_EXTRA", "").split()      # e.g. -DORBIT_GEMM_TRACE for scripts/gemm_trace.py
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-lcuda"] if os.environ.get("ORBIT_LINK_LIBCUDA") else []
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building liborbit_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(STAMP, "w") as f:
        f.write(_flags())
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
