"""Builds litiv_b200/liblitiv_b200.so (CUDA kernels + C ABI) for sm_100a with nvcc. In-tree, no JIT cache."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "litiv_b200.cu")
SO = os.environ.get("LVB_SO") or os.path.join(HERE, "liblitiv_b200.so")  # LVB_SO: load an alternative build (tuning experiments)

# -fmad=false: the feedback arithmetic mirrors the reference's separate float mul/add (parity tier 3)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "550"]


def _sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in os.listdir(d)] + [os.path.join(ROOT, "include", "litiv_b200.h")]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, SRC]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
