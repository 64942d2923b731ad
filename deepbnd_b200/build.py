"""Build librve_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "rve_abi.cu")
DEPS = [SRC] + [os.path.join(HERE, "csrc", f) for f in ("rve_kernels.cuh", "rve_devsym.cuh", "rve_math.cuh", "rve_symbolic.hpp", "rve_matfree.cuh", "rve_stream.cuh")] + \
       [os.path.join(HERE, "..", "include", "rve_b200.h")]
LIB = os.environ.get("DEEPBND_RVE_LIB") or os.path.join(HERE, "lib", "librve_b200.so")   # override: kernel-variant experiments

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-pthread", "-shared"]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_lib(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("DEEPBND_NVCC_EXTRA", "").split()          # kernel-variant experiments: -DRVE_...=...
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building librve_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
