"""Builds the native libraries in-tree with nvcc for sm_100a (cross-compiles without a GPU):

  libcova_b200.so      the product: the hot path only
  libcova_b200_val.so  the same sources with -DCOVA_VALIDATION: adds the fp32 CUDA-core validation kernels
                       (COVA_IMPL_SIMT) the tests compare every tcgen05 layer against; never loaded by the product path
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcova_b200.so")
SO_VAL = os.path.join(HERE, "libcova_b200_val.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC,-O2,-Wall", "-cudart", "static", "--expt-relaxed-constexpr"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))


def _stale(so: str) -> bool:
    if not os.path.exists(so):
        return True
    t = os.path.getmtime(so)
    deps = sources() + [os.path.join(os.path.dirname(HERE), "include", "cova_b200.h")]
    return any(os.path.getmtime(f) > t for f in deps)


def needs_build() -> bool:
    return _stale(SO) or _stale(SO_VAL)


def build(force: bool = False, verbose: bool = False) -> str:
    units = [os.path.join(CSRC, "cova_abi.cu")] + [f for f in sources() if f.endswith(".cpp")]  # kernels + host-only C++
    jobs = []
    for so, extra in ((SO, []), (SO_VAL, ["-DCOVA_VALIDATION"])):
        if force or _stale(so):
            cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", so] + units
            jobs.append((so, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for so, proc in jobs:                                  # the two builds run side by side
        out, err = proc.communicate()
        if proc.returncode != 0:
            sys.stderr.write(out + err)
            raise RuntimeError(f"nvcc failed building {os.path.basename(so)}")
        if verbose:
            sys.stderr.write(err)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
