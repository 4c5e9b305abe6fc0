"""Builds libreinfocus_b200.so (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc.

The library is compiled for sm_100a only; there is no other backend and no fallback.
``python -m reinfocus_b200.build`` or ``__graft_entry__.build()`` run this; nvcc
cross-compiles without a GPU.
"""

import os
import shutil
import subprocess
import sys

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PACKAGE_DIR, "csrc")
LIB_NAME = "libreinfocus_b200.so"
LIB_PATH = os.path.join(PACKAGE_DIR, LIB_NAME)

SOURCES = ["rf_api.cu"]
HEADERS = ["rf_rng.cuh", "rf_tracer.cuh", "rf_tracer_mp.cuh", "rf_focus.cuh", "rf_generic.cuh", "rf_env.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # every FP op of the reference arithmetic is an explicit round-to-nearest intrinsic, which
    # nvcc never contracts; library math (the float64 sin of the general-scene tracer) keeps
    # the default contraction so that it matches libdevice as numba links it
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: reinfocus_b200 needs the CUDA toolkit to build")
    return nvcc


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(PACKAGE_DIR, "..", "include", "reinfocus_b200.h"))
    return any(os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    # the image exports CC/CXX=/opt/gcc wrappers that cannot link OpenMP; nvcc only needs g++
    result = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or result.returncode != 0:
        sys.stderr.write(result.stdout + result.stderr)
    if result.returncode != 0:
        raise RuntimeError(f"nvcc failed ({result.returncode}): {' '.join(cmd)}")
    with open(os.path.join(PACKAGE_DIR, "csrc", "ptxas_info.txt"), "w") as f:
        f.write(result.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
