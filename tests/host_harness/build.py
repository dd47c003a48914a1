"""Builds the host logic harness (TEST INFRASTRUCTURE): the __host__ __device__ cores of the CUDA kernels compiled for
the CPU with g++, strict IEEE (no FMA contraction), so their sequential logic can be checked against the oracle on a
machine without a GPU.  It is never loaded by the product package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libctag_hostharness.so")
ROOT = os.path.join(HERE, "..", "..")


def build_harness(force=False):
    src = os.path.join(HERE, "harness.cpp")
    deps = [src] + [os.path.join(ROOT, "cylindertag_b200", "csrc", f) for f in
                    ("libm_core.cuh", "fit_core.cuh", "quad_core.cuh", "feature_core.cuh", "decode_core.cuh", "jpeg_core.cuh")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, src]
    subprocess.run(cmd, check=True)
    return LIB
