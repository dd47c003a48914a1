"""Builds the C++ CPU restatement (oracle/cpu_ref/cpu_ref.cpp) with the flags of the reference's last build:
-O3 -DNDEBUG, no -march, no -ffast-math (build/CMakeCache.txt:77); FMA contraction is off."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libctag_cpu_ref.so")
ROOT = os.path.join(HERE, "..", "..")


def build_cpu_ref(force=False):
    src = os.path.join(HERE, "cpu_ref.cpp")
    deps = [src] + [os.path.join(ROOT, "cylindertag_b200", "csrc", f) for f in
                    ("libm_core.cuh", "fit_core.cuh", "quad_core.cuh", "feature_core.cuh", "decode_core.cuh")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    subprocess.run(["g++", "-O3", "-DNDEBUG", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-o", LIB, src],
                   check=True)
    return LIB
