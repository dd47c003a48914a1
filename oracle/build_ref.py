"""TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/libctag_ref.so: the reference's OWN sources, compiled unmodified from
where they lie (/root/reference/corner_detector.cpp, CylinderTag.cpp, pose_estimation.cpp), against the OpenCV / Ceres
stand-ins in oracle/ref_shim/ and the ctypes driver oracle/ref_shim/ref_capi.cpp.  Flags follow the reference's last
build (build/CMakeCache.txt:51,57,77: g++, Release, -O3 -DNDEBUG; CMakeLists.txt:4-6: C++17 without extensions; no
-march, no -ffast-math, so no FMA contraction).  The reference's own build system is not run (it needs cmake
find_package(OpenCV) / find_package(Ceres), which this container cannot satisfy).  Outputs go to oracle/_ref/ only
(git-ignored; it travels to the GPU box with the snapshot, where /root/reference does not exist and the prebuilt
library is used as is)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "ref_shim")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libctag_ref.so")
REFERENCE = os.environ.get("CTAG_REFERENCE_DIR", "/root/reference")
REF_SOURCES = ["corner_detector.cpp", "CylinderTag.cpp", "pose_estimation.cpp"]
SHIM_SOURCES = ["shim_cv.cpp", "shim_ceres.cpp", "ref_capi.cpp"]
CXXFLAGS = ["-O3", "-DNDEBUG", "-std=c++17", "-ffp-contract=off", "-fPIC", "-fvisibility=hidden", "-pthread", "-w"]


def reference_present():
    return all(os.path.exists(os.path.join(REFERENCE, s)) for s in REF_SOURCES)


def _deps():
    d = [os.path.join(SHIM, s) for s in SHIM_SOURCES]
    for root, _, files in os.walk(SHIM):
        d += [os.path.join(root, f) for f in files]
    if reference_present():
        d += [os.path.join(REFERENCE, s) for s in REF_SOURCES]
        d += [os.path.join(REFERENCE, "header", f) for f in os.listdir(os.path.join(REFERENCE, "header"))]
    return d


def build_ref(force=False):
    """Returns the library path.  Rebuilds when the reference is present and anything is newer than the library;
    without the reference (GPU box) the prebuilt library is returned, and its absence is an error."""
    if not reference_present():
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("oracle/_ref/libctag_ref.so is missing and %s is not available to build it from" % REFERENCE)
    if not force and os.path.exists(LIB) and all(os.path.getmtime(p) <= os.path.getmtime(LIB) for p in _deps()):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    objs = []
    jobs = []
    for src in [os.path.join(REFERENCE, s) for s in REF_SOURCES] + [os.path.join(SHIM, s) for s in SHIM_SOURCES]:
        obj = os.path.join(OUT, os.path.basename(src) + ".o")
        cmd = ["g++"] + CXXFLAGS + ["-I", SHIM, "-I", REFERENCE, "-I", os.path.join(HERE, "..", "include"), "-c", src, "-o", obj]
        jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in jobs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("g++ failed on " + src)
    subprocess.run(["g++", "-shared", "-pthread", "-o", LIB] + objs, check=True)
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
