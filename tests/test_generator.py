"""include/cylindertag/generator.h (dictionary generator, checker and .marker writer in C++) against the
Python rules in cylindertag_b200.synth and the oracle's .marker loader -- CPU only."""
import os
import subprocess

import numpy as np
import pytest

from cylindertag_b200 import synth
from oracle import ctag_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gen(tmp_path_factory):
    exe = tmp_path_factory.mktemp("gen") / "gen_test"
    libdir = os.path.join(ROOT, "cylindertag_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cxx", "test_generator.cpp"),
                    "-o", str(exe), "-L", libdir, "-lctag_b200", f"-Wl,-rpath,{libdir}"], check=True)
    return str(exe)


@pytest.mark.parametrize("cols,fsz,rows", [(12, 2, 20), (15, 3, 30), (18, 4, 30)])
def test_generated_dictionary_obeys_the_rules(gen, tmp_path, cols, fsz, rows):
    path = tmp_path / f"CTag_{fsz}f{cols}c.marker"
    out = subprocess.run([gen, str(cols), str(fsz), str(rows), str(path)], capture_output=True, text=True, check=True).stdout
    assert out.strip() == f"rows={rows} cols={cols} ok=1 broken_ok=0 illegal_ok=0 written=1"
    state, fs = o.load_marker_file(str(path))          # what CylinderTag::load_from_file would read
    assert fs == fsz and state.shape == (rows, cols)
    assert synth.check_codebook(np.asarray(state), fsz)  # the Python statement of the same rules agrees
    assert all((int(s) // 8 <= 3) == (int(s) % 8 <= 3) for s in np.asarray(state).ravel())


def test_shipped_dictionary_passes_the_cxx_checker_rules(marker_path):
    state, fs = o.load_marker_file(marker_path)
    assert synth.check_codebook(np.asarray(state), fs)
