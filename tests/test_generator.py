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


def test_dfs_generator_reaches_the_capacity_of_the_shipped_book(marker_path):
    """CylinderTag_generator.m:36-39 caps a 2f12c book at 41 rows -- the size of the shipped CTag_2f12c.marker; the
    library's restatement of its depth-first search gets there, and the larger layouts give the 100 rows the .m asks for."""
    import ctypes
    from cylindertag_b200 import _capi as C
    lib = C.load()
    assert lib.ctag_codebook_capacity(12, 2) == 41 and lib.ctag_codebook_capacity(15, 3) == 1092
    shipped, fs = o.load_marker_file(marker_path)
    assert shipped.shape == (41, 12)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.ctag_check_codebook(vp(np.ascontiguousarray(shipped, np.int32)), 41, 12, 2) == 1
    for seed in (7, 8, 2024):
        book = synth.generate_codebook_dfs(12, 2, 41, seed=seed)
        assert book.shape == (41, 12) and synth.check_codebook(book, 2)
        assert lib.ctag_check_codebook(vp(np.ascontiguousarray(book, np.int32)), 41, 12, 2) == 1
        # 984 of the 992 usable windows are taken, like the shipped book
        fw = {(int(r[j]), int(r[(j + 1) % 12])) for r in book for j in range(12)}
        assert len(fw) == 41 * 12
    for cols, fsz in ((15, 3), (18, 4)):
        book = synth.generate_codebook_dfs(cols, fsz, 100, seed=7)
        assert book.shape == (100, cols) and synth.check_codebook(book, fsz)
    broken = synth.generate_codebook_dfs(12, 2, 10).copy()
    broken[1] = broken[0]
    assert lib.ctag_check_codebook(vp(np.ascontiguousarray(broken, np.int32)), 10, 12, 2) == 0
