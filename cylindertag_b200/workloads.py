"""The synthetic workloads of BASELINE.json configs 3-5 as SURVEY 8(d) writes them, in one place for the golden
generator (tests/golden/make_golden_ref.py), the parity tests and bench.py.  Input synthesis on the host (through
synth.py); not part of the detection path.

config 3: 256 frames 1920x1080 BGR, one 2f12c marker each, frame i from default_rng(1000 + i)
config 4: 3840x2160 BGR, 4..8 markers per frame (4 + i % 5), seeds 2000 + i, three codebooks: 2f12c (shipped) and
          15c3f / 18c4f generated with default_rng(7) (30 rows each; one detector holds one dictionary,
          header/CylinderTag.h:44)
config 5: the ring of 64 distinct config-4 2f12c frames (seeds 2000..2063)."""
import functools
import os

import numpy as np

from . import synth

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "data")
CODEBOOKS = {"2f12c": None, "15c3f": (15, 3), "18c4f": (18, 4)}
CONFIG3_FRAMES = 256
CONFIG4_FRAMES = 8


@functools.lru_cache(maxsize=None)
def codebook(name):
    """(state, feature_size) of a config-4 codebook."""
    if name == "2f12c":
        toks = open(os.path.join(DATA, "CTag_2f12c.marker")).read().split()
        n, cols, fs = int(toks[0]), int(toks[1]), int(toks[2])
        return np.array([int(t) for t in toks[3:3 + n * cols]], np.int32).reshape(n, cols), fs
    cols, fs = CODEBOOKS[name]
    return synth.generate_codebook(cols, fs, 30, seed=7), fs


def config3_frame(i):
    """Frame i of config 3 (BGR u8 1080x1920x3)."""
    state, _ = codebook("2f12c")
    return synth.synthetic_frame(1000 + i, 1920, 1080, state, 1, channels=3)[0]


def config4_markers(i):
    return 4 + i % 5


def config4_frame(name, i, w=3840, h=2160):
    """Frame i of the config-4 run with codebook `name`: (BGR frame, rendered dictionary rows)."""
    state, _ = codebook(name)
    frame, specs = synth.synthetic_frame(2000 + i, w, h, state, config4_markers(i), channels=3)
    return frame, [row for row, _ in specs]


CONFIG5_DISTINCT = 64


def config5_frame(i):
    """Frame i (0..63) of the config-5 ring: the config-4 2f12c frames with seeds 2000..2063."""
    return config4_frame("2f12c", i)[0]


def _job(args):
    kind, name, i = args
    cache = os.path.join(os.environ.get("CTAG_FRAME_CACHE", "/tmp"), f"ctag_frame_c{kind}_{name}_{i}.npy")
    if os.environ.get("CTAG_FRAME_CACHE") != "off" and os.path.exists(cache):
        try:
            return np.load(cache)
        except Exception:
            pass
    frame = config3_frame(i) if kind == 3 else config4_frame(name, i)[0]
    if os.environ.get("CTAG_FRAME_CACHE") != "off":
        try:
            tmp = cache + f".{os.getpid()}.tmp.npy"
            np.save(tmp, frame)
            os.replace(tmp, cache)
        except Exception:
            pass
    return frame


def render_many(jobs, workers=None):
    """jobs: [(3, None, i) | (4, codebook, i)] -> list of frames, rendered by a process pool (the renderer is numpy);
    rendered frames are cached as .npy under $CTAG_FRAME_CACHE (default /tmp; "off" disables)."""
    workers = workers or min(len(jobs), os.cpu_count() or 1)
    if workers <= 1 or len(jobs) <= 1:
        return [_job(j) for j in jobs]
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(workers) as pool:
        return pool.map(_job, jobs, chunksize=1)
