"""Synthetic CylinderTag inputs: codebook generator, marker geometry, cylinder renderer, .marker/.model writers.

The reference ships one photo and one dictionary; its MATLAB generator (CylinderTag_generator.m) writes bitmaps only.
The benchmark configs of BASELINE.json (1080p / 4K frames with rendered markers, 15c3f / 18c4f codebooks) therefore
need inputs made here.  Rules followed:
  * a state is 8*left + right, digits 0..7, legal iff both digits lie in the same half (<=3 or >=4)
    (CylinderTag_generator.m:18,96,114,164);
  * inverse(s) = (7 - s % 8) * 8 + (7 - s // 8)  (:198 and corner_detector.cpp:1299);
  * every cyclic window of `feature_size` states, read forward and read as flipped+inverted, is globally unique and no
    window equals its own inverse (:247-286, :27,179);
  * geometry of one column (:221-245): width W at pitch 1.5 W, height L, two black quads separated by a white band of
    height 0.2 L whose centre is p*L with p the root of -p^2 + p + (0.11 - 0.2 cr) = 0 (cross ratio cr of the digit).
This is host-side input synthesis (numpy/cv2); it is not part of the detection path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import cv2
import numpy as np

CR = (1.47, 1.54, 1.61, 1.68)


def legal_states():
    return [8 * a + b for a in range(8) for b in range(8) if (a <= 3) == (b <= 3)]


def inverse_state(s: int) -> int:
    return (7 - s % 8) * 8 + (7 - s // 8)


def _windows(row, f):
    n = len(row)
    fw = [tuple(row[(j + k) % n] for k in range(f)) for j in range(n)]
    inv = [tuple(inverse_state(row[(j - k) % n]) for k in range(f)) for j in range(n)]
    return fw, inv


def generate_codebook(cols: int, feature_size: int, rows: int, seed: int = 7) -> np.ndarray:
    """Random search for a dictionary satisfying the uniqueness rules above."""
    rng = np.random.default_rng(seed)
    states = legal_states()
    used = set()
    out = []
    attempts = 0
    while len(out) < rows:
        attempts += 1
        if attempts > 200000:
            raise RuntimeError("codebook search did not converge")
        row = [int(states[i]) for i in rng.integers(0, len(states), cols)]
        fw, inv = _windows(row, feature_size)
        allw = fw + inv
        if len(set(allw)) != len(allw):
            continue  # a window repeats inside the row or equals an inverse reading of the same row
        if any(w in used for w in allw):
            continue
        used.update(allw)
        out.append(row)
    return np.array(out, dtype=np.int32)


def generate_codebook_dfs(cols: int, feature_size: int, rows: int, seed: int = 7) -> np.ndarray:
    """The reference generator's depth-first search (CylinderTag_generator.m:34-216) in the library
    (ctag_generate_codebook, host code): reaches the full capacity, e.g. 41 rows for 2f12c like the shipped book."""
    import ctypes
    from . import _capi as C
    lib = C.load()
    cap = lib.ctag_codebook_capacity(cols, feature_size)
    if cap < 0:
        raise ValueError("cols / feature_size outside the generator's range")
    rows = min(rows, cap)
    out = np.zeros((rows, cols), np.int32)
    n = ctypes.c_int()
    C.check(lib.ctag_generate_codebook(cols, feature_size, rows, seed, out.ctypes.data_as(ctypes.c_void_p), rows, ctypes.byref(n)),
            "ctag_generate_codebook")
    return out[:n.value]


def check_codebook(state: np.ndarray, feature_size: int) -> bool:
    seen = set()
    for row in state.tolist():
        if any(s not in set(legal_states()) for s in row):
            return False
        fw, inv = _windows(row, feature_size)
        for w in fw + inv:
            if w in seen:
                return False
            seen.add(w)
    return True


def write_marker_file(path: str, state: np.ndarray, feature_size: int):
    """.marker format read by CylinderTag::load_from_file (CylinderTag.cpp:24-32)."""
    with open(path, "w") as fh:
        fh.write(f"{state.shape[0]} {state.shape[1]} {feature_size}\n")
        for row in state:
            fh.write("\t".join(str(int(v)) for v in row) + "\n")


def band_centre(digit: int) -> float:
    cr = CR[digit] if digit <= 3 else CR[7 - digit]
    disc = math.sqrt(1.0 + 4.0 * (0.11 - 0.2 * cr))
    return (1.0 - disc) / 2.0 if digit <= 3 else (1.0 + disc) / 2.0


@dataclass
class MarkerSpec:
    states: np.ndarray  # one dictionary row
    radius: float = 20.0  # mm
    ratio: float = 10.0  # L / W
    rvec: np.ndarray = field(default_factory=lambda: np.zeros(3))
    tvec: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 350.0]))
    black: float = 25.0
    white: float = 225.0

    @property
    def cols(self):
        return len(self.states)

    @property
    def W(self):
        return 2.0 * math.pi * self.radius / (1.5 * self.cols)

    @property
    def L(self):
        return self.ratio * self.W


def model_corners(spec: MarkerSpec) -> np.ndarray:
    """3-D corners [cols*8, 3] in the object frame, index 8*c + k, corner order of SURVEY D.4 / D.6."""
    W, L, r, n = spec.W, spec.L, spec.radius, spec.cols
    out = np.zeros((n * 8, 3))
    for c in range(n):
        s = int(spec.states[c])
        pl, pr = band_centre(s // 8) * L, band_centre(s % 8) * L
        x0, x1 = 1.5 * W * c, 1.5 * W * c + W
        uv = [(x0, 0), (x1, 0), (x1, pr - 0.1 * L), (x0, pl - 0.1 * L), (x1, L), (x0, L), (x0, pl + 0.1 * L), (x1, pr + 0.1 * L)]
        for k, (u, v) in enumerate(uv):
            th = 2.0 * math.pi * u / (1.5 * W * n)
            out[8 * c + k] = (r * math.sin(th), v - L / 2.0, -r * math.cos(th))
    return out


def write_model_file(path: str, specs_with_ids):
    """.model format read by CylinderTag::loadModel (CylinderTag.cpp:168-188)."""
    with open(path, "w") as fh:
        size = specs_with_ids[0][1].cols
        fh.write(f"{len(specs_with_ids)} {size}\n\n")
        for mid, spec in specs_with_ids:
            fh.write(f"{mid}\n0 0 0\n0 1 0\n")
            for i, p in enumerate(model_corners(spec)):
                fh.write(f"{i} {p[0]:.6f} {p[1]:.6f} {p[2]:.6f}\n")
            fh.write("\n")


def _texture(spec: MarkerSpec, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Analytic marker texture: value in {black, white}; u along the circumference, v along the axis (0..L)."""
    W, L, n = spec.W, spec.L, spec.cols
    pitch = 1.5 * W
    c = np.floor(u / pitch).astype(np.int64) % n
    x = u - np.floor(u / pitch) * pitch
    st = np.asarray(spec.states, np.int64)[c]
    pl = np.array([band_centre(d) for d in range(8)])[st // 8] * L
    pr = np.array([band_centre(d) for d in range(8)])[st % 8] * L
    centre = pl + (pr - pl) * np.clip(x / W, 0.0, 1.0)
    in_col = (x <= W) & (v >= 0) & (v <= L)
    black = in_col & ((v < centre - 0.1 * L) | (v > centre + 0.1 * L))
    return np.where(black, spec.black, spec.white)


def render_marker(img: np.ndarray, K: np.ndarray, spec: MarkerSpec, ss: int = 3):
    """Renders one marker wrapped once around a cylinder into img (float32, in place). Camera looks along +Z."""
    h, w = img.shape
    R, _ = cv2.Rodrigues(np.asarray(spec.rvec, np.float64).reshape(3, 1))
    t = np.asarray(spec.tvec, np.float64).reshape(3)
    L, r = spec.L, spec.radius
    margin = 0.15 * L
    # projected bounding box of the cylinder surface
    th = np.linspace(0, 2 * np.pi, 48)
    ys = np.array([-L / 2 - margin, L / 2 + margin])
    P = np.array([[r * np.sin(a), y, -r * np.cos(a)] for a in th for y in ys])
    Pc = P @ R.T + t
    if (Pc[:, 2] <= 1e-3).any():
        return
    px = Pc[:, 0] / Pc[:, 2] * K[0, 0] + K[0, 2]
    py = Pc[:, 1] / Pc[:, 2] * K[1, 1] + K[1, 2]
    x0, x1 = int(max(0, math.floor(px.min()) - 2)), int(min(w, math.ceil(px.max()) + 3))
    y0, y1 = int(max(0, math.floor(py.min()) - 2)), int(min(h, math.ceil(py.max()) + 3))
    if x1 <= x0 or y1 <= y0:
        return
    sub = (np.arange(ss) + 0.5) / ss - 0.5
    xs = (np.arange(x0, x1)[:, None] + sub[None, :]).reshape(-1)
    ysub = (np.arange(y0, y1)[:, None] + sub[None, :]).reshape(-1)
    X, Y = np.meshgrid(xs, ysub)
    d = np.stack([(X - K[0, 2]) / K[0, 0], (Y - K[1, 2]) / K[1, 1], np.ones_like(X)], -1)
    dO = d @ R  # R^T d
    oO = -(R.T @ t)
    a = dO[..., 0] ** 2 + dO[..., 2] ** 2
    b = 2 * (oO[0] * dO[..., 0] + oO[2] * dO[..., 2])
    c = oO[0] ** 2 + oO[2] ** 2 - r * r
    disc = b * b - 4 * a * c
    hit = disc > 0
    s = (-b - np.sqrt(np.where(hit, disc, 0))) / (2 * np.where(a > 0, a, 1))
    Pobj = oO + s[..., None] * dO
    v = Pobj[..., 1] + L / 2
    hit &= (s > 0) & (v >= -margin) & (v <= L + margin)
    theta = np.arctan2(Pobj[..., 0], -Pobj[..., 2])
    u = (theta % (2 * np.pi)) / (2 * np.pi) * (1.5 * spec.W * spec.cols)
    val = _texture(spec, u, v)
    hh, ww = y1 - y0, x1 - x0
    val = np.where(hit, val, 0.0).reshape(hh, ss, ww, ss)
    cov = hit.astype(np.float64).reshape(hh, ss, ww, ss)
    vs, cs = val.sum(axis=(1, 3)), cov.sum(axis=(1, 3))
    frac = cs / (ss * ss)
    region = img[y0:y1, x0:x1]
    region[:] = np.where(cs > 0, (vs / np.maximum(cs, 1)) * frac + region * (1 - frac), region)


def background(h: int, w: int, rng, lo=140.0, hi=220.0, sigma=12.0) -> np.ndarray:
    """Smooth texture: Gaussian-filtered uniform noise stretched to [lo, hi] (generated at 1/4 scale for speed)."""
    hs, ws = (h + 3) // 4, (w + 3) // 4
    n = rng.random((hs, ws)).astype(np.float32)
    n = cv2.GaussianBlur(n, (0, 0), sigma / 4.0)
    n = (n - n.min()) / max(float(n.max() - n.min()), 1e-6)
    n = cv2.resize(n, (w, h), interpolation=cv2.INTER_LINEAR)
    return lo + (hi - lo) * n


def camera_matrix(w: int, h: int, f: float | None = None) -> np.ndarray:
    f = f if f is not None else 2200.0 * w / 1920.0
    return np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1.0]])


def render_frame(w: int, h: int, specs, rng, blur_sigma=0.8, noise_sigma=2.0, K=None, channels=1) -> np.ndarray:
    K = camera_matrix(w, h) if K is None else K
    img = background(h, w, rng).astype(np.float64)
    # far markers first so that nearer ones occlude them
    for spec in sorted(specs, key=lambda s: -float(np.asarray(s.tvec)[2])):
        render_marker(img, K, spec)
    img = img.astype(np.float32)
    if blur_sigma > 0:
        img = cv2.GaussianBlur(img, (0, 0), blur_sigma)
    if noise_sigma > 0:
        img = img + rng.normal(0, noise_sigma, img.shape).astype(np.float32)
    gray = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if channels == 1:
        return gray
    # BGR frame whose cvtColor(BGR2GRAY) stays close to `gray`: small chroma offsets around the luminance
    off = rng.integers(-6, 7, (h, w, 3)).astype(np.int16)
    return np.clip(gray[..., None].astype(np.int16) + off, 0, 255).astype(np.uint8)


def random_specs(state: np.ndarray, n_markers: int, w: int, h: int, rng, K=None, px_width=(14.0, 40.0)):
    """Random cylinders spread over the frame: projected column width in px_width full-res px."""
    K = camera_matrix(w, h) if K is None else K
    f = K[0, 0]
    specs = []
    # place markers on a jittered grid so that they do not overlap
    gx = int(math.ceil(math.sqrt(n_markers * w / h)))
    gy = int(math.ceil(n_markers / gx))
    cells = [(i, j) for j in range(gy) for i in range(gx)]
    rng.shuffle(cells)
    for m in range(n_markers):
        i, j = cells[m]
        row = int(rng.integers(0, state.shape[0]))
        radius = float(rng.uniform(15, 40))
        ratio = float(rng.uniform(7, 12))
        cols = state.shape[1]
        W = 2 * math.pi * radius / (1.5 * cols)
        # the marker (height ratio*W) must fit its grid cell
        cell_h = h / gy
        wpx_max = min(px_width[1], 0.8 * cell_h / ratio)
        wpx = float(rng.uniform(min(px_width[0], wpx_max), wpx_max))
        z = f * W / wpx
        u = (i + rng.uniform(0.35, 0.65)) * w / gx
        v = (j + rng.uniform(0.4, 0.6)) * h / gy
        tx, ty = (u - K[0, 2]) / f * z, (v - K[1, 2]) / f * z
        rvec = np.array([rng.uniform(-0.35, 0.35), rng.uniform(-3, 3), rng.uniform(-0.5, 0.5)])
        specs.append((row, MarkerSpec(state[row], radius, ratio, rvec, np.array([tx, ty, z]),
                                      black=float(rng.uniform(15, 40)), white=float(rng.uniform(200, 240)))))
    return specs


def synthetic_frame(seed: int, w: int, h: int, state: np.ndarray, n_markers: int, channels: int = 1):
    """Frame `seed` of the synthetic sets: default_rng(seed); returns (frame, [(dictionary row, MarkerSpec)])."""
    rng = np.random.default_rng(seed)
    specs = random_specs(state, n_markers, w, h, rng)
    frame = render_frame(w, h, [s for _, s in specs], rng, blur_sigma=float(rng.uniform(0.3, 1.2)),
                         noise_sigma=float(rng.uniform(0.5, 3.0)), channels=channels)
    return frame, specs


def render_frames_gpu(det, frames_ptr: int, seeds, w: int, h: int, n_markers, channels: int = 3, pitch: int | None = None,
                      frame_stride: int = 0):
    """The same scenes as synthetic_frame(seed, ...) -- random_specs drawn from default_rng(seed), blur / noise levels from the
    same stream -- rendered by the library's CUDA renderer (ctag_render_frames) straight into device memory at `frames_ptr`
    (e.g. a torch tensor's data_ptr()).  n_markers: int or one int per frame.  Returns the rendered dictionary rows per frame.
    Pixel values differ from the host renderer's (other background texture / noise stream); the geometry is the same."""
    import ctypes
    from . import _capi as C
    state = det.state
    seeds = list(seeds)
    nm = [n_markers] * len(seeds) if isinstance(n_markers, int) else list(n_markers)
    K = camera_matrix(w, h)
    specs, start, fparams, truth = [], [0], [], []
    for seed, k in zip(seeds, nm):
        rng = np.random.default_rng(seed)
        sp = random_specs(state, k, w, h, rng, K)
        blur, noise = float(rng.uniform(0.3, 1.2)), float(rng.uniform(0.5, 3.0))
        for row, s in sorted(sp, key=lambda rs: -float(np.asarray(rs[1].tvec)[2])):  # far markers first
            specs.append(list(np.asarray(s.rvec, np.float64)) + list(np.asarray(s.tvec, np.float64)) +
                         [s.radius, s.ratio, s.black, s.white, float(row)] + [0.0] * 5)
        start.append(len(specs))
        truth.append([row for row, _ in sp])
        bits = np.array([seed * 2654435761 % (1 << 32), (seed * 40503 + 12345) % (1 << 32)], np.uint32).view(np.float32)
        fparams.append([K[0, 0], K[1, 1], K[0, 2], K[1, 2], blur, noise, bits[0], bits[1]])
    sp_arr = np.ascontiguousarray(np.array(specs, np.float32).reshape(-1, 16))
    st_arr = np.ascontiguousarray(np.array(start, np.int32))
    fp_arr = np.ascontiguousarray(np.array(fparams, np.float32))
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    pitch = pitch or w * channels
    C.check(C.load().ctag_render_frames(det._h, ctypes.c_void_p(frames_ptr), len(seeds), w, h, pitch, frame_stride, channels, vp(sp_arr),
                                        vp(st_arr), vp(fp_arr)), "ctag_render_frames")
    return truth


def video_sequence(gray: np.ndarray, n_frames: int = 120, seed: int = 2024, first: int = 0, count: int | None = None):
    """BASELINE.json config 2 substitute: the reference's test.avi is not shipped, so a deterministic
    `n_frames`-long sequence is made from test.bmp by seeded small homographies (rotation <= 3 deg, scale 0.95-1.05,
    shift <= 20 px, perspective terms <= 2e-5), Gaussian blur sigma in [0, 1] and noise sigma in [0, 3] DN, all drawn
    from ONE default_rng(seed) stream in frame order.  Returns frames [first, first + count) as a (count, h, w) u8 array
    (the draws of the skipped frames are still consumed, so frame i is the same whatever slice is asked for)."""
    h, w = gray.shape
    rng = np.random.default_rng(seed)
    count = n_frames - first if count is None else count
    out = []
    for i in range(min(n_frames, first + count)):
        ang = math.radians(rng.uniform(-3.0, 3.0))
        sc = rng.uniform(0.95, 1.05)
        tx, ty = rng.uniform(-20.0, 20.0, size=2)
        px, py = rng.uniform(-2e-5, 2e-5, size=2)
        blur = rng.uniform(0.0, 1.0)
        nsig = rng.uniform(0.0, 3.0)
        nseed = int(rng.integers(0, 2**31 - 1))
        if i < first:
            continue
        c, s = sc * math.cos(ang), sc * math.sin(ang)
        cx, cy = 0.5 * w, 0.5 * h
        H = np.array([[c, -s, cx - c * cx + s * cy + tx], [s, c, cy - s * cx - c * cy + ty], [px, py, 1.0 - px * cx - py * cy]])
        fr = cv2.warpPerspective(gray, H, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        if blur > 0.05:
            fr = cv2.GaussianBlur(fr, (0, 0), blur)
        noise = np.random.default_rng(nseed).normal(0.0, nsig, size=fr.shape)
        out.append(np.clip(np.rint(fr.astype(np.float64) + noise), 0, 255).astype(np.uint8))
    return np.stack(out)
