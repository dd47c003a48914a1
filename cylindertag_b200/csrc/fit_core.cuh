// Line-fit arithmetic shared by the device kernels and by the host-side logic tests (tests/host_harness).
//
// Restates cv::fitLine for 2-D points (OpenCV imgproc linefit.cpp as shipped in the 4.x series; the reference calls it
// at corner_detector.cpp:136,151,163 with DIST_L2 and at :358 with DIST_WELSCH, param 0, reps = aeps = 0.01).
// The OpenCV source is not part of the reference tree; the arithmetic below follows SURVEY Appendix B.4 and is
// validated against cv2.fitLine in tests/test_fit_core.py.
//
// float/double placement matters and mirrors the library: points are float, moment sums are double, the angle is
// rounded to float before cos/sin, residuals and weights are float, their sums are double.
#pragma once
#include <math.h>
#include <stdint.h>

#include "libm_core.cuh"

#if defined(__CUDACC__)
#define CT_HD __host__ __device__ __forceinline__
#else
#define CT_HD inline
#endif

namespace ctag {
namespace core {

// ---- cv::RNG (multiply-with-carry), operations.hpp: RNG::next / RNG::uniform(int,int) -------------------------------
struct Rng {
  uint64_t state;
};
CT_HD uint32_t rng_next(Rng& r) {
  r.state = (uint64_t)(uint32_t)r.state * 4164903690U + (uint32_t)(r.state >> 32);
  return (uint32_t)r.state;
}
CT_HD int rng_uniform(Rng& r, int a, int b) { return a == b ? a : (int)(rng_next(r) % (uint32_t)(b - a) + a); }

// cos/sin/exp of a float argument: the library ends up in the C float routines; libm_core.cuh reproduces glibc's
// results bit for bit on host and device alike (see the note there on why that matters for DIST_WELSCH).
CT_HD float cos_f(float t) { return libm_cosf(t); }
CT_HD float sin_f(float t) { return libm_sinf(t); }
CT_HD float exp_f(float a) { return libm_expf(a); }

// ---- moments -> line, the tail of fitLine2D_wods ------------------------------------------------------------------
// x,y,x2,y2,xy are the (weighted) sums, w the weight sum.  line = (vx, vy, x0, y0).
CT_HD void line_from_moments(double x, double y, double x2, double y2, double xy, double w, float* line) {
  x /= w;
  y /= w;
  x2 /= w;
  y2 /= w;
  xy /= w;
  double dx2 = x2 - x * x;
  double dy2 = y2 - y * y;
  double dxy = xy - x * y;
  float t = (float)atan2(2 * dxy, dx2 - dy2) / 2;
  line[0] = cos_f(t);
  line[1] = sin_f(t);
  line[2] = (float)x;
  line[3] = (float)y;
}

// Unweighted fit over integer points: all sums are exact integers (|coord| < 4096, n < 2^20), so they can be kept
// incrementally as 64-bit integers and still reproduce the library's double accumulation bit for bit.
struct IntMoments {
  long long sx, sy, sxx, syy, sxy;
  int n;
};
CT_HD void im_reset(IntMoments& m) { m.sx = m.sy = m.sxx = m.syy = m.sxy = 0, m.n = 0; }
CT_HD void im_add(IntMoments& m, int x, int y) {
  m.sx += x;
  m.sy += y;
  m.sxx += (long long)x * x;
  m.syy += (long long)y * y;
  m.sxy += (long long)x * y;
  m.n += 1;
}
CT_HD void im_fit(const IntMoments& m, float* line) {
  line_from_moments((double)m.sx, (double)m.sy, (double)m.sxx, (double)m.syy, (double)m.sxy, (double)(float)m.n, line);
}

// packed point: x | y << 16 (absolute half-res coordinates, both < 4096)
CT_HD int pt_x(int p) { return p & 0xFFFF; }
CT_HD int pt_y(int p) { return (p >> 16) & 0xFFFF; }
// (float) of a packed coordinate (< 2^16).  On the device through the 2^23 trick (exact) instead of the conversion unit.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ float pt_xf(int p) { return __fsub_rn(__uint_as_float(0x4B000000u | (uint32_t)(p & 0xFFFF)), 8388608.0f); }
__device__ __forceinline__ float pt_yf(int p) { return __fsub_rn(__uint_as_float(0x4B000000u | ((uint32_t)p >> 16)), 8388608.0f); }
#else
CT_HD float pt_xf(int p) { return (float)pt_x(p); }
CT_HD float pt_yf(int p) { return (float)pt_y(p); }
#endif
CT_HD int pt_pack(int x, int y) { return x | (y << 16); }

CT_HD float welsch_exp(float a) { return exp_f(a); }

// One of the 20 random restarts of fitLine2D(DIST_WELSCH).  `pts` are the cluster points in library order.
// Visits at most 30 iterates; for each iterate i it reports err[i] and the line that produced it.
// Returns the number of iterates visited.  `stop_eps` > 0 makes the restart stop after the first iterate whose error is
// below it (the library's `err < EPS` early exit can only trigger then; see welsch_combine).
struct WelschIter {
  double err;
  float line[4];
};

// x % d for 32-bit x without the integer-division sequence: one double multiply and a +-1 correction (exact).
struct FastMod {
  uint32_t d;
  double inv;
};
CT_HD FastMod fastmod_make(int d) { return FastMod{(uint32_t)d, 1.0 / (double)d}; }
CT_HD uint32_t fastmod(uint32_t x, const FastMod& m) {
  uint32_t q = (uint32_t)((double)x * m.inv);
  int64_t r = (int64_t)x - (int64_t)q * m.d;
  if (r < 0) r += m.d;
  else if (r >= (int64_t)m.d) r -= m.d;
  return (uint32_t)r;
}

// Subset selection of one restart: draws indices until min(count,10) distinct ones are marked (the library's
// `w[j] < FLT_EPSILON` test on a zeroed weight array).  Advances the generator; returns the picks in draw order.
CT_HD int welsch_pick(Rng& rng, int count, const FastMod& fm, int* picked /*[10]*/) {
  const int need = count < 10 ? count : 10;
  int np = 0;
  if (count <= 64) {
    uint64_t seen = 0;
    while (np < need) {
      int j = (int)fastmod(rng_next(rng), fm);  // RNG::uniform(0, count) = next() % count
      uint64_t b = 1ull << j;
      if (!(seen & b)) {
        seen |= b;
        picked[np++] = j;
      }
    }
  } else {
    while (np < need) {
      int j = (int)fastmod(rng_next(rng), fm);
      bool dup = false;
      for (int q = 0; q < np; ++q) dup |= (picked[q] == j);
      if (!dup) picked[np++] = j;
    }
  }
  return np;
}

// Advances the generator over one restart's subset selection (so that restart k+1 can start from the right state).
CT_HD void welsch_skip_restart(Rng& rng, int count) {
  int picked[10];
  FastMod fm = fastmod_make(count);
  welsch_pick(rng, count, fm, picked);
}

// The generator is re-seeded with the same constant for every fitLine call, so the subset picked by restart k depends
// on the point count only.  welsch_pick_table fills table[(count-1)*200 + k*10 + a] with the picks of every restart for
// count = 1..max_count, sorted ascending (the order in which the library accumulates them).
CT_HD void welsch_sort_picks(int* picked, int np) {
  for (int a = 1; a < np; ++a) {
    int v = picked[a], b = a - 1;
    while (b >= 0 && picked[b] > v) {
      picked[b + 1] = picked[b];
      --b;
    }
    picked[b + 1] = v;
  }
}
inline void welsch_pick_table(uint16_t* table, int max_count) {
  for (int count = 1; count <= max_count; ++count) {
    Rng rng{0xFFFFFFFFFFFFFFFFull};
    FastMod fm = fastmod_make(count);
    for (int k = 0; k < 20; ++k) {
      int picked[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      int np = welsch_pick(rng, count, fm, picked);
      welsch_sort_picks(picked, np);
      for (int a = 0; a < 10; ++a) table[(size_t)(count - 1) * 200 + k * 10 + a] = (uint16_t)(a < np ? picked[a] : 0);
    }
  }
}

constexpr int kWelschCache = 128;  // raw weights kept per thread up to this cluster size

// One restart from its (sorted) initial subset; `visit(i, err, line)` is called for every iterate (at most 30).
template <typename PtFn, typename Visitor>
CT_HD int welsch_restart_from_picks(PtFn pt, int count, const int* picked, int np, Visitor& visit);

// One restart; the subset is drawn from the generator state at the start of the restart.
template <typename PtFn, typename Visitor>
CT_HD int welsch_restart_visit(PtFn pt, int count, Rng rng_at_restart, Visitor& visit) {
  Rng rng = rng_at_restart;
  int picked[10];
  FastMod fm = fastmod_make(count);
  int np = welsch_pick(rng, count, fm, picked);
  welsch_sort_picks(picked, np);  // the library accumulates over i = 0..count-1 with w[i] in {0,1}: ascending order
  return welsch_restart_from_picks(pt, count, picked, np, visit);
}

// State of one restart between iterations (lets a GPU lane interleave restarts of different fits).
struct WelschState {
  float line[4], prev[4];
  int i;     // next iteration
  int nvis;  // iterates reported so far
};

// Initial line of a restart: plain moments of its (sorted) subset.
template <typename PtFn>
CT_HD void welsch_init(PtFn pt, const int* picked, int np, WelschState& st) {
  double x = 0, y = 0, x2 = 0, y2 = 0, xy = 0, w = 0;
  for (int a = 0; a < np; ++a) {
    int p = pt(picked[a]);
    float px = pt_xf(p), py = pt_yf(p);
    x += px;
    y += py;
    x2 += px * px;
    y2 += py * py;
    xy += px * py;
    w += 1.0f;
  }
  line_from_moments(x, y, x2, y2, xy, w, st.line);
  st.prev[0] = st.prev[1] = st.prev[2] = st.prev[3] = 0.f;
  st.i = 0;
  st.nvis = 0;
}

// One iteration of the reweighting loop; returns false when the restart is over (converged or 30 iterations done).
// `wcache` holds kWelschCache floats of thread-private scratch.
template <typename PtFn, typename Visitor>
CT_HD bool welsch_step(PtFn pt, int count, WelschState& st, float* wcache, Visitor& visit) {
  float* line = st.line;
  float* prev = st.prev;
  if (st.i >= 30) return false;
  if (st.i > 0) {
    double t = line[0] * prev[0] + line[1] * prev[1];
    t = t > -1. ? t : -1.;
    t = t < 1. ? t : 1.;
    if (fabs(acos(t)) < 0.01f) {
      float dx = (float)fabs(line[2] - prev[2]);
      float dy = (float)fabs(line[3] - prev[3]);
      float d = dx > dy ? dx : dy;
      if (d < 0.01f) return false;
    }
  }
  const float c = 1 / 2.9846f;
  // residuals, error, raw weights.  For clusters of up to kWelschCache points the raw weights are kept (thread-local
  // array) so that the refit pass does not have to evaluate exp again; larger clusters recompute them.
  const float px0 = line[2], py0 = line[3], nx = line[1], ny = -line[0];
  double err = 0, sum_w = 0;
  const bool cached = count <= kWelschCache;
  // unrolled so that the (independent) exp evaluations of neighbouring points overlap; the accumulations keep the
  // library's order
#pragma unroll 4
  for (int j = 0; j < count; ++j) {
    int p = pt(j);
    float x = pt_xf(p) - px0, y = pt_yf(p) - py0;
    float r = (float)fabs(nx * x + ny * y);
    err += r;
    float wr = welsch_exp(-r * r * c * c);
    if (cached) wcache[j] = wr;
    sum_w += wr;
  }
  visit(st.nvis, err, line);
  ++st.nvis;
  // normalised weights + refit
  double x = 0, y = 0, x2 = 0, y2 = 0, xy = 0, wsum = 0;
  const bool norm = fabs(sum_w) > 1.1920928955078125e-07;
  const double inv = norm ? 1. / sum_w : 0.;
#pragma unroll 4
  for (int j = 0; j < count; ++j) {
    int p = pt(j);
    float fx = pt_xf(p), fy = pt_yf(p);
    float wr;
    if (cached) {
      wr = wcache[j];
    } else {
      float xx = fx - px0, yy = fy - py0;
      float r = (float)fabs(nx * xx + ny * yy);
      wr = welsch_exp(-r * r * c * c);
    }
    float w = norm ? (float)(wr * inv) : 1.f;
    x += w * fx;
    y += w * fy;
    x2 += w * fx * fx;
    y2 += w * fy * fy;
    xy += w * fx * fy;
    wsum += w;
  }
  prev[0] = line[0], prev[1] = line[1], prev[2] = line[2], prev[3] = line[3];
  line_from_moments(x, y, x2, y2, xy, wsum, line);
  ++st.i;
  return true;
}

template <typename PtFn, typename Visitor>
CT_HD int welsch_restart_from_picks(PtFn pt, int count, const int* picked, int np, Visitor& visit) {
  WelschState st;
  welsch_init(pt, picked, np, st);
  float wcache[kWelschCache];
  while (welsch_step(pt, count, st, wcache, visit)) {
  }
  return st.nvis;
}

// Visitor that stores the whole trajectory (exact library bookkeeping through welsch_combine).
struct WelschStore {
  WelschIter* out;
  int stride;
  CT_HD void operator()(int i, double err, const float* line) {
    WelschIter& o = out[i * stride];
    o.err = err;
    o.line[0] = line[0], o.line[1] = line[1], o.line[2] = line[2], o.line[3] = line[3];
  }
};

// Visitor that keeps only the restart's first minimum and whether any error fell below EPS.  When no restart of a fit
// sees a sub-EPS error the library's result is simply the first occurrence of the global minimum, so nothing else has
// to be stored; otherwise the caller recomputes with WelschStore and runs welsch_combine.
struct WelschBest {
  double err, eps;
  float line[4];
  bool sub_eps;
  CT_HD void operator()(int, double e, const float* l) {
    if (e < err) {
      err = e;
      line[0] = l[0], line[1] = l[1], line[2] = l[2], line[3] = l[3];
    }
    if (e < eps) sub_eps = true;
  }
};

template <typename PtFn>
CT_HD int welsch_restart(PtFn pt, int count, Rng rng_at_restart, WelschIter* out /*[30]*/, int out_stride, double) {
  WelschStore st{out, out_stride};
  return welsch_restart_visit(pt, count, rng_at_restart, st);
}

// Sequential bookkeeping of fitLine2D over the 20 restarts: `if (err < min_err) { keep; if (err < EPS) stop }`.
// iters[k*30*stride ...] / nvis[k] come from welsch_restart.  mode 0: a sub-EPS error stops everything;
// mode 1: it only ends the current restart.
CT_HD void welsch_combine(const WelschIter* iters, int stride_k, int stride_i, const int* nvis, int nvis_stride, int count,
                          int mode, float* line) {
  const double EPS = count * 1.1920928955078125e-07;
  double min_err = 1.7976931348623157e308;
  line[0] = line[1] = line[2] = line[3] = 0.f;
  for (int k = 0; k < 20; ++k) {
    int nv = nvis[k * nvis_stride];
    bool stop_all = false;
    for (int i = 0; i < nv; ++i) {
      const WelschIter& it = iters[k * stride_k + i * stride_i];
      if (it.err < min_err) {
        min_err = it.err;
        line[0] = it.line[0], line[1] = it.line[1], line[2] = it.line[2], line[3] = it.line[3];
        if (it.err < EPS) {
          stop_all = (mode == 0);
          break;
        }
      }
    }
    if (stop_all) break;
  }
}

}  // namespace core
}  // namespace ctag
