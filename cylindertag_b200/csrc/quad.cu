// K4: quad extraction (reference row a5, corner_detector.cpp:171-405) as three kernels over device work lists:
//   quad_edges_kernel  one warp per legal component: silhouettes, oriented trace, extended RDP -> four clusters
//   quad_fit_kernel    one thread per (component, edge, restart) of cv::fitLine(DIST_WELSCH) + merge / exact fallback
//   quad_select_kernel one thread per component: six intersections, best 4-subset
// plus the ordered compaction of the surviving quads.  The per-component arithmetic lives in quad_core.cuh /
// fit_core.cuh (shared with the host logic tests).
#include "common.cuh"
#include "kernels.cuh"
#include <cstdlib>
#include "quad_core.cuh"

namespace ctag {

using namespace core;

// Per-warp global scratch layout (bytes) for components too large for shared memory; depends on the geometry only.
struct QuadScratchLayout {
  size_t vis, col_top, col_bot, pts_a, pts_b, stack, cl, rng, iters, nvis, lines, total;
};

static inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

static QuadScratchLayout make_layout(const FrameGeom& g) {
  QuadScratchLayout L;
  const size_t pmax = 2 * (size_t)(g.hw + g.hh) + 16;
  size_t o = 0;
  L.vis = o;
  o = align16(o + 4 * ((size_t)((g.hw + 31) / 32) * g.hh + 4));
  L.col_top = o;
  o = align16(o + 2 * (size_t)g.hw);
  L.col_bot = o;
  o = align16(o + 2 * (size_t)g.hw);
  L.pts_a = o;
  o = align16(o + 4 * pmax);
  L.pts_b = o;
  o = align16(o + 4 * pmax);
  L.stack = o;
  o = align16(o + 4 * pmax);
  L.cl = o;
  o = align16(o + 4 * pmax);
  L.rng = o;
  o = align16(o + 8 * 80);
  L.iters = o;
  o = align16(o + sizeof(WelschIter) * 80 * 30);
  L.nvis = o;
  o = align16(o + 4 * 80);
  L.lines = o;
  o = align16(o + 4 * 16);
  L.total = (o + 255) & ~(size_t)255;
  return L;
}

size_t quad_scratch_bytes_per_warp(const FrameGeom& g) { return make_layout(g).total; }  // per persistent CTA

// control words of the quad stage (device ints)
enum { QC_WORK_EDGES = 0, QC_EXACT_COUNT = 1, QC_FIT_COUNT = 2, QC_POOL_CURSOR = 3, QC_OVERFLOW = 4, QC_WORK_FIT = 5,
       QC_UNITS_G32 = 6, QC_TAIL_COUNT = 7, QC_WORK_G32 = 8, QC_FIT_FALLBACKS = 10, QC_WORDS = 16 };

// Restarts of clusters with more than kGroup32Min points are fitted by a whole warp each (quad_fitwarp_kernel), the rest
// by one lane each (quad_fit_kernel); once the work list has run dry the lane kernel parks its stragglers in a tail list
// and the warp kernel finishes them, one restart per warp.
constexpr int kGroup32Min = 256;
constexpr int kTailLanes = 4;  // a warp parks its restarts when the list is dry and at most this many lanes are busy

// A restart handed from the lane kernel to the warp kernel between two reweighting passes.
struct TailRec {
  float line[4], prev[4], best_line[4];
  double best_err;
  int i, nvis, t, count, sub_eps, pool_index;
};

// Component that reached four edges: what the line fits and the corner selection need.
struct FitRec {
  int o;          // frame * legal_cap + component index
  int area;       // pixel count of the component
  int pool_off;   // first cluster point in the point pool
  int cl_off[5];  // cluster c = pool[pool_off + cl_off[c] .. pool_off + cl_off[c+1])
  float cx, cy;   // area_center
  int cols, rows;
};

// prefix[f] = number of legal components in frames < f; prefix[n] = total.  Also resets the control words.
// frame_fit[f] = four-edge components of frame f so far, frame_fit[n + f] = 1 when frame f ran out of its fit quota.
__global__ void quad_prefix_kernel(const int* __restrict__ counters, int n, int* __restrict__ prefix, int* __restrict__ qctl,
                                   int* __restrict__ frame_fit) {
  for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) frame_fit[i] = 0;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int f = 0; f < n; ++f) {
      prefix[f] = acc;
      acc += counters[f * 4 + 1];
    }
    prefix[n] = acc;
  }
  if (threadIdx.x < QC_WORDS) qctl[threadIdx.x] = 0;
}

// ---- K4a: edge stage, one warp per legal component ------------------------------------------------------------------
// Persistent warps pull components from a global counter.  Small components keep the boundary bit map and the point
// lists in shared memory; large ones use the warp's global scratch (same code, different pointers).  Components with
// four edges append a FitRec and copy their clusters to the point pool.
constexpr int kEdgeWarps = 4;
constexpr int kEdgeCtasPerSm = 6;  // 80 registers per thread: at 8 CTAs (64 registers) the trace loop spills and rematerialises pointers
constexpr int kSmemPts = 256;       // points per list on the shared-memory fast path
constexpr int kSmemVisWords = 256;  // bit-map words on the shared-memory fast path

__global__ void __launch_bounds__(32 * kEdgeWarps, kEdgeCtasPerSm) quad_edges_kernel(
    int n_frames, FrameGeom g, const uint8_t* __restrict__ bin, size_t bin_fstride, const int* __restrict__ labels,
    const int* __restrict__ legal, int legal_cap, const int* __restrict__ prefix, int* __restrict__ qctl,
    uint8_t* __restrict__ scratch, QuadScratchLayout L, int* __restrict__ quad_status, FitRec* __restrict__ fits, int fit_cap,
    int fit_per_frame, int* __restrict__ frame_fit, int* __restrict__ pool, int pool_cap) {
  __shared__ uint32_t s_vis[kEdgeWarps][kSmemVisWords];
  __shared__ int s_pts_a[kEdgeWarps][kSmemPts + 8], s_pts_b[kEdgeWarps][kSmemPts + 8], s_stack[kEdgeWarps][kSmemPts + 8],
      s_cl[kEdgeWarps][kSmemPts + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warp_global = blockIdx.x * kEdgeWarps + warp;
  uint8_t* base = scratch + (size_t)warp_global * L.total;
  QuadScratch gsc;
  gsc.vis = reinterpret_cast<uint32_t*>(base + L.vis);
  gsc.col_top = reinterpret_cast<int16_t*>(base + L.col_top);
  gsc.col_bot = reinterpret_cast<int16_t*>(base + L.col_bot);
  gsc.pts_a = reinterpret_cast<int*>(base + L.pts_a);
  gsc.pts_b = reinterpret_cast<int*>(base + L.pts_b);
  gsc.stack = reinterpret_cast<int*>(base + L.stack);
  gsc.cl = reinterpret_cast<int*>(base + L.cl);
  gsc.rng = nullptr;
  gsc.iters = nullptr;
  gsc.nvis = nullptr;
  gsc.lines = nullptr;
  const int total = prefix[n_frames];
  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(&qctl[QC_WORK_EDGES], 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    int lo = 0, hi = n_frames - 1;  // frame of this item: largest f with prefix[f] <= item
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= item) lo = mid;
      else hi = mid - 1;
    }
    const int fr = lo, ci = item - prefix[fr];
    const int* lg = legal + ((size_t)fr * legal_cap + ci) * 6;
    CompView cv;
    cv.bin = bin + (size_t)fr * bin_fstride;
    cv.bpitch = g.bpitch;
    cv.labels = labels + (size_t)fr * g.nblocks;
    cv.bw = g.bw;
    cv.cols = g.hw;
    cv.rows = g.hh;
    cv.root = lg[0];
    cv.area = lg[1];
    cv.x0 = lg[2];
    cv.y0 = lg[3];
    cv.x1 = lg[4];
    cv.y1 = lg[5];
    const int bw_ = cv.x1 - cv.x0 + 1, bh_ = cv.y1 - cv.y0 + 1;
    QuadScratch sc = gsc;
    if (((bw_ + 31) >> 5) * bh_ <= kSmemVisWords) sc.vis = s_vis[warp];
    if (2 * (bw_ + bh_) <= kSmemPts) {  // boundary points <= silhouette pixels <= 2 * (w + h)
      sc.pts_a = s_pts_a[warp];
      sc.pts_b = s_pts_b[warp];
      sc.stack = s_stack[warp];
      sc.cl = s_cl[warp];
    }
    QuadEdges ed;
    quad_stage_edges(cv, sc, Lanes{lane, 32}, &ed);
    const int o = fr * legal_cap + ci;
    int slot = -1, poff = 0;
    if (ed.cnt == 4) {
      if (lane == 0) {
        // every frame owns a quota of fit slots (the batch capacity is the sum of the quotas), so a cluttered frame
        // can only exhaust -- and flag -- itself, never a neighbour in the batch
        if (atomicAdd(&frame_fit[fr], 1) >= fit_per_frame) {
          frame_fit[n_frames + fr] = 1;
          slot = -2;
        } else {
          slot = atomicAdd(&qctl[QC_FIT_COUNT], 1);
          poff = atomicAdd(&qctl[QC_POOL_CURSOR], ed.cl_off[4]);
          if (slot >= fit_cap || poff + ed.cl_off[4] > pool_cap) {  // cannot happen with the sizes api.cu allocates
            atomicExch(&qctl[QC_OVERFLOW], 1);
            slot = -2;
          }
        }
      }
      slot = __shfl_sync(0xffffffffu, slot, 0);
      poff = __shfl_sync(0xffffffffu, poff, 0);
    }
    if (slot >= 0) {
      for (int i = lane; i < ed.cl_off[4]; i += 32) pool[poff + i] = sc.cl[i];
      if (lane == 0) {
        FitRec r;
        r.o = o;
        r.area = cv.area;
        r.pool_off = poff;
        for (int c = 0; c < 5; ++c) r.cl_off[c] = ed.cl_off[c];
        r.cx = ed.cx;
        r.cy = ed.cy;
        r.cols = cv.cols;
        r.rows = cv.rows;
        fits[slot] = r;
      }
    } else if (lane == 0) {
      quad_status[o] = Q_FEW_EDGES;  // also the fate of a component dropped because a pool overflowed (frame flagged)
    }
    __syncwarp();
  }
}

// ---- K4b: DIST_WELSCH fits (cv::fitLine, corner_detector.cpp:358) -----------------------------------------------------
// quad_fit_kernel     one THREAD per (component, edge, restart): the 80 restarts of a component are independent given
//                     their initial subsets, which depend on the point count only and come from a table built once
//                     per detector.  Each thread reports its first minimum and whether any error fell below EPS.
// quad_fitmerge_kernel one thread per (component, edge): without a sub-EPS error the library result is the first
//                     occurrence of the minimum (lowest restart on ties); otherwise the edge is queued for ...
// quad_fitexact_kernel one warp per queued edge: restarts replayed with stored trajectories, then the library's exact
//                     sequential bookkeeping (welsch_combine).
struct FitResult {
  double err;
  float line[4];
  int sub_eps;
  int pad;
};

struct PoolPts {
  const int* p;
  __device__ __forceinline__ int operator()(int j) const { return p[j]; }
};

__device__ __forceinline__ int load_picks(const uint16_t* __restrict__ table, int table_max, int count, int k, int* picked) {
  const int np = count < 10 ? count : 10;
  if (count <= table_max) {
    const uint16_t* t = table + (size_t)(count - 1) * 200 + k * 10;
    for (int a = 0; a < np; ++a) picked[a] = t[a];
  } else {
    // beyond the table: replay the generator from the seed (rare: > table_max boundary points on one edge)
    Rng rng{0xFFFFFFFFFFFFFFFFull};
    FastMod fm = fastmod_make(count);
    for (int q = 0; q <= k; ++q) welsch_pick(rng, count, fm, picked);
    welsch_sort_picks(picked, np);
  }
  return np;
}

// Fit edges sorted by point count, largest first (counting sort, one CTA): the lanes of a warp then walk clusters of
// similar size, and the long restarts start first instead of forming the tail of the fit kernel.
__global__ void __launch_bounds__(1024) quad_fitorder_kernel(int* __restrict__ qctl, const FitRec* __restrict__ fits,
                                                              int fit_cap, int* __restrict__ order) {
  __shared__ int hist[1024];
  __shared__ int wsum[32];
  int nfit = qctl[QC_FIT_COUNT];
  if (nfit > fit_cap) nfit = fit_cap;
  const int nedge = 4 * nfit, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  hist[tid] = 0;
  __syncthreads();
  for (int e = tid; e < nedge; e += 1024) {
    const FitRec* fr = fits + (e >> 2);
    const int cnt = fr->cl_off[(e & 3) + 1] - fr->cl_off[e & 3];
    atomicAdd(&hist[1023 - (cnt < 1023 ? cnt : 1023)], 1);
  }
  __syncthreads();
  // exclusive scan of the 1024 buckets
  const int v = hist[tid];
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = wsum[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += u;
    }
    wsum[lane] = winc - w;
  }
  __syncthreads();
  hist[tid] = wsum[warp] + inc - v;
  __syncthreads();
  // the sorted list starts with the largest clusters: their 20 restarts each are the work units of
  // quad_fitwarp_kernel, the lane-per-restart kernel starts behind them
  if (tid == 0) {
    const int u32 = 20 * hist[1023 - kGroup32Min];
    qctl[QC_UNITS_G32] = u32;  // work units [0, u32): one restart per warp (quad_fitwarp_kernel)
    qctl[QC_WORK_G32] = 0;
    qctl[QC_TAIL_COUNT] = 0;
    qctl[QC_WORK_FIT] = u32;   // [u32, 80 * nfit): one restart per lane (quad_fit_kernel)
  }
  __syncthreads();
  for (int e = tid; e < nedge; e += 1024) {
    const FitRec* fr = fits + (e >> 2);
    const int cnt = fr->cl_off[(e & 3) + 1] - fr->cl_off[e & 3];
    order[atomicAdd(&hist[1023 - (cnt < 1023 ? cnt : 1023)], 1)] = e;
  }
}

// ---- larger clusters: one restart per GROUP of G lanes ---------------------------------------------------------------
// A lane that walks a 100-point cluster through 30 reweighting passes alone needs about 600k cycles and is the tail of
// the whole quad stage, so the restarts of larger clusters are spread over G = 8 or 32 lanes: residuals, weights and
// products per point in parallel, the sums by shuffle.  cv::fitLine accumulates its sums sequentially in double, and the
// result has to stay bit-identical, so a parallel sum is only accepted when it is provably EXACT: every term is a
// non-negative float, hence a multiple of ulp(smallest non-zero term) >= t_min * 2^-24, and every partial sum in any order
// is a multiple of that quantum bounded by the total S; with S <= 2^28 * t_min all of them fit the 53-bit significand,
// no addition rounds, and any summation order gives the same double.  An iteration in which one of its eight sums
// fails that test (a point on the line to within float noise next to far outliers) is redone by the scalar routine,
// redundantly on every lane of the group (counted in QC_FIT_FALLBACKS).
struct GroupSum {
  double s;
  float tmin;  // smallest non-zero term
};
__device__ __forceinline__ void gsum_add(GroupSum& a, float t) {
  a.s += t;
  if (t > 0.f && t < a.tmin) a.tmin = t;
}
template <int G>
__device__ __forceinline__ bool gsum_reduce(GroupSum& a) {  // every lane of the group gets the total; true = provably exact
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    a.s += __shfl_xor_sync(0xffffffffu, a.s, o);
    a.tmin = fminf(a.tmin, __shfl_xor_sync(0xffffffffu, a.tmin, o));
  }
  return a.s <= 268435456.0 * (double)a.tmin;  // 2^28 * t_min (an all-zero sum has tmin = inf and is exact)
}

// One reweighting pass of the group's restart (welsch_step, fit_core.cuh, with the point loops dealt over the lanes).
// Called by ALL lanes of the warp (the reductions are warp-wide shuffles); groups without work pass have = false.
// Returns false when the restart is over.
template <int G, int CACHE>
__device__ __forceinline__ bool welsch_step_group(PoolPts pt, int count, bool have, WelschState& st, WelschBest& best, int gl,
                                                  int* __restrict__ fallbacks) {
  float* line = st.line;
  float* prev = st.prev;
  bool run = have && st.i < 30;
  if (run && st.i > 0) {
    double t = line[0] * prev[0] + line[1] * prev[1];
    t = t > -1. ? t : -1.;
    t = t < 1. ? t : 1.;
    if (fabs(acos(t)) < 0.01f) {
      float dx = (float)fabs(line[2] - prev[2]);
      float dy = (float)fabs(line[3] - prev[3]);
      float d = dx > dy ? dx : dy;
      if (d < 0.01f) run = false;
    }
  }
  const int n = run ? count : 0;
  const float c = 1 / 2.9846f;
  const float px0 = line[2], py0 = line[3], nx = line[1], ny = -line[0];
  const float kInf = __int_as_float(0x7f800000);
  float wloc[CACHE];
  const bool cached = count <= G * CACHE;
  GroupSum s_err{0., kInf}, s_w{0., kInf};
  for (int j = gl, m = 0; j < n; j += G, ++m) {
    const int p = pt(j);
    const float x = pt_xf(p) - px0, y = pt_yf(p) - py0;
    const float r = (float)fabs(nx * x + ny * y);
    const float wr = welsch_exp(-r * r * c * c);
    if (cached) wloc[m] = wr;
    gsum_add(s_err, r);
    gsum_add(s_w, wr);
  }
  bool exact = gsum_reduce<G>(s_err);
  exact &= gsum_reduce<G>(s_w);
  const double sum_w = s_w.s;
  const bool norm = fabs(sum_w) > 1.1920928955078125e-07;
  const double inv = norm ? 1. / sum_w : 0.;
  GroupSum sx{0., kInf}, sy{0., kInf}, sxx{0., kInf}, syy{0., kInf}, sxy{0., kInf}, sw{0., kInf};
  if (exact) {
    for (int j = gl, m = 0; j < n; j += G, ++m) {
      const int p = pt(j);
      const float fx = pt_xf(p), fy = pt_yf(p);
      float wr;
      if (cached) {
        wr = wloc[m];
      } else {
        const float xx = fx - px0, yy = fy - py0;
        const float r = (float)fabs(nx * xx + ny * yy);
        wr = welsch_exp(-r * r * c * c);
      }
      const float w = norm ? (float)(wr * inv) : 1.f;
      gsum_add(sx, w * fx);
      gsum_add(sy, w * fy);
      gsum_add(sxx, w * fx * fx);
      gsum_add(syy, w * fy * fy);
      gsum_add(sxy, w * fx * fy);
      gsum_add(sw, w);
    }
  }
  exact &= gsum_reduce<G>(sx);
  exact &= gsum_reduce<G>(sy);
  exact &= gsum_reduce<G>(sxx);
  exact &= gsum_reduce<G>(syy);
  exact &= gsum_reduce<G>(sxy);
  exact &= gsum_reduce<G>(sw);
  if (!run) return false;
  if (!exact) {
    if (gl == 0) atomicAdd(fallbacks, 1);
    float wc[kWelschCache];
    return welsch_step(pt, count, st, wc, best);
  }
  best(st.nvis, s_err.s, line);
  ++st.nvis;
  prev[0] = line[0], prev[1] = line[1], prev[2] = line[2], prev[3] = line[3];
  line_from_moments(sx.s, sy.s, sxx.s, syy.s, sxy.s, sw.s, line);
  ++st.i;
  return true;
}

// One restart per warp: the fresh restarts of the largest clusters (work units [0, qctl[QC_UNITS_G32]) of the size-sorted
// list) and the stragglers the lane kernel parked in the tail list.  Runs after quad_fit_kernel.
__global__ void __launch_bounds__(128) quad_fitwarp_kernel(int* __restrict__ qctl, const FitRec* __restrict__ fits,
                                                           const int* __restrict__ pool, const uint16_t* __restrict__ pick_table,
                                                           int table_max, const int* __restrict__ order,
                                                           const TailRec* __restrict__ tails, int tail_cap,
                                                           FitResult* __restrict__ results) {
  const int lane = threadIdx.x & 31;
  const int fresh = qctl[QC_UNITS_G32];
  int ntail = qctl[QC_TAIL_COUNT];
  if (ntail > tail_cap) ntail = tail_cap;
  while (true) {
    int u = 0;
    if (lane == 0) u = atomicAdd(&qctl[QC_WORK_G32], 1);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= fresh + ntail) break;
    WelschState st;
    WelschBest best;
    PoolPts pa{pool};
    int t, count;
    if (u < fresh) {
      const int ei = u / 20, k = u - 20 * ei, e = order[ei], slot = e >> 2, c = e & 3;
      t = 20 * e + k;
      const FitRec* fr = fits + slot;
      count = fr->cl_off[c + 1] - fr->cl_off[c];
      best.err = 1.7976931348623157e308;
      best.sub_eps = false;
      best.line[0] = best.line[1] = best.line[2] = best.line[3] = 0.f;
      int picked[10];
      const int np = load_picks(pick_table, table_max, count, k, picked);
      pa.p = pool + fr->pool_off + fr->cl_off[c];
      welsch_init(pa, picked, np, st);  // ten points: every lane computes the same initial line
    } else {
      const TailRec r = tails[u - fresh];
#pragma unroll
      for (int q = 0; q < 4; ++q) st.line[q] = r.line[q], st.prev[q] = r.prev[q], best.line[q] = r.best_line[q];
      st.i = r.i;
      st.nvis = r.nvis;
      best.err = r.best_err;
      best.sub_eps = r.sub_eps != 0;
      t = r.t;
      count = r.count;
      pa.p = pool + r.pool_index;
    }
    best.eps = count * 1.1920928955078125e-07;
    while (welsch_step_group<32, 32>(pa, count, true, st, best, lane, &qctl[QC_FIT_FALLBACKS])) {
    }
    if (lane == 0) {
      FitResult o;
      o.err = best.err;
      o.line[0] = best.line[0], o.line[1] = best.line[1], o.line[2] = best.line[2], o.line[3] = best.line[3];
      o.sub_eps = best.sub_eps ? 1 : 0;
      o.pad = 0;
      results[t] = o;
    }
  }
}

// Restarts converge after anything between 2 and 30 iterations, and fits with at most 10 points run a single restart:
// a lane that handled one restart from start to end would idle most of the time.  Instead every lane keeps the state of
// its current restart, all lanes advance by one iteration per trip, and a lane whose restart is over takes the next one
// from the global counter right away.
// (no minimum-blocks bound: 56 registers, eight CTAs per SM; forcing 12 or 16 CTAs per SM spills and is 13-20 % slower)
__global__ void __launch_bounds__(128) quad_fit_kernel(int* __restrict__ qctl, const FitRec* __restrict__ fits,
                                                       int fit_cap, const int* __restrict__ pool,
                                                       const uint16_t* __restrict__ pick_table, int table_max,
                                                       const int* __restrict__ order, FitResult* __restrict__ results,
                                                       TailRec* __restrict__ tails, int tail_cap) {
  int nfit = qctl[QC_FIT_COUNT];
  if (nfit > fit_cap) nfit = fit_cap;
  const int total = 80 * nfit;
  const int lane = threadIdx.x & 31;
  float wcache[kWelschCache];
  WelschState st;
  WelschBest best;
  PoolPts pa{pool};
  int t = 0, count = 0;
  bool have = false, drained = false;
  while (true) {
    // refill: lanes without a restart draw consecutive items (one atomic per warp and trip)
    while (true) {
      const unsigned need = __ballot_sync(0xffffffffu, !have && !drained);
      if (need == 0) break;
      int base = 0;
      if (lane == __ffs(need) - 1) base = atomicAdd(&qctl[QC_WORK_FIT], __popc(need));
      base = __shfl_sync(0xffffffffu, base, __ffs(need) - 1);
      if (!have && !drained) {
        const int u = base + __popc(need & ((1u << lane) - 1u));  // position in the sorted work list
        if (u >= total) {
          drained = true;
        } else {
          const int ei = u / 20, k = u - 20 * ei, e = order[ei], slot = e >> 2, c = e & 3;
          t = 20 * e + k;  // results stay indexed by (component, edge, restart)
          const FitRec* fr = fits + slot;
          count = fr->cl_off[c + 1] - fr->cl_off[c];
          best.err = 1.7976931348623157e308;
          best.eps = count * 1.1920928955078125e-07;
          best.sub_eps = false;
          best.line[0] = best.line[1] = best.line[2] = best.line[3] = 0.f;
          // <= 10 points: every restart starts from all points and follows the same trajectory; restart 0 stands for all
          if (count > 10 || k == 0) {
            int picked[10];
            const int np = load_picks(pick_table, table_max, count, k, picked);
            pa.p = pool + fr->pool_off + fr->cl_off[c];
            welsch_init(pa, picked, np, st);
            have = true;
          } else {
            FitResult o;
            o.err = best.err;
            o.line[0] = o.line[1] = o.line[2] = o.line[3] = 0.f;
            o.sub_eps = 0;
            o.pad = 0;
            results[t] = o;
          }
        }
      }
    }
    const unsigned act = __ballot_sync(0xffffffffu, have);
    if (act == 0) break;
    if (__any_sync(0xffffffffu, drained) && __popc(act) <= kTailLanes) {
      // The work list has run dry and only a few lanes of this warp still own a restart -- the tail of the kernel, where
      // one lane walks a whole cluster while 31 idle.  Park the restarts (between two passes their state is a handful of
      // numbers); quad_fitwarp_kernel resumes each of them with the points dealt over a whole warp.
      if (have) {
        const int idx = atomicAdd(&qctl[QC_TAIL_COUNT], 1);
        if (idx < tail_cap) {
          TailRec r;
#pragma unroll
          for (int q = 0; q < 4; ++q) r.line[q] = st.line[q], r.prev[q] = st.prev[q], r.best_line[q] = best.line[q];
          r.best_err = best.err;
          r.i = st.i;
          r.nvis = st.nvis;
          r.t = t;
          r.count = count;
          r.sub_eps = best.sub_eps ? 1 : 0;
          r.pool_index = (int)(pa.p - pool);
          tails[idx] = r;
          have = false;
        }
      }
      if (__all_sync(0xffffffffu, !have)) break;
    }
    if (have && !welsch_step(pa, count, st, wcache, best)) {
      FitResult o;
      o.err = best.err;
      o.line[0] = best.line[0], o.line[1] = best.line[1], o.line[2] = best.line[2], o.line[3] = best.line[3];
      o.sub_eps = best.sub_eps ? 1 : 0;
      o.pad = 0;
      results[t] = o;
      have = false;
    }
  }
}

__global__ void __launch_bounds__(128) quad_fitmerge_kernel(int* __restrict__ qctl, int fit_cap,
                                                            const FitResult* __restrict__ results, float* __restrict__ lines,
                                                            int* __restrict__ exact_list) {
  int nfit = qctl[QC_FIT_COUNT];
  if (nfit > fit_cap) nfit = fit_cap;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // edge id = slot * 4 + c
  if (e >= 4 * nfit) return;
  const FitResult* r = results + (size_t)e * 20;
  double best = 1.7976931348623157e308;
  int bk = 0, any_sub = 0;
  for (int k = 0; k < 20; ++k) {
    const double err = r[k].err;
    any_sub |= r[k].sub_eps;
    if (err < best) best = err, bk = k;  // strict: first occurrence wins, as in `if (err < min_err)`
  }
  if (any_sub) {
    exact_list[atomicAdd(&qctl[QC_EXACT_COUNT], 1)] = e;
    return;
  }
  *reinterpret_cast<float4*>(lines + (size_t)e * 4) = make_float4(r[bk].line[0], r[bk].line[1], r[bk].line[2], r[bk].line[3]);
}

__global__ void __launch_bounds__(32) quad_fitexact_kernel(int* __restrict__ qctl, const int* __restrict__ exact_list,
                                                           const FitRec* __restrict__ fits, const int* __restrict__ pool,
                                                           const uint16_t* __restrict__ pick_table, int table_max,
                                                           WelschIter* __restrict__ traj /* per CTA 20*30 */,
                                                           float* __restrict__ lines) {
  __shared__ int s_nvis[20];
  const int lane = threadIdx.x;
  const int total = qctl[QC_EXACT_COUNT];
  WelschIter* tw = traj + (size_t)blockIdx.x * 600;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int e = exact_list[item], slot = e >> 2, c = e & 3;
    const FitRec* fr = fits + slot;
    const int count = fr->cl_off[c + 1] - fr->cl_off[c];
    if (lane < 20) {
      int nv = 0;
      if (count > 10 || lane == 0) {
        int picked[10];
        const int np = load_picks(pick_table, table_max, count, lane, picked);
        PoolPts pa{pool + fr->pool_off + fr->cl_off[c]};
        WelschStore st{tw + lane * 30, 1};
        nv = welsch_restart_from_picks(pa, count, picked, np, st);
      }
      s_nvis[lane] = nv;
    }
    __syncwarp();
    if (lane == 0) {
      float out4[4];
      welsch_combine(tw, 30, 1, s_nvis, 1, count, 1, out4);
      *reinterpret_cast<float4*>(lines + (size_t)e * 4) = make_float4(out4[0], out4[1], out4[2], out4[3]);
    }
    __syncwarp();
  }
}

// ---- K4c: six intersections -> best 4-subset, one thread per fitted component ---------------------------------------
__global__ void __launch_bounds__(128) quad_select_kernel(const int* __restrict__ qctl, const FitRec* __restrict__ fits,
                                                          int fit_cap, float* __restrict__ lines, int* __restrict__ quad_status,
                                                          float* __restrict__ quad_corners) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  int nfit = qctl[QC_FIT_COUNT];
  if (nfit > fit_cap) nfit = fit_cap;
  if (slot >= nfit) return;
  const FitRec fr = fits[slot];
  CompView cv;
  cv.cols = fr.cols;
  cv.rows = fr.rows;
  cv.area = fr.area;
  QuadScratch sc;
  sc.lines = lines + (size_t)slot * 16;
  QuadEdges ed;
  ed.cnt = 4;
  ed.cx = fr.cx;
  ed.cy = fr.cy;
  ed.n_trace = 0;
  QuadResult r;
  quad_stage_select(cv, sc, ed, &r);
  quad_status[fr.o] = r.status;
  if (r.status == Q_OK) {
    float4* dst = reinterpret_cast<float4*>(quad_corners + (size_t)fr.o * 8);
    dst[0] = make_float4(r.c[0], r.c[1], r.c[2], r.c[3]);
    dst[1] = make_float4(r.c[4], r.c[5], r.c[6], r.c[7]);
  }
}

// One warp per frame: ordered compaction (ballot + prefix popcount) of the components that produced a quad.
__global__ void __launch_bounds__(32) quad_compact_kernel(const int* __restrict__ counters, int legal_cap,
                                                          const int* __restrict__ quad_status,
                                                          const float* __restrict__ quad_corners, int quad_cap,
                                                          float* __restrict__ quads, int* __restrict__ quad_comp,
                                                          int* __restrict__ n_quads) {
  const int fr = blockIdx.x, lane = threadIdx.x;
  const int nl = counters[fr * 4 + 1];
  int outn = 0;
  for (int c0 = 0; c0 < nl; c0 += 32) {
    int c = c0 + lane;
    bool ok = c < nl && quad_status[(size_t)fr * legal_cap + c] == Q_OK;
    unsigned bal = __ballot_sync(0xffffffffu, ok);
    int pos = outn + __popc(bal & ((1u << lane) - 1));
    if (ok && pos < quad_cap) {
      const float4* src = reinterpret_cast<const float4*>(quad_corners + ((size_t)fr * legal_cap + c) * 8);
      float4* dst = reinterpret_cast<float4*>(quads + ((size_t)fr * quad_cap + pos) * 8);
      dst[0] = src[0];
      dst[1] = src[1];
      quad_comp[(size_t)fr * quad_cap + pos] = c;
    }
    outn += __popc(bal);
  }
  if (lane == 0) n_quads[fr] = outn;  // true count; > quad_cap means the reference's isVisited[1000] would overflow
}

size_t quad_fitrec_bytes() { return sizeof(FitRec); }
size_t quad_tailrec_bytes() { return sizeof(TailRec); }
int quad_tail_cap(int sms) { return sms * 8 * 4 * kTailLanes; }  // every warp of the lane kernel parks at most kTailLanes restarts
size_t quad_fitresult_bytes() { return sizeof(FitResult); }
size_t quad_traj_bytes_per_cta() { return sizeof(WelschIter) * 600; }
int quad_edge_warps(int sms) { return sms * kEdgeCtasPerSm * kEdgeWarps; }  // persistent warps
int quad_exact_ctas(int sms) { return sms * 8; }
void quad_build_pick_table(uint16_t* host_table, int max_count) { welsch_pick_table(host_table, max_count); }

// CTAs per SM actually launched for the two persistent kernels (work-list driven: any grid size is correct).  Fewer
// than the maximum leaves registers for the kernels of the other batches in flight.
static int env_ctas(const char* name, int dflt, int maxv) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : dflt;
  return v < 1 ? 1 : (v > maxv ? maxv : v);
}

int launch_quad(int n, const FrameGeom& g, const uint8_t* bin, size_t bin_fstride, const int* labels, const int* legal,
                int legal_cap, const int* counters, int* prefix, int* qctl, uint8_t* scratch, int edge_warps, void* fits,
                int fit_cap, int fit_per_frame, int* frame_fit, int* pool, int pool_cap, const uint16_t* pick_table, int table_max, void* results,
                int* exact_list, int* fit_order, void* tails, int tail_cap, void* traj, int exact_ctas, int sms, float* lines, int* quad_status,
                float* quad_corners, int quad_cap, float* quads, int* quad_comp, int* n_quads, cudaStream_t stream,
                int* launches) {
  QuadScratchLayout L = make_layout(g);
  quad_prefix_kernel<<<1, 32, 0, stream>>>(counters, n, prefix, qctl, frame_fit);
  static const int edge_ctas = env_ctas("CTAG_EDGE_CTAS", kEdgeCtasPerSm, kEdgeCtasPerSm);
  static const int fit_ctas = env_ctas("CTAG_FIT_CTAS", 8, 8);
  int edge_grid = sms * edge_ctas;
  if (edge_grid > edge_warps / kEdgeWarps) edge_grid = edge_warps / kEdgeWarps;
  quad_edges_kernel<<<edge_grid, 32 * kEdgeWarps, 0, stream>>>(
      n, g, bin, bin_fstride, labels, legal, legal_cap, prefix, qctl, scratch, L, quad_status, static_cast<FitRec*>(fits),
      fit_cap, fit_per_frame, frame_fit, pool, pool_cap);
  quad_fitorder_kernel<<<1, 1024, 0, stream>>>(qctl, static_cast<const FitRec*>(fits), fit_cap, fit_order);
  quad_fit_kernel<<<sms * fit_ctas, 128, 0, stream>>>(qctl, static_cast<const FitRec*>(fits), fit_cap, pool, pick_table, table_max,
                                                      fit_order, static_cast<FitResult*>(results), static_cast<TailRec*>(tails), tail_cap);
  // one restart per warp: the largest clusters and the parked stragglers of the lane kernel
  quad_fitwarp_kernel<<<sms * 8, 128, 0, stream>>>(qctl, static_cast<const FitRec*>(fits), pool, pick_table, table_max, fit_order,
                                                   static_cast<const TailRec*>(tails), tail_cap, static_cast<FitResult*>(results));
  quad_fitmerge_kernel<<<(4 * fit_cap + 127) / 128, 128, 0, stream>>>(qctl, fit_cap, static_cast<const FitResult*>(results),
                                                                     lines, exact_list);
  quad_fitexact_kernel<<<exact_ctas, 32, 0, stream>>>(qctl, exact_list, static_cast<const FitRec*>(fits), pool, pick_table,
                                                      table_max, static_cast<WelschIter*>(traj), lines);
  quad_select_kernel<<<(fit_cap + 127) / 128, 128, 0, stream>>>(qctl, static_cast<const FitRec*>(fits), fit_cap, lines,
                                                                quad_status, quad_corners);
  quad_compact_kernel<<<n, 32, 0, stream>>>(counters, legal_cap, quad_status, quad_corners, quad_cap, quads, quad_comp,
                                            n_quads);
  CTAG_CUDA_CHECK(cudaGetLastError());
  if (launches) *launches += 9;
  return CTAG_OK;
}

}  // namespace ctag
