// Per-component quad extraction (reference row a5: corner_detector::edgeExtraction, corner_detector.cpp:171-405,
// with get_orientedEdgePoints :407-418, expand_line :125-169, get_permutation/quadJudgment :420-463).
//
// The code is written once for a cooperating group of lanes: on the device one warp (32 lanes) executes it for one
// component; the host build (tests/host_harness) runs it with a single lane so that the sequential logic can be
// checked against the oracle without a GPU.  Scalar control flow is executed redundantly by all lanes (warp-uniform),
// data-parallel sections stride over lanes and meet in warp reductions, and the two inherently serial chains
// (boundary trace, span expansion) run on lane 0.
//
// Redesign notes versus the reference (same results, different work):
//  * no per-component pixel lists / sorts: the bbox comes from the CCL stats, membership from (binary, block label);
//  * the ray-cast "visited" image is a bit map; the four silhouettes are found in one coalesced sweep;
//  * expand_line keeps exact integer moments, so the O(n) refit after every appended point becomes O(1) and is still
//    bit-identical to refitting from scratch (all sums are exact in double);
//  * the 20 random restarts of each DIST_WELSCH fit are independent given the generator state at their start, so the
//    4 x 20 restarts of a component run on different lanes and are merged in library order afterwards.
#pragma once
#include "fit_core.cuh"

namespace ctag {
namespace core {

struct Lanes {
  int id, n;
};

CT_HD float kRdpLine() { return 1.8f; }  // threshold_line,   corner_detector.h:90
CT_HD float kExpand() { return 1.2f; }   // threshold_expand, corner_detector.h:90
CT_HD float kRac() { return 0.3f; }      // threshold_RAC,    corner_detector.h:110

// ---- lane collectives --------------------------------------------------------------------------------------------
CT_HD void w_sync() {
#ifdef __CUDA_ARCH__
  __syncwarp();
#endif
}
CT_HD int w_min_i(int v) {
#ifdef __CUDA_ARCH__
  return __reduce_min_sync(0xffffffffu, v);
#else
  return v;
#endif
}
CT_HD int w_max_i(int v) {
#ifdef __CUDA_ARCH__
  return __reduce_max_sync(0xffffffffu, v);
#else
  return v;
#endif
}
CT_HD unsigned w_min_u(unsigned v) {
#ifdef __CUDA_ARCH__
  return __reduce_min_sync(0xffffffffu, v);
#else
  return v;
#endif
}
CT_HD unsigned w_max_u(unsigned v) {
#ifdef __CUDA_ARCH__
  return __reduce_max_sync(0xffffffffu, v);
#else
  return v;
#endif
}
CT_HD int w_sum_i(int v) {
#ifdef __CUDA_ARCH__
  return __reduce_add_sync(0xffffffffu, v);
#else
  return v;
#endif
}
// max of v over the lanes below this one (identity `none` for lane 0)
CT_HD int w_prefix_max_excl_i(int v, int lane, int none) {
#ifdef __CUDA_ARCH__
  int inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = inc > u ? inc : u;
  }
  const int ex = __shfl_up_sync(0xffffffffu, inc, 1);
  return lane == 0 ? none : ex;
#else
  (void)v;
  (void)lane;
  return none;
#endif
}
CT_HD int popc32(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
CT_HD int w_bcast_i(int v, int src) {
#ifdef __CUDA_ARCH__
  return __shfl_sync(0xffffffffu, v, src);
#else
  (void)src;
  return v;
#endif
}
CT_HD void bit_or(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
  atomicOr(p, v);
#else
  *p |= v;
#endif
}
CT_HD uint32_t f2u(float f) { return as_u32(f); }

// ---- inputs / scratch / result ------------------------------------------------------------------------------------
struct CompView {
  const uint8_t* bin;  // half-res binary image of the frame ({0,255})
  int bpitch;
  const int* labels;  // one label per 2x2 block: index of the component's root block, -1 = background
  int bw;             // blocks per row
  int cols, rows;     // half-res image size (what the reference passes as img.cols/img.rows)
  int root, area, x0, y0, x1, y1;
};

struct QuadScratch {
  uint32_t* vis;     // boundary bit map of the bbox, ((w+31)/32)*h words
  int16_t* col_top;  // [cols]
  int16_t* col_bot;  // [cols]
  int* pts_a;        // [pmax + 8] packed points
  int* pts_b;        // [pmax + 8]
  int* stack;        // [pmax + 8] trace frames: x | y << 12 | next_dir << 24
  int* cl;           // [pmax + 8] the four clusters back to back
  uint64_t* rng;     // [80] generator state at the start of each (cluster, restart)
  WelschIter* iters; // [80 * 30]
  int* nvis;         // [80]
  float* lines;      // [16]
};

enum { Q_OK = 0, Q_FEW_EDGES = 1, Q_NO_QUAD = 2 };

struct QuadResult {
  int status;
  int n_trace;  // traced boundary points
  int n_edges;  // accepted edges (cnt_boundary)
  float c[8];   // 4 corners (x,y), angular order, half-res coordinates
};

CT_HD bool in_comp(const CompView& cv, int x, int y) {
  int ax = cv.x0 + x, ay = cv.y0 + y;
  return cv.bin[(size_t)ay * cv.bpitch + ax] != 0 && cv.labels[(ay >> 1) * cv.bw + (ax >> 1)] == cv.root;
}

// squared norm of p[i] + p[(i+2)%n] - 2 p[(i+1)%n]; the reference compares its float sqrt with 1.05:
// cost > 1.05 <=> d2 >= 2, cost < 1.05 <=> d2 <= 1 (integer geometry).
CT_HD int second_diff2(const int* P, int n, int i) {
  int a = P[i], b = P[(i + 2) % n], c = P[(i + 1) % n];
  int dx = pt_x(a) + pt_x(b) - 2 * pt_x(c);
  int dy = pt_y(a) + pt_y(b) - 2 * pt_y(c);
  return dx * dx + dy * dy;
}

// cv::solve / cv::determinant for 2x2 CV_32F
CT_HD bool solve2x2(float a00, float a01, float a10, float a11, float b0, float b1, float* x0, float* x1) {
  double det = (double)a00 * a11 - (double)a01 * a10;
  if (det == 0) return false;
  double d = 1. / det;
  *x0 = (float)(((double)b0 * a11 - (double)b1 * a01) * d);
  *x1 = (float)(((double)b1 * a00 - (double)b0 * a10) * d);
  return true;
}

// bits (x-1, x, x+1) of bit-map row y as bits 0..2; rows outside the box and columns outside the row read as 0
CT_HD uint32_t row_bits3(const uint32_t* vis, int wpr, int h, int y, int x) {
  if (y < 0 || y >= h) return 0u;
  const uint32_t* row = vis + y * wpr;
  const int xm = x - 1;
  if (xm < 0) return (row[0] << 1) & 7u;
  const int wi = xm >> 5, sh = xm & 31;
  uint64_t v = row[wi];
  if (sh > 29 && wi + 1 < wpr) v |= (uint64_t)row[wi + 1] << 32;
  return (uint32_t)(v >> sh) & 7u;
}

// The 8-neighbour mask as a function of the three 3-bit row windows, v = t0 | t1 << 3 | t2 << 6 (t = bits x-1, x, x+1):
// one table look-up instead of eight bit moves in the trace's dependent chain.
struct NbrLut {
  uint8_t m[512];
  constexpr NbrLut() : m() {
    for (int v = 0; v < 512; ++v) {
      const int t0 = v & 7, t1 = (v >> 3) & 7, t2 = v >> 6;
      m[v] = (uint8_t)(((t0 >> 1) & 1) | (((t0 >> 2) & 1) << 1) | (((t1 >> 2) & 1) << 2) | (((t2 >> 2) & 1) << 3) |
                       (((t2 >> 1) & 1) << 4) | ((t2 & 1) << 5) | ((t1 & 1) << 6) | ((t0 & 1) << 7));
    }
  }
};
#if defined(__CUDACC__)
static __constant__ NbrLut kNbrLutDev = NbrLut();
#endif
static const NbrLut kNbrLutHost = NbrLut();
CT_HD uint32_t nbr_lut(uint32_t v) {
#ifdef __CUDA_ARCH__
  return kNbrLutDev.m[v];
#else
  return kNbrLutHost.m[v];
#endif
}

// 8-neighbour occupancy of (x,y) in the reference's probing order N,NE,E,SE,S,SW,W,NW (corner_detector.h:84-85)
CT_HD uint32_t nbr_mask(const uint32_t* vis, int wpr, int h, int x, int y) {
  uint32_t t0, t1, t2;
  if (wpr == 1) {  // boxes up to 32 px wide (most quads): one word per row, no straddling
    const uint32_t r0 = y > 0 ? vis[y - 1] : 0u, r1 = vis[y], r2 = y + 1 < h ? vis[y + 1] : 0u;
    if (x == 0) t0 = (r0 << 1) & 7u, t1 = (r1 << 1) & 7u, t2 = (r2 << 1) & 7u;
    else t0 = (r0 >> (x - 1)) & 7u, t1 = (r1 >> (x - 1)) & 7u, t2 = (r2 >> (x - 1)) & 7u;
  } else {
    t0 = row_bits3(vis, wpr, h, y - 1, x), t1 = row_bits3(vis, wpr, h, y, x), t2 = row_bits3(vis, wpr, h, y + 1, x);
  }
  return ((t0 >> 1) & 1u) | (((t0 >> 2) & 1u) << 1) | (((t1 >> 2) & 1u) << 2) | (((t2 >> 2) & 1u) << 3) |
         (((t2 >> 1) & 1u) << 4) | ((t2 & 1u) << 5) | ((t1 & 1u) << 6) | ((t0 & 1u) << 7);
}

struct PtArray {
  const int* p;
  CT_HD int operator()(int j) const { return p[j]; }
};

CT_HD float expand_dist(int px, int py, const float* line) {  // dist_expand of expand_line (:144,:156)
  return fabsf((float)px * line[1] - (float)py * line[0] + line[0] * line[3] - line[1] * line[2]);
}

// expand_line (:125-169), literal sequential form with exact incremental moments.  Returns how many points were
// appended on the left (indices init-1, init-2, ... wrapping) and on the right (end+1, ... wrapping).
CT_HD void expand_span_seq(const int* P, int n, int init, int end, int* nl_out, int* nr_out) {
  IntMoments mom;
  im_reset(mom);
  for (int i = init; i <= end; ++i) im_add(mom, pt_x(P[i]), pt_y(P[i]));
  float line[4];
  im_fit(mom, line);
  bool fl = false, fr = false;
  int left = init - 1, right = end + 1, nl = 0, nr = 0;
  while ((!fl || !fr) && left != right) {
    if (!fl) {
      if (left == -1) left = n - 1;
      int px = pt_x(P[left]), py = pt_y(P[left]);
      if (expand_dist(px, py, line) > kExpand()) {
        fl = true;
        continue;
      }
      im_add(mom, px, py);
      ++nl;
      --left;
      im_fit(mom, line);
      if (mom.n == n) break;
    }
    if (!fr) {
      if (right == n) right = 0;
      int px = pt_x(P[right]), py = pt_y(P[right]);
      if (expand_dist(px, py, line) > kExpand()) {
        fr = true;
        continue;
      }
      im_add(mom, px, py);
      ++nr;
      ++right;
      im_fit(mom, line);
      if (mom.n == n) break;
    }
  }
  *nl_out = nl;
  *nr_out = nr;
}

// The same function, speculatively parallel: the order in which points would be appended is known as long as no
// distance test fails (left/right alternately, or one side only), so 32 lanes each assume "everything before me was
// accepted", build that prefix fit from exact integer prefix moments and test their own candidate.  The first failing
// lane ends the round; everything before it is accepted exactly as the sequential loop would have.  A span of k points
// costs about k/32 + 3 fits of latency instead of k.  (The host build emulates the 32 lanes with a loop.)
CT_HD void expand_span(const int* P, int n, int init, int end, Lanes ln, int* nl_out, int* nr_out) {
  const int SW = 32;
  IntMoments mom;
  im_reset(mom);
#ifdef __CUDA_ARCH__
  {
    int sx = 0, sy = 0;
    long long sxx = 0, syy = 0, sxy = 0;
    for (int i = init + ln.id; i <= end; i += ln.n) {
      int x = pt_x(P[i]), y = pt_y(P[i]);
      sx += x, sy += y, sxx += x * x, syy += y * y, sxy += x * y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sxx += __shfl_xor_sync(0xffffffffu, sxx, o);
      syy += __shfl_xor_sync(0xffffffffu, syy, o);
      sxy += __shfl_xor_sync(0xffffffffu, sxy, o);
    }
    mom.sx = sx, mom.sy = sy, mom.sxx = sxx, mom.syy = syy, mom.sxy = sxy, mom.n = end - init + 1;
  }
#else
  (void)ln;
  for (int i = init; i <= end; ++i) im_add(mom, pt_x(P[i]), pt_y(P[i]));
#endif
  bool fl = false, fr = false;
  int nl = 0, nr = 0;
  int frozen_l = 0, frozen_r = 0;  // a finished side keeps the (normalised) index of the candidate that failed
  const int left0 = init - 1, right0 = end + 1;
  // raw cursor values as the reference's `left != right` sees them: an active cursor that has just stepped past the
  // end is still -1 / n (it is normalised at its next use), a finished one stays on its failed candidate
#define CT_RAW_L(q) (fl ? frozen_l : ((left0 - (q)) >= -1 ? (left0 - (q)) : (left0 - (q)) + n))
#define CT_RAW_R(q) (fr ? frozen_r : ((right0 + (q)) <= n ? (right0 + (q)) : (right0 + (q)) - n))
  while (true) {
    if ((fl && fr) || CT_RAW_L(nl) == CT_RAW_R(nr)) break;
    const bool both = !fl && !fr;
    const int m0 = mom.n;
    int limit = n - m0 < SW ? n - m0 : SW;  // candidates that exist before the list is fully covered
    int acc, fail_at = -1, fail_idx = 0;
    IntMoments add;  // moments of the accepted candidates of this round
    im_reset(add);
#ifdef __CUDA_ARCH__
    {
      const int i = ln.id;
      const bool is_left = both ? !(i & 1) : !fl;
      const int q = both ? (i >> 1) : i;
      int raw = is_left ? left0 - (nl + q) : right0 + (nr + q);
      int idx = is_left ? (raw >= 0 ? raw : raw + n) : (raw < n ? raw : raw - n);
      // loop-top equality in front of the iteration this candidate starts
      bool topstop = false;
      if (i > 0 && (!both || !(i & 1)))
        topstop = CT_RAW_L(nl + (both ? q : (is_left ? q : 0))) == CT_RAW_R(nr + (both ? q : (is_left ? 0 : q)));
      unsigned tmask = __ballot_sync(0xffffffffu, topstop);
      if (tmask) {
        int t = __ffs(tmask) - 1;
        limit = t < limit ? t : limit;
      }
      const bool exist = i < limit;
      int px = 0, py = 0;
      if (exist) px = pt_x(P[idx]), py = pt_y(P[idx]);
      // inclusive prefix sums over lanes (all fit 32 bits: 32 points, coordinates < 4096)
      int sx = px, sy = py, sxx = px * px, syy = py * py, sxy = px * py;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int a = __shfl_up_sync(0xffffffffu, sx, o), b = __shfl_up_sync(0xffffffffu, sy, o);
        int c = __shfl_up_sync(0xffffffffu, sxx, o), d = __shfl_up_sync(0xffffffffu, syy, o);
        int e = __shfl_up_sync(0xffffffffu, sxy, o);
        if (i >= o) sx += a, sy += b, sxx += c, syy += d, sxy += e;
      }
      bool fail = false;
      if (exist) {
        IntMoments pm = mom;  // exclusive prefix: everything before this candidate
        pm.sx += sx - px, pm.sy += sy - py, pm.sxx += sxx - px * px, pm.syy += syy - py * py, pm.sxy += sxy - px * py;
        pm.n += i;
        float line[4];
        im_fit(pm, line);
        fail = expand_dist(px, py, line) > kExpand();
      }
      unsigned fmask = __ballot_sync(0xffffffffu, fail);
      if (fmask) {
        fail_at = __ffs(fmask) - 1;
        fail_idx = __shfl_sync(0xffffffffu, idx, fail_at);
      }
      acc = fail_at >= 0 ? fail_at : limit;
      if (acc > 0) {
        add.sx = __shfl_sync(0xffffffffu, sx, acc - 1);
        add.sy = __shfl_sync(0xffffffffu, sy, acc - 1);
        add.sxx = __shfl_sync(0xffffffffu, sxx, acc - 1);
        add.syy = __shfl_sync(0xffffffffu, syy, acc - 1);
        add.sxy = __shfl_sync(0xffffffffu, sxy, acc - 1);
      }
    }
#else
    {
      IntMoments run;
      im_reset(run);
      acc = -1;
      for (int i = 0; i < SW && acc < 0; ++i) {
        const bool is_left = both ? !(i & 1) : !fl;
        const int q = both ? (i >> 1) : i;
        if (i > 0 && (!both || !(i & 1)) &&
            CT_RAW_L(nl + (both ? q : (is_left ? q : 0))) == CT_RAW_R(nr + (both ? q : (is_left ? 0 : q))) && i < limit)
          limit = i;
        if (i >= limit) {
          acc = limit;
          break;
        }
        int raw = is_left ? left0 - (nl + q) : right0 + (nr + q);
        int idx = is_left ? (raw >= 0 ? raw : raw + n) : (raw < n ? raw : raw - n);
        int px = pt_x(P[idx]), py = pt_y(P[idx]);
        IntMoments pm = mom;
        pm.sx += run.sx, pm.sy += run.sy, pm.sxx += run.sxx, pm.syy += run.syy, pm.sxy += run.sxy, pm.n += i;
        float line[4];
        im_fit(pm, line);
        if (expand_dist(px, py, line) > kExpand()) {
          fail_at = i;
          fail_idx = idx;
          acc = i;
          break;
        }
        im_add(run, px, py);
      }
      if (acc < 0) acc = limit;
      add = run;
    }
#endif
    mom.sx += add.sx, mom.sy += add.sy, mom.sxx += add.sxx, mom.syy += add.syy, mom.sxy += add.sxy, mom.n += acc;
    if (both) nl += (acc + 1) >> 1, nr += acc >> 1;
    else if (!fl) nl += acc;
    else nr += acc;
    if (fail_at >= 0) {
      const bool left_failed = both ? !(fail_at & 1) : !fl;
      if (left_failed) fl = true, frozen_l = fail_idx;
      else fr = true, frozen_r = fail_idx;
      continue;  // the reference's `continue`: back to the loop top, the failed side is finished
    }
    if (mom.n == n) break;  // Slide.size() == edge_point.size()
  }
#undef CT_RAW_L
#undef CT_RAW_R
  *nl_out = nl;
  *nr_out = nr;
}

// Result of the edge stage: the four point clusters (in sc.cl, back to back) and the boundary centre.
struct QuadEdges {
  int cnt;        // accepted edges (cnt_boundary); a quad needs 4
  int cl_off[5];  // cluster c = sc.cl[cl_off[c] .. cl_off[c+1])
  float cx, cy;   // area_center
  int n_trace;
};

// Steps 1-4: silhouettes, oriented trace, centre/rotation, extended RDP with span expansion.  One lane group.
CT_HD void quad_stage_edges(const CompView& cv, const QuadScratch& sc, Lanes ln, QuadEdges* out) {
  const int w = cv.x1 - cv.x0 + 1, h = cv.y1 - cv.y0 + 1, wpr = (w + 31) >> 5;
  QuadEdges* res = out;
  res->cnt = 0;
  res->n_trace = 0;

  // ---- 1. four silhouettes -> boundary bit map (corner_detector.cpp:197-232) --------------------------------------
  // The ray-cast's `if (visited) break` can never fire before the first mask pixel (visited is a subset of mask), so
  // the result is the union of the top/bottom pixel of every column and the left/right pixel of every row.
  for (int i = ln.id; i < h * wpr; i += ln.n) sc.vis[i] = 0;
#ifdef __CUDA_ARCH__
  const int xa = cv.x0 & ~3;  // 4-pixel aligned start: one binary word and two block labels per lane and row
  if (ln.n == 32 && cv.x1 - xa < 128) {
    // boxes up to 128 aligned columns (nearly all quads): four columns per lane, column extents in registers
    w_sync();
    const int cx0 = xa + 4 * ln.id;
    uint32_t cmask = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (cx0 + b >= cv.x0 && cx0 + b <= cv.x1) cmask |= 1u << b;
    const bool second = (cx0 >> 1) + 1 < cv.bw;
    int top[4] = {-1, -1, -1, -1}, bot[4] = {-1, -1, -1, -1};
#pragma unroll 2
    for (int y = 0; y < h; ++y) {
      const int ay = cv.y0 + y;
      uint32_t m = 0;
      if (cmask) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(cv.bin + (size_t)ay * cv.bpitch + cx0);
        const int* lb = cv.labels + (ay >> 1) * cv.bw + (cx0 >> 1);
        const bool in0 = lb[0] == cv.root, in1 = second && lb[1] == cv.root;
        m = (((v & 0x000000FFu) && in0) ? 1u : 0u) | (((v & 0x0000FF00u) && in0) ? 2u : 0u) |
            (((v & 0x00FF0000u) && in1) ? 4u : 0u) | (((v & 0xFF000000u) && in1) ? 8u : 0u);
        m &= cmask;
      }
      int lmin = m ? cx0 + __ffs(m) - 1 : 0x7fffffff, lmax = m ? cx0 + 31 - __clz(m) : -1;
      lmin = w_min_i(lmin);
      lmax = w_max_i(lmax);
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (m & (1u << b)) {
          if (top[b] < 0) top[b] = y;
          bot[b] = y;
        }
      if (ln.id == 0 && lmax >= 0) {
        lmin -= cv.x0, lmax -= cv.x0;
        sc.vis[y * wpr + (lmin >> 5)] |= 1u << (lmin & 31);
        sc.vis[y * wpr + (lmax >> 5)] |= 1u << (lmax & 31);
      }
    }
    w_sync();
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int x = cx0 + b - cv.x0;
      if (top[b] >= 0) {
        bit_or(&sc.vis[top[b] * wpr + (x >> 5)], 1u << (x & 31));
        bit_or(&sc.vis[bot[b] * wpr + (x >> 5)], 1u << (x & 31));
      }
      if (x == 0) sc.col_top[0] = (int16_t)top[b];
    }
    w_sync();
  } else
#endif
  {
    for (int x = ln.id; x < w; x += ln.n) sc.col_top[x] = -1, sc.col_bot[x] = -1;
    w_sync();
    for (int y = 0; y < h; ++y) {
      int lmin = 0x7fffffff, lmax = -1;
      for (int x = ln.id; x < w; x += ln.n) {
        if (in_comp(cv, x, y)) {
          if (sc.col_top[x] < 0) sc.col_top[x] = (int16_t)y;
          sc.col_bot[x] = (int16_t)y;
          lmin = x < lmin ? x : lmin;
          lmax = x > lmax ? x : lmax;
        }
      }
      lmin = w_min_i(lmin);
      lmax = w_max_i(lmax);
      if (ln.id == 0 && lmax >= 0) {
        sc.vis[y * wpr + (lmin >> 5)] |= 1u << (lmin & 31);
        sc.vis[y * wpr + (lmax >> 5)] |= 1u << (lmax & 31);
      }
    }
    w_sync();
    for (int x = ln.id; x < w; x += ln.n) {
      int t = sc.col_top[x], b = sc.col_bot[x];
      if (t >= 0) {
        bit_or(&sc.vis[t * wpr + (x >> 5)], 1u << (x & 31));
        bit_or(&sc.vis[b * wpr + (x >> 5)], 1u << (x & 31));
      }
    }
    w_sync();
  }
  // pixels on the bit map: once the trace has consumed them all, unwinding its stack cannot find anything new
  int remaining = 0;
  for (int i = ln.id; i < h * wpr; i += ln.n) remaining += popc32(sc.vis[i]);
  remaining = w_sum_i(remaining);

  // ---- 2. start pixel + oriented trace (corner_detector.cpp:235-247, 407-418) -------------------------------------
  // get_orientedEdgePoints follows 8-neighbours in the order N,NE,E,SE,S,SW,W,NW, recursing on every hit; after the
  // recursive call returns, the caller keeps scanning the remaining directions from the NEW position.  Frames of that
  // recursion live on an explicit stack.
  int* P = sc.pts_a;
  int* Q = sc.pts_b;
  int n = 0;
  if (ln.id == 0) {
    int fx = 0, fy = sc.col_top[0], fj = 0, sp = 0;
    P[n++] = pt_pack(fx + cv.x0, fy + cv.y0);
    sc.vis[fy * wpr] &= ~1u;
    --remaining;
    // The walk moves one pixel at a time, so the three bit-map words around the current position (word column cw, rows
    // cy-1..cy+1) are kept in registers: a horizontal move needs no load at all, a vertical one loads a single new word,
    // and clearing a visited bit is a plain store.  Memory stays authoritative (every change is stored); the registers
    // are reloaded after a pop and when the walk crosses into another word column.
    int cw = 0, cy = fy;
    uint32_t w0 = cy > 0 ? sc.vis[(cy - 1) * wpr] : 0u, w1 = sc.vis[cy * wpr], w2 = cy + 1 < h ? sc.vis[(cy + 1) * wpr] : 0u;
    while (remaining > 0) {
      // Directions fj..7 are probed at a fixed position and the bit map only changes inside a recursive call, so
      // the first hit is the first set bit of the 8-neighbour mask at or after fj.
      uint32_t m = 0u;
      if (fj < 8) {
        const int xb = fx & 31;
        if (xb >= 1 && xb <= 30) {  // the 3x3 neighbourhood lies inside the cached words
          const int sh = xb - 1;
          m = nbr_lut(((w0 >> sh) & 7u) | (((w1 >> sh) & 7u) << 3) | (((w2 >> sh) & 7u) << 6));
        } else {
          m = nbr_mask(sc.vis, wpr, h, fx, fy);
        }
        m >>= fj;
      }
      if (m == 0) {
        if (sp == 0) break;
        int fr = sc.stack[--sp];
        fx = fr & 0xFFF, fy = (fr >> 12) & 0xFFF, fj = fr >> 24;
        cw = fx >> 5, cy = fy;
        w0 = cy > 0 ? sc.vis[(cy - 1) * wpr + cw] : 0u, w1 = sc.vis[cy * wpr + cw];
        w2 = cy + 1 < h ? sc.vis[(cy + 1) * wpr + cw] : 0u;
        continue;
      }
#ifdef __CUDA_ARCH__
      const int j = fj + __ffs(m) - 1;
#else
      const int j = fj + __builtin_ctz(m);
#endif
      const int nx = fx + (int)((0x1A9u >> (2 * j)) & 3u) - 1;
      const int ny = fy + (int)((0x1A90u >> (2 * j)) & 3u) - 1;
      const uint32_t clr = ~(1u << (nx & 31));
      if ((nx >> 5) == cw) {
        // the visited bit lies in one of the cached words (ny is cy-1, cy or cy+1): clear it there and store that word
        const int dy = ny - cy;
        if (dy < 0) w0 &= clr;
        else if (dy == 0) w1 &= clr;
        else w2 &= clr;
        sc.vis[ny * wpr + cw] = dy < 0 ? w0 : (dy == 0 ? w1 : w2);
        if (dy < 0) {
          w2 = w1, w1 = w0;
          w0 = ny > 0 ? sc.vis[(ny - 1) * wpr + cw] : 0u;
        } else if (dy > 0) {
          w0 = w1, w1 = w2;
          w2 = ny + 1 < h ? sc.vis[(ny + 1) * wpr + cw] : 0u;
        }
        cy = ny;
      } else {
        sc.vis[ny * wpr + (nx >> 5)] &= clr;
        cw = nx >> 5, cy = ny;
        w0 = cy > 0 ? sc.vis[(cy - 1) * wpr + cw] : 0u, w1 = sc.vis[cy * wpr + cw];
        w2 = cy + 1 < h ? sc.vis[(cy + 1) * wpr + cw] : 0u;
      }
      P[n++] = pt_pack(nx + cv.x0, ny + cv.y0);
      sc.stack[sp++] = nx | (ny << 12) | ((j + 1) << 24);  // caller resumes at direction j+1 from the new position
      fx = nx, fy = ny, fj = 0;
      --remaining;
    }
  }
  w_sync();
  n = w_bcast_i(n, 0);
  res->n_trace = n;

  // ---- 3. centre of the traced boundary, rotate the list to the point closest to it (:250-275) --------------------
  float cx, cy;
  {
    int sx = 0, sy = 0;
    for (int i = ln.id; i < n; i += ln.n) sx += pt_x(P[i]), sy += pt_y(P[i]);
    sx = w_sum_i(sx);
    sy = w_sum_i(sy);
    cx = (float)(1.0 * sx / n);
    cy = (float)(1.0 * sy / n);
    float best = 3.0e38f;
    int bidx = 0x7fffffff;
    for (int i = ln.id; i < n; i += ln.n) {
      float dx = (float)pt_x(P[i]) - cx, dy = (float)pt_y(P[i]) - cy;
      float d = sqrtf(dx * dx + dy * dy);
      if (d < best) best = d, bidx = i;  // strict: first minimum (stable ascending sort)
    }
    unsigned bm = w_min_u(f2u(best));  // distances are >= 0: unsigned order == float order
    int b0 = w_min_i(f2u(best) == bm ? bidx : 0x7fffffff);
    if (b0 > 0) {
      for (int i = ln.id; i < n; i += ln.n) {
        int s = i + b0;
        Q[i] = P[s >= n ? s - n : s];
      }
      w_sync();
      int* t = P;
      P = Q;
      Q = t;
    }
  }

  // ---- 4. extended Ramer-Douglas-Peucker: up to four edges (:278-349) ----------------------------------------------
  int cnt = 0, init = 0;
  bool failed = false;
  int cl_off[5] = {0, 0, 0, 0, 0};
  while (n > 0 && !failed && cnt < 4) {
    if (n > 2) {
      while (second_diff2(P, n, init) >= 2 && init < n - 3) ++init;
    } else {
      failed = true;
      break;
    }
    int end = init + n / 2;
    if (end > n - 1) end = n - 1;
    while (true) {
      if (end <= init + 1) {
        failed = true;
        break;
      }
      const int xi = pt_x(P[init]), yi = pt_y(P[init]), xe = pt_x(P[end]), ye = pt_y(P[end]);
      const float nl0 = (xi == xe) ? 100.f : (float)(1.0 * (ye - yi) / (xe - xi));  // C-13: vertical chord = slope 100
      const float nl1 = -1.f;
      const float d_line = -(nl0 * (float)xi + nl1 * (float)yi);
      const float den = sqrtf(nl0 * nl0 + 1.f);
      float best = 0.f;
      int bidx = -1;
      for (int it = init + 1 + ln.id; it < end; it += ln.n) {
        float d = fabsf(nl0 * (float)pt_x(P[it]) + nl1 * (float)pt_y(P[it]) + d_line) / den;
        if (d >= best) best = d, bidx = it;  // >=: last index among ties of the maximum (stable descending sort)
      }
      unsigned bm = w_max_u(f2u(best));
      int last = w_max_i(f2u(best) == bm ? bidx : -1);
      float dmax;
      {
        // recover the float from its bits without type punning through memory on the device
#ifdef __CUDA_ARCH__
        dmax = __uint_as_float(bm);
#else
        memcpy(&dmax, &bm, 4);
#endif
      }
      const int m = end - init - 1;
      if (dmax > kRdpLine() && m > 1) {
        end = last - (init + 1);  // C-5: the index relative to init+1 is used as an absolute index
        continue;
      }
      // ---- expand_line (:125-169): speculative lane-parallel form, exact incremental moments ---------------------
      int nl = 0, nr = 0;
      expand_span(P, n, init, end, ln, &nl, &nr);
      // the span is the cyclic interval [a, a+len) of indices; the reference sorts it descending
      const int len = (end - init + 1) + nl + nr;
      int a = init - nl;
      if (a < 0) a += n;
      const bool wrap = a + len > n;
      const int nhi = n - a;              // wrap: indices a..n-1 come first in descending order
      const int e = a + len - n - 1;      // wrap: then e..0
      const int s0 = wrap ? n - 1 : a + len - 1;
      int* cl = sc.cl + cl_off[cnt];
      for (int q = ln.id; q < len; q += ln.n) {
        int idx = !wrap ? (a + len - 1 - q) : (q < nhi ? n - 1 - q : e - (q - nhi));
        cl[q] = P[idx];
      }
      cl_off[cnt + 1] = cl_off[cnt] + len;
      // endpoint keeping judgment (:337-339): the largest index stays in the list if the boundary is straight there
      const int keep = second_diff2(P, n, s0) <= 1 ? 1 : 0;
      for (int i = ln.id; i < n; i += ln.n) {
        if (!wrap) {
          if (i < a) Q[i] = P[i];
          else if (i == s0) { if (keep) Q[a] = P[i]; }
          else if (i > s0) Q[i - len + keep] = P[i];
        } else {
          if (i > e && i < a) Q[i - (e + 1)] = P[i];
          else if (i == s0 && keep) Q[a - e - 1] = P[i];
        }
      }
      w_sync();
      {
        int* t = P;
        P = Q;
        Q = t;
      }
      n = n - len + keep;
      const int back = wrap ? 0 : a;  // smallest erased index
      ++cnt;
      init = back >= n ? 0 : back;
      break;
    }
  }
  res->cnt = cnt;
  res->cx = cx;
  res->cy = cy;
  for (int c = 0; c < 5; ++c) res->cl_off[c] = cl_off[c];
}

// ---- 5. four DIST_WELSCH fits (:358): 4 x 20 independent restarts = 80 tasks ----------------------------------------
// (a) generator state at the start of every (cluster, restart): run by 4 lanes, one per cluster
CT_HD void quad_welsch_prepare(const QuadEdges& ed, const QuadScratch& sc, int c) {
  Rng rng{0xFFFFFFFFFFFFFFFFull};
  const int count = ed.cl_off[c + 1] - ed.cl_off[c];
  for (int k = 0; k < 20; ++k) {
    sc.rng[c * 20 + k] = rng.state;
    welsch_skip_restart(rng, count);
  }
}
// (b) one restart, t = cluster * 20 + restart.  With <= 10 points every restart starts from all points and follows the
//     same trajectory, so only restart 0 is computed and the others report no iterates (the bookkeeping only ever
//     takes strictly smaller errors, so repeats never change it).
CT_HD void quad_welsch_task(const QuadEdges& ed, const QuadScratch& sc, int t) {
  const int c = t / 20, k = t - 20 * c;
  const int count = ed.cl_off[c + 1] - ed.cl_off[c];
  if (count <= 10 && k > 0) {
    sc.nvis[t] = 0;
    return;
  }
  PtArray pa{sc.cl + ed.cl_off[c]};
  sc.nvis[t] = welsch_restart(pa, count, Rng{sc.rng[t]}, sc.iters + t * 30, 1, 0.0);
}
// (c) library bookkeeping over the 20 restarts of cluster c
CT_HD void quad_welsch_combine(const QuadEdges& ed, const QuadScratch& sc, int c) {
  welsch_combine(sc.iters + c * 600, 30, 1, sc.nvis + c * 20, 1, ed.cl_off[c + 1] - ed.cl_off[c], 1, sc.lines + 4 * c);
}

// ---- 6. six intersections, angular sort, best 4-subset (:362-403, 420-463); scalar, any single lane ----------------
CT_HD void quad_stage_select(const CompView& cv, const QuadScratch& sc, const QuadEdges& ed, QuadResult* res) {
  const float cx = ed.cx, cy = ed.cy;
  res->status = Q_NO_QUAD;
  res->n_trace = ed.n_trace;
  res->n_edges = ed.cnt;
  float ix[6], iy[6], ia[6];
  int nc = 0;
  for (int j = 0; j < 3; ++j)
    for (int k = j + 1; k < 4; ++k) {
      const float* lj = sc.lines + 4 * j;
      const float* lk = sc.lines + 4 * k;
      float b0 = lj[1] * lj[2] - lj[0] * lj[3];
      float b1 = lk[1] * lk[2] - lk[0] * lk[3];
      float sx, sy;
      if (!solve2x2(lj[1], -lj[0], lk[1], -lk[0], b0, b1, &sx, &sy)) continue;
      float dx = sx - cx, dy = sy - cy;
      float dis = sqrtf(dx * dx + dy * dy);
      float ang = (float)atan2_deg(dy, dx);
      if (dis < (float)cv.cols && dis < (float)cv.rows) {
        // insertion keeps the list sorted by angle, stable for ties (C-11)
        int pos = nc;
        while (pos > 0 && ang < ia[pos - 1]) {
          ix[pos] = ix[pos - 1], iy[pos] = iy[pos - 1], ia[pos] = ia[pos - 1];
          --pos;
        }
        ix[pos] = sx, iy[pos] = sy, ia[pos] = ang;
        ++nc;
      }
    }
  float rac_min = kRac();
  int best[4] = {-1, -1, -1, -1};
  for (int a = 0; a < nc; ++a)
    for (int b = a + 1; b < nc; ++b)
      for (int c = b + 1; c < nc; ++c)
        for (int d = c + 1; d < nc; ++d) {
          const float X[4] = {ix[a], ix[b], ix[c], ix[d]}, Y[4] = {iy[a], iy[b], iy[c], iy[d]};
          auto tri = [&](int i0, int i1, int i2) {
            return X[i0] * Y[i1] + X[i1] * Y[i2] + X[i2] * Y[i0] - X[i0] * Y[i2] - X[i1] * Y[i0] - X[i2] * Y[i1];
          };
          if (fabsf(tri(0, 1, 2)) < 1 || fabsf(tri(1, 2, 3)) < 1 || fabsf(tri(2, 3, 0)) < 1 || fabsf(tri(0, 1, 3)) < 1)
            continue;
          float qa = 0;
          for (int i = 0; i < 3; ++i) qa += X[i] * Y[i + 1] - Y[i] * X[i + 1];
          qa += X[3] * Y[0] - Y[3] * X[0];
          qa /= 2;
          float rac = fabsf(fabsf(qa) - (float)cv.area) / (float)cv.area;
          if (rac < rac_min) {
            rac_min = rac;
            best[0] = a, best[1] = b, best[2] = c, best[3] = d;
          }
        }
  if (best[0] < 0) {
    res->status = Q_NO_QUAD;
    return;
  }
  for (int k = 0; k < 4; ++k) {
    float x = ix[best[k]], y = iy[best[k]];
    if (x < 0 || y < 0 || x > (float)cv.cols || y > (float)cv.rows) {
      res->status = Q_NO_QUAD;
      return;
    }
    res->c[2 * k] = x;
    res->c[2 * k + 1] = y;
  }
  res->status = Q_OK;
}

// All stages with one lane group (host harness; the device kernel spreads stage 5 over a whole CTA instead).
CT_HD void quad_extract(const CompView& cv, const QuadScratch& sc, Lanes ln, QuadResult* res) {
  QuadEdges ed;
  quad_stage_edges(cv, sc, ln, &ed);
  res->status = Q_FEW_EDGES;
  res->n_trace = ed.n_trace;
  res->n_edges = ed.cnt;
  if (ed.cnt < 4) return;  // some cluster has < 2 points (:351-361)
  for (int c = ln.id; c < 4; c += ln.n) quad_welsch_prepare(ed, sc, c);
  w_sync();
  for (int t = ln.id; t < 80; t += ln.n) quad_welsch_task(ed, sc, t);
  w_sync();
  for (int c = ln.id; c < 4; c += ln.n) quad_welsch_combine(ed, sc, c);
  w_sync();
  quad_stage_select(cv, sc, ed, res);
}

}  // namespace core
}  // namespace ctag
