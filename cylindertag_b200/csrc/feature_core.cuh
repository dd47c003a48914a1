// Quad pairing into features, feature organisation, coordinate lift and edge refinement
// (reference rows a6-a8: corner_detector.cpp:465-559 featureRecovery, :571-598 featureOrganization,
//  :561-569 cornerObtain, :600-951 edgeRefine).  Host/device shared arithmetic, see quad_core.cuh for the conventions.
#pragma once
#include "quad_core.cuh"

namespace ctag {
namespace core {

// per-quad precomputation of featureRecovery (:473-481)
struct QuadGeom {
  float cx, cy;
  float d[4];
  float ang1, ang2;
  // Angles that depend on the quad alone, computed once per quad (in parallel over the quads) instead of inside the
  // sequential pairing loop: ea = the four edge directions pair_side may pick (:495,:499,:508,:512: c0-c3, c1-c2, c0-c1,
  // c2-c3), ca = direction from each corner to the centre (featureOrganization, :577-578).
  float ea[4];
  float ca[4];
};

CT_HD float dist_pts(float ax, float ay, float bx, float by) {  // distance_2points (:1252-1254)
  return sqrtf((ax - bx) * (ax - bx) + (ay - by) * (ay - by));
}

CT_HD void quad_geom(const float* q, QuadGeom* g) {
  g->cx = (q[0] + q[2] + q[4] + q[6]) / 4;
  g->cy = (q[1] + q[3] + q[5] + q[7]) / 4;
  for (int j = 0; j < 4; ++j) {
    int k = (j + 1) & 3;
    g->d[j] = sqrtf((q[2 * j] - q[2 * k]) * (q[2 * j] - q[2 * k]) + (q[2 * j + 1] - q[2 * k + 1]) * (q[2 * j + 1] - q[2 * k + 1]));
  }
  // plain arithmetic means of two angles, no wrap handling (SURVEY C-14)
  const double a01 = atan2_deg(q[1] - q[3], q[0] - q[2]), a12 = atan2_deg(q[3] - q[5], q[2] - q[4]),
               a03 = atan2_deg(q[1] - q[7], q[0] - q[6]);
  g->ang1 = (float)((a01 + atan2_deg(q[7] - q[5], q[6] - q[4])) / 2);
  g->ang2 = (float)((a12 + a03) / 2);
  g->ea[0] = (float)a03;
  g->ea[1] = (float)a12;
  g->ea[2] = (float)a01;
  g->ea[3] = (float)atan2_deg(q[5] - q[7], q[4] - q[6]);
  for (int i = 0; i < 4; ++i) g->ca[i] = (float)atan2_deg(g->cy - q[2 * i + 1], g->cx - q[2 * i]);
}

// |d| < thr  or  ||d| - 180| < thr  or  ||d| - 360| < thr   (:490 and friends)
CT_HD bool near_mod(float diff, float thr) {
  float a = fabsf(diff);
  return a < thr || fabsf(a - 180) < thr || fabsf(a - 360) < thr;
}

// role of one quad in a candidate pair (:490-515 / :516-541): returns tag; the second test overrides the first
CT_HD bool pair_side(const float*, const QuadGeom& g, float fa, float* lng, float* sht, float* ea) {
  bool tag = false;
  if (near_mod(fa - g.ang1, 5.0f)) {
    tag = true;
    *lng = (g.d[0] + g.d[2]) / 2;
    *sht = g.d[1] < g.d[3] ? g.d[1] : g.d[3];  // std::min(a,b): b < a ? b : a -- same value for floats
    *ea = g.d[1] < g.d[3] ? g.ea[0] : g.ea[1];
  }
  if (near_mod(fa - g.ang2, 5.0f)) {
    tag = true;
    *sht = g.d[0] < g.d[2] ? g.d[0] : g.d[2];
    *lng = (g.d[1] + g.d[3]) / 2;
    *ea = g.d[0] > g.d[2] ? g.ea[2] : g.ea[3];
  }
  return tag;
}

// acceptance test of a quad pair (:486-548); *fa_out = feature_angle
CT_HD bool pair_test(const float* qi, const QuadGeom& gi, const float* qj, const QuadGeom& gj, float* fa_out) {
  float fa = (float)atan2_deg(gi.cy - gj.cy, gi.cx - gj.cx);
  *fa_out = fa;
  float l1 = 0, s1 = 0, e1 = 0, l2 = 0, s2 = 0, e2 = 0;
  bool t1 = pair_side(qi, gi, fa, &l1, &s1, &e1);
  bool t2 = pair_side(qj, gj, fa, &l2, &s2, &e2);
  if (!(t1 && t2)) return false;
  float flen = dist_pts(gi.cx, gi.cy, gj.cx, gj.cy);
  float lsum = l1 + l2, ssum = s1 + s2, half = lsum / 2;
  float smin = s2 < s1 ? s2 : s1;
  return (l1 > s1 || l2 > s2) && near_mod(e1 - e2, 50.0f) && ((double)fabsf(s1 - s2) < (double)smin * 0.33) &&
         (lsum > ssum) && (lsum < 15 * ssum) && ((double)(flen - half) < 0.3 * (double)(flen + half));
}

// featureOrganization (:571-598): rotate both quads so that corners 2,3 / 6,7 face each other; out = 8 corners + centre
CT_HD void feature_organize(const float* q1, const float* q2, const QuadGeom& g1, const QuadGeom& g2, float fa, float* out16,
                            float* center2) {
  const float* a1 = g1.ca;
  const float* a2 = g2.ca;
  float amax = 0, amin = 360;
  int p1 = -1, p2 = -1;
  for (int i = 0; i < 4; ++i) {
    float da = fabsf(a1[(i + 2) & 3] - fa), db = fabsf(a1[(i + 3) & 3] - fa);
    float s1 = ((360 - da) < da ? (360 - da) : da) + ((360 - db) < db ? (360 - db) : db);
    if (s1 < amin) amin = s1, p1 = i;
    float dc = fabsf(a2[(i + 2) & 3] - fa), dd = fabsf(a2[(i + 3) & 3] - fa);
    float s2 = ((360 - dc) < dc ? (360 - dc) : dc) + ((360 - dd) < dd ? (360 - dd) : dd);
    if (s2 > amax) amax = s2, p2 = i;
  }
  // p1/p2 stay -1 only for NaN input; the reference would index with -1 % 4 -- clamp instead (unreachable in practice)
  if (p1 < 0) p1 = 0;
  if (p2 < 0) p2 = 0;
  for (int i = 0; i < 4; ++i) {
    out16[2 * i] = q1[2 * ((i + p1) & 3)];
    out16[2 * i + 1] = q1[2 * ((i + p1) & 3) + 1];
    out16[8 + 2 * i] = q2[2 * ((i + p2) & 3)];
    out16[8 + 2 * i + 1] = q2[2 * ((i + p2) & 3) + 1];
  }
  center2[0] = (out16[0] + out16[2] + out16[8] + out16[10]) / 4;
  center2[1] = (out16[1] + out16[3] + out16[9] + out16[11]) / 4;
}

// cornerObtain (:561-569): half-res -> full-res coordinates, centre recomputed from corners 0,1,4,5
CT_HD void corner_obtain(float* c16, float* center2) {
  for (int k = 0; k < 16; ++k) c16[k] = (c16[k] - 0.5f) * 2 + 0.5f;
  center2[0] = (c16[0] + c16[2] + c16[8] + c16[10]) / 4;
  center2[1] = (c16[1] + c16[3] + c16[9] + c16[11]) / 4;
}

// ---- edgeRefine (:600-951) --------------------------------------------------------------------------------------
// img_float(y,x) = float(gray(y,x)) * float(1.0/255): the float image is never materialised.
CT_HD float lut255f(int v) {
#ifdef __CUDA_ARCH__
  return __fmul_rn((float)v, (float)(1.0 / 255));
#else
  return (float)v * (float)(1.0 / 255);
#endif
}

struct EdgeMoments {  // one weighting of one edge
  double Mx, My, Mxx, Mxy, Myy, N;
};
CT_HD void em_zero(EdgeMoments& m) { m.Mx = m.My = m.Mxx = m.Mxy = m.Myy = m.N = 0; }

// Accumulates one located sample (bestx, besty) into both weightings (1-alpha: "next", alpha: "last").
CT_HD void em_add(EdgeMoments& nextm, EdgeMoments& lastm, double bestx, double besty, double alpha) {
  double wn = 1 - alpha, wl = alpha;
  nextm.Mx += bestx * wn;
  nextm.My += besty * wn;
  nextm.Mxx += bestx * bestx * wn;
  nextm.Mxy += bestx * besty * wn;
  nextm.Myy += besty * besty * wn;
  nextm.N += wn;
  lastm.Mx += bestx * wl;
  lastm.My += besty * wl;
  lastm.Mxx += bestx * bestx * wl;
  lastm.Mxy += bestx * besty * wl;
  lastm.Myy += besty * besty * wl;
  lastm.N += wl;
}

// Samples s = first, first+step, ... of the edge a->b; accumulates both weightings.
// The reference runs the identical sampling twice, once per weighting (:605-679 vs :681-755).
// Along the normal it compares the pixels at offsets n+1 and n-1 for n = -win..win step 1/4: both belong to the one
// ladder m = -(win+1)..(win+1) step 1/4, so every pixel is located and read once (n+1 = m, n-1 = m-2 = 8 steps back)
// and kept in a 9-entry ring; the arithmetic per pixel and the order of the sums are the reference's.
// float(v) * float(1/255) for a byte v without the conversion unit: 2^23 + v is exact in float, so is the subtraction.
// (Measured on the refine kernel at equal register budget: -4 %.  Replacing the double -> int and float -> double
// conversions by a round-toward-zero add / integer re-biasing as well was slower: +3 %.)
#ifdef __CUDA_ARCH__
__device__ __forceinline__ float lut255_byte(uint32_t v) {
  return __fmul_rn(__fsub_rn(__uint_as_float(0x4B000000u | v), 8388608.0f), (float)(1.0 / 255));
}
#else
CT_HD float lut255_byte(uint32_t v) { return lut255f((int)v); }
#endif

// One sample: the ladder of NM pixels along the normal through (x0, y0).  INSIDE = the caller has checked that both ends
// of the ladder (hence, x and y being monotone in m, all of it) lie inside the image: no per-pixel bounds tests then.
template <int WIN, bool INSIDE>
CT_HD void edge_ladder(const uint8_t* gray, int pitch, int cols, int rows, double x0, double y0, double nx, double ny, double& Mn,
                       double& Mcount) {
  constexpr int NM = 8 * WIN + 9;
  float ring[9];
  bool okr[9];
  double mb = -(double)(WIN + 1);  // m of the first pixel of the block: multiples of 0.25 are exact
#pragma unroll
  for (int blk = 0; blk < NM; blk += 9, mb += 2.25) {  // ring slot = i mod 9 is a compile-time constant inside the body
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int i = blk + j;
      if (i < NM) {
        const double m = mb + 0.25 * j;
        const int x = (int)(x0 + m * nx);
        const int y = (int)(y0 + m * ny);
        const bool ok = INSIDE || !(x < 0 || x >= cols || y < 0 || y >= rows);
        // a frame is smaller than 2^32 bytes: a 32-bit offset keeps the address arithmetic short
        const float gv = ok ? lut255_byte(gray[(uint32_t)(y * pitch + x)]) : 0.f;
        if (i >= 8) {
          const double n = m - 1;
          const float g1 = gv, g2 = ring[(j + 1) % 9];
          if (ok && (INSIDE || okr[(j + 1) % 9]) && !(g1 < g2)) {
            double weight = (g2 - g1) * (g2 - g1);
            Mn += weight * n;
            Mcount += weight;
          }
        }
        ring[j] = gv;
        okr[j] = ok;
      }
    }
  }
}

template <int WIN>
CT_HD void edge_samples_w(const uint8_t* gray, int pitch, int cols, int rows, float ax, float ay, float bx, float by,
                          int first, int step, double nx, double ny, int nsamples, EdgeMoments& nextm, EdgeMoments& lastm) {
  for (int s = first; s < nsamples; s += step) {
    double alpha = (15.0 + s) / (nsamples + 30);
    double x0 = alpha * ax + (1 - alpha) * bx;
    double y0 = alpha * ay + (1 - alpha) * by;
    double Mn = 0, Mcount = 0;
    // the two ends of the ladder, computed exactly as the ladder computes them
    const double me = (double)(WIN + 1);
    const int xa = (int)(x0 + (-me) * nx), xb = (int)(x0 + me * nx), ya = (int)(y0 + (-me) * ny), yb = (int)(y0 + me * ny);
    const bool inside = xa >= 0 && xa < cols && xb >= 0 && xb < cols && ya >= 0 && ya < rows && yb >= 0 && yb < rows;
    if (inside) edge_ladder<WIN, true>(gray, pitch, cols, rows, x0, y0, nx, ny, Mn, Mcount);
    else edge_ladder<WIN, false>(gray, pitch, cols, rows, x0, y0, nx, ny, Mn, Mcount);
    if (Mcount == 0) continue;
    double n0 = Mn / Mcount;
    em_add(nextm, lastm, x0 + n0 * nx, y0 + n0 * ny, alpha);
  }
}

CT_HD void edge_samples(const uint8_t* gray, int pitch, int cols, int rows, float ax, float ay, float bx, float by, int win,
                        int first, int step, double nx, double ny, int nsamples, EdgeMoments& nextm, EdgeMoments& lastm) {
  if (win == 5) {  // main.cpp's detect(..., 5, true, 5)
    edge_samples_w<5>(gray, pitch, cols, rows, ax, ay, bx, by, first, step, nx, ny, nsamples, nextm, lastm);
    return;
  }
  const double range = win;
  for (int s = first; s < nsamples; s += step) {
    double alpha = (15.0 + s) / (nsamples + 30);
    double x0 = alpha * ax + (1 - alpha) * bx;
    double y0 = alpha * ay + (1 - alpha) * by;
    double Mn = 0, Mcount = 0;
    for (double n = -range; n <= range; n += 0.25) {
      int x1 = (int)(x0 + (n + 1) * nx);
      int y1 = (int)(y0 + (n + 1) * ny);
      if (x1 < 0 || x1 >= cols || y1 < 0 || y1 >= rows) continue;
      int x2 = (int)(x0 + (n - 1) * nx);
      int y2 = (int)(y0 + (n - 1) * ny);
      if (x2 < 0 || x2 >= cols || y2 < 0 || y2 >= rows) continue;
      float g1 = lut255f(gray[(size_t)y1 * pitch + x1]);
      float g2 = lut255f(gray[(size_t)y2 * pitch + x2]);
      if (g1 < g2) continue;
      double weight = (g2 - g1) * (g2 - g1);
      Mn += weight * n;
      Mcount += weight;
    }
    if (Mcount == 0) continue;
    double n0 = Mn / Mcount;
    em_add(nextm, lastm, x0 + n0 * nx, y0 + n0 * ny, alpha);
  }
}

// moments -> (Ex, Ey, nx, ny)  (:667-678)
CT_HD void edge_line(const EdgeMoments& m, double* line4) {
  double Ex = m.Mx / m.N, Ey = m.My / m.N;
  double Cxx = m.Mxx / m.N - Ex * Ex;
  double Cxy = m.Mxy / m.N - Ex * Ey;
  double Cyy = m.Myy / m.N - Ey * Ey;
  double theta = .5 * atan2_f((float)(-2 * Cxy), (float)(Cyy - Cxx));
  line4[0] = Ex;
  line4[1] = Ey;
  line4[2] = libm_cosf((float)theta);
  line4[3] = libm_sinf((float)theta);
}

// normal / sample count of the edge a->b (:609-615)
CT_HD void edge_setup(float ax, float ay, float bx, float by, double* nx, double* ny, int* nsamples) {
  double x = by - ay;       // float subtraction, then widened
  double y = -bx + ax;
  double mag = sqrt(x * x + y * y);
  *nx = x / mag;
  *ny = y / mag;
  double ns = mag / 8;
  *nsamples = (int)(128.0 > ns ? 128.0 : ns);
}

// corner (it+1) = intersection of next[it] and last[it+1]  (:757-776); returns false if the corner is kept
CT_HD bool edge_corner(const double* nxt, const double* lst, float* cx, float* cy) {
  double A00 = nxt[3], A01 = -lst[3];
  double A10 = -nxt[2], A11 = lst[2];
  double B0 = -nxt[0] + lst[0];
  double B1 = -nxt[1] + lst[1];
  double det = A00 * A11 - A10 * A01;
  if (!(fabs(det) > 0.001)) return false;
  double W00 = A11 / det, W01 = -A01 / det;
  double L0 = W00 * B0 + W01 * B1;
  *cx = (float)(nxt[0] + L0 * A00);
  *cy = (float)(nxt[1] + L0 * A10);
  return true;
}

}  // namespace core
}  // namespace ctag
