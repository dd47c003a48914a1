// Feature grouping into markers, cross-ratio IDs and dictionary decoding
// (reference rows a9-a10: corner_detector.cpp:976-1052 markerOrganization, :1054-1209 featureExtraction,
//  :1211-1250 markerDecoder, :1269-1324 match_dictionary).  Host/device shared, see quad_core.cuh for the conventions.
#pragma once
#include "../../include/ctag.h"
#include "feature_core.cuh"

namespace ctag {
namespace core {

struct FeatureRec {
  float c[16];      // 8 corners (x,y), full-res
  float cx, cy;     // feature_center
  float angle;      // feature_angle
  int qi, qj;       // source quads (debug)
};

// cv::fastAtan2 (degrees in [0,360)), scalar path of OpenCV's mathfuncs: 7th-order odd polynomial.
// Only feeds the 45/135 degree ordering decision (:1028-1034).
CT_HD float fast_atan2_deg(float y, float x) {
  const float s = (float)(180 / 3.1415926535897932384626433832795);
  const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s,
              p7 = -0.04432655554792128f * s;
  float ax = fabsf(x), ay = fabsf(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)2.2204460492503131e-16);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)2.2204460492503131e-16);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// link test between features i < j (:982-985)
CT_HD bool feature_link(const FeatureRec& fi, const FeatureRec& fj) {
  float vcx = fi.cx - fj.cx, vcy = fi.cy - fj.cy;
  float vlx = fi.c[0] - fi.c[10], vly = fi.c[1] - fi.c[11];  // corners[0] - corners[5]
  float ca = (vcx * vlx + vcy * vly) / sqrtf((vcx * vcx + vcy * vcy) * (vlx * vlx + vly * vly));
  float da = fabsf(fi.angle - fj.angle);
  bool c1 = da < 10.0f || fabsf(180 - da) < 5.0f;
  bool c2 = (double)dist_pts(fi.cx, fi.cy, fj.cx, fj.cy) < 0.3 * (double)dist_pts(fi.c[0], fi.c[1], fi.c[10], fi.c[11]);
  bool c3 = fabsf(ca) < 0.5f;
  return c1 && c2 && c3;
}

CT_HD int uf_find_c(int* father, int x) {  // union_find with full path compression (:1256-1258)
  int r = x;
  while (father[r] != r) r = father[r];
  while (father[x] != r) {
    int nx = father[x];
    father[x] = r;
    x = nx;
  }
  return r;
}

struct IdState {  // the members ID_left / ID_right (corner_detector.h:133), reset per frame (SURVEY C-2)
  int left, right, stale;
};

CT_HD void line3(float px, float py, float qx, float qy, float rx, float ry, float* l) {
  l[0] = py - qy;
  l[1] = qx - px;
  l[2] = -l[0] * rx - l[1] * ry;
}

CT_HD int cr_band(float cr, bool is_long, int cur, bool* hit) {
  const float idc[4] = {1.47f, 1.54f, 1.61f, 1.68f};
  const float covl[4] = {0.1f, 0.035f, 0.035f, 0.035f};
  const float covr[4] = {0.035f, 0.035f, 0.035f, 0.1f};
  *hit = false;
  for (int j = 0; j < 4; ++j) {
    if (idc[j] >= cr && idc[j] - cr < covl[j]) cur = is_long ? 7 - j : j, *hit = true;
    if (idc[j] < cr && cr - idc[j] < covr[j]) cur = is_long ? 7 - j : j, *hit = true;
  }
  return cur;
}

// featureExtraction for one feature (:1056-1207).  c16 is modified in place (the halves are swapped for direction 0
// when c0.x > c4.x; src and dst alias in the reference).  Returns feature_ID.
CT_HD int feature_extract_one(float* c, int direction, IdState& ids, float* cr_left, float* cr_right, int* id_left,
                              int* id_right) {
  if (!direction && c[0] > c[8]) {
    for (int k = 0; k < 8; ++k) {
      float t = c[k];
      c[k] = c[8 + k];
      c[8 + k] = t;
    }
  }
#define CX(k) c[2 * (k)]
#define CY(k) c[2 * (k) + 1]
#define DP(a, b) dist_pts(CX(a), CY(a), CX(b), CY(b))
  float l1[4] = {DP(0, 3), DP(3, 6), DP(6, 5), DP(0, 5)};
  float l2[4] = {DP(1, 2), DP(2, 7), DP(7, 4), DP(1, 4)};
  float crl = (l1[0] + l1[1]) * (l1[2] + l1[1]) / ((l1[1] * l1[3]));
  float crr = (l2[0] + l2[1]) * (l2[2] + l2[1]) / ((l2[1] * l2[3]));
  float line1[3], line2[3], cross1[3], cross2[3], lleft[3];
  line3(CX(5), CY(5), CX(4), CY(4), CX(5), CY(5), line1);
  line3(CX(0), CY(0), CX(1), CY(1), CX(0), CY(0), line2);
  line3(CX(0), CY(0), CX(4), CY(4), CX(0), CY(0), cross1);
  line3(CX(5), CY(5), CX(1), CY(1), CX(5), CY(5), cross2);
  line3(CX(5), CY(5), CX(0), CY(0), CX(5), CY(5), lleft);
  float vpx = 0, vpy = 0, mpx = 0, mpy = 0, mlx_ = 0, mly_ = 0;  // cv::Point2f default-constructs to (0,0)
  solve2x2(line1[0], line1[1], line2[0], line2[1], -line1[2], -line2[2], &vpx, &vpy);
  solve2x2(cross1[0], cross1[1], cross2[0], cross2[1], -cross1[2], -cross2[2], &mpx, &mpy);
  float ml[3];
  ml[0] = mpy - vpy;
  ml[1] = vpx - mpx;
  ml[2] = -ml[0] * mpx - ml[1] * mpy;
  solve2x2(ml[0], ml[1], lleft[0], lleft[1], -ml[2], -lleft[2], &mlx_, &mly_);
  // middle_right is computed by the reference but the right side reuses middle_left (SURVEY C-7)
#define DM(k) dist_pts(mlx_, mly_, CX(k), CY(k))
  bool hit;
  bool is_long = DM(3) * DM(5) < DM(0) * DM(6);
  ids.left = cr_band(crl, is_long, ids.left, &hit);
  ids.stale += hit ? 0 : 1;
  is_long = DM(2) * DM(4) < DM(1) * DM(7);
  ids.right = cr_band(crr, is_long, ids.right, &hit);
  ids.stale += hit ? 0 : 1;
#undef DM
#undef DP
#undef CX
#undef CY
  *cr_left = crl;
  *cr_right = crr;
  if ((double)fabsf(l1[1] - l2[1]) > 0.05 * (double)(l1[1] + l2[1])) {
    *id_left = -1;
    *id_right = -1;
    return -2;
  }
  *id_left = ids.left;
  *id_right = ids.right;
  return ids.left * 8 + ids.right;
}

CT_HD int c_mod(int a, int b) { return a % b; }  // C semantics (sign of the dividend), as in the reference

// Scratch for one frame.
struct DecodeScratch {
  int* father;      // [100]
  uint8_t* link;    // 16 bytes, 4-byte aligned: link bit mask of the feature being joined (up to 128 features)
  int* group_of;    // [100] group index of every feature
  int* order;       // [100] feature indices of the current group, sorted
  int* cover;       // [2 * rows * cols] coverage table of match_dictionary
  ctag_marker* mk;  // working record of the current group (shared memory on the GPU: zeroed and stored by all lanes)
};

// Whole tail of detect() for one frame.  Returns the number of decoded markers written
// (at most out_cap are stored, the true count is returned).
CT_HD int organize_and_decode(const FeatureRec* feats, int nf, const int* state, int srows, int scols, int fsz, Lanes ln,
                              const DecodeScratch& sc, ctag_marker* out, int out_cap, int frame, int* n_groups, int* flagged,
                              int* stale) {
  // ---- union-find over linked feature pairs (:977-991) ----
  for (int i = ln.id; i < nf; i += ln.n) sc.father[i] = i;
  w_sync();
  // links of feature i to the features behind it as a bit mask (a feature links to a few neighbours only), then lane 0
  // joins them in ascending j like the reference's inner loop
  uint32_t* lmask = reinterpret_cast<uint32_t*>(sc.link);  // 4 words: up to 128 features
  for (int w = ln.id; w < 4; w += ln.n) lmask[w] = 0u;
  w_sync();
  for (int i = 0; i < nf - 1; ++i) {
    for (int j = i + 1 + ln.id; j < nf; j += ln.n)
      if (feature_link(feats[i], feats[j])) bit_or(&lmask[j >> 5], 1u << (j & 31));
    w_sync();
    if (ln.id == 0) {
      for (int w = (i + 1) >> 5; w < 4 && 32 * w < nf; ++w) {
        uint32_t m = lmask[w];
        lmask[w] = 0u;
        while (m) {
          const int j = 32 * w + popc32((m & (0u - m)) - 1u);  // lowest set bit
          m &= m - 1u;
          int a = uf_find_c(sc.father, i), b = uf_find_c(sc.father, j);
          if (a != b) sc.father[b] = a;
        }
      }
    }
    w_sync();
  }
  int ngroups = 0, nout = 0, flag = 0;
  IdState ids{0, 0, 0};
  // everything below is short and order dependent: lane 0 runs it, the other lanes only help in match_dictionary
  // ---- groups in first-member order (:993-1019).  The database is seeded with father[0] BEFORE the compression
  //      loop, so feature 0 is split from its group whenever its stored parent is not the final root. ----
  int db[CTAG_MAX_FRAME_FEATURES];
  if (ln.id == 0) {
    db[0] = sc.father[0];
    sc.group_of[0] = 0;
    ngroups = 1;
    for (int i = 1; i < nf; ++i) {
      int now = sc.father[i];
      while (now != sc.father[now]) now = uf_find_c(sc.father, now);
      sc.father[i] = now;
    }
    for (int i = 1; i < nf; ++i) {
      int g = -1;
      for (int j = 0; j < ngroups; ++j)
        if (sc.father[i] == db[j]) {
          g = j;
          break;
        }
      if (g < 0) {
        g = ngroups;
        db[ngroups++] = sc.father[i];
      }
      sc.group_of[i] = g;
    }
  }
  w_sync();
  ngroups = w_bcast_i(ngroups, 0);
  for (int g = 0; g < ngroups; ++g) {
    // per-group working record (lane 0 fills it); unused slots of the record are zero
    ctag_marker& mk = *sc.mk;
    constexpr int kMkWords = (int)(sizeof(ctag_marker) / 4);
    {
      int* z = reinterpret_cast<int*>(&mk);
      for (int q = ln.id; q < kMkWords; q += ln.n) z[q] = 0;
    }
    w_sync();
    int m = 0, pos_now = 0, legal = 0, decode_ok = 0;
    int code[20];
    if (ln.id == 0) {
      // members in feature order
      for (int i = 0; i < nf; ++i)
        if (sc.group_of[i] == g) sc.order[m++] = i;
      // marker angle (:1023-1032)
      float marker_angle = 0;
      for (int k = 0; k < m; ++k) {
        const FeatureRec& f = feats[sc.order[k]];
        double a = fast_atan2_deg(f.c[1] - f.c[11], f.c[0] - f.c[10]);
        if (a > 180) a -= 180;
        marker_angle = (float)((double)marker_angle + a);
      }
      marker_angle /= (float)m;
      const int direc = (fabsf(marker_angle) < 45 || fabsf(marker_angle) > 135) ? 0 : 1;
      // stable insertion sort: y descending (direction 0) or x ascending (direction 1)  (:1034-1049, C-11)
      for (int a = 1; a < m; ++a) {
        int v = sc.order[a], b = a - 1;
        while (b >= 0 && (direc == 0 ? feats[v].cy > feats[sc.order[b]].cy : feats[v].cx < feats[sc.order[b]].cx)) {
          sc.order[b + 1] = sc.order[b];
          --b;
        }
        sc.order[b + 1] = v;
      }
      if (m > CTAG_MAX_FEATURES) {
        // more members than code[20] could ever hold: the ID state still advances, the group is dropped and flagged
        for (int k = 0; k < m; ++k) {
          float crl, crr, cc[16];
          int il, ir;
          for (int q = 0; q < 16; ++q) cc[q] = feats[sc.order[k]].c[q];
          feature_extract_one(cc, direc, ids, &crl, &crr, &il, &ir);
        }
        flag = 1;
        m = -1;
      } else {
        mk.marker_id = -1;
        mk.n_features = m;
        mk.inverse = 0;
        mk.frame = frame;
        for (int k = 0; k < m; ++k) {
          const FeatureRec& f = feats[sc.order[k]];
          float cc[16];
          for (int q = 0; q < 16; ++q) cc[q] = f.c[q];
          mk.center[k][0] = f.cx;
          mk.center[k][1] = f.cy;
          // edge_length with the reference's precedence (SURVEY C-6): d(c0,c1) + d(c4,c5)/2
          mk.edge_length[k] = dist_pts(f.c[0], f.c[1], f.c[2], f.c[3]) + dist_pts(f.c[8], f.c[9], f.c[10], f.c[11]) / 2;
          mk.feature_id[k] = feature_extract_one(cc, direc, ids, &mk.cr_left[k], &mk.cr_right[k], &mk.id_left[k], &mk.id_right[k]);
          for (int q = 0; q < 16; ++q) mk.corners[k][q >> 1][q & 1] = cc[q];
          mk.feature_pos[k] = -1;
        }
        // ---- markerDecoder (:1214-1231) ----
        if (m >= fsz) {
          for (int q = 0; q < 20; ++q) code[q] = -1;
          code[0] = mk.feature_id[0];
          bool bad = false;
          for (int k = 1; k < m && !bad; ++k) {
            float dfe = dist_pts(mk.center[k][0], mk.center[k][1], mk.center[k - 1][0], mk.center[k - 1][1]);
            float val = dfe / ((mk.edge_length[k] + mk.edge_length[k - 1]) * 3 / 4);
            if (!(val == val) || fabsf(val) > 1.0e6f) {
              bad = true;
              break;
            }
            int gap = (int)roundf(val);
            pos_now += gap;
            if (pos_now < 0 || pos_now >= 20) {
              bad = true;  // code[20] overflow in the reference (SURVEY C-4): marker dropped, frame flagged
              break;
            }
            code[pos_now] = mk.feature_id[k];
          }
          if (bad) {
            flag = 1;
          } else {
            for (int q = 0; q < 20; ++q) legal += code[q] > -1 ? 1 : 0;
            decode_ok = 1;
          }
        }
      }
      for (int q = 0; q < 20; ++q) sc.cover[2 * srows * scols + q] = code[q];  // hand the code to the other lanes
    }
    w_sync();
    decode_ok = w_bcast_i(decode_ok, 0);
    pos_now = w_bcast_i(pos_now, 0);
    if (decode_ok) {
      // ---- match_dictionary (:1269-1324): coverage of every (direction, row, shift) in parallel ... ----
      const int* cd = sc.cover + 2 * srows * scols;
      // (dir, i, j) of entry t = ln.id, ln.id + ln.n, ... kept by carries instead of two divisions per entry
      int dir = 0, i = ln.id / scols, j = ln.id - (ln.id / scols) * scols;
      while (i >= srows) i -= srows, ++dir;
      const int step_i = ln.n / scols, step_j = ln.n - step_i * scols;
      for (int t = ln.id; t < 2 * srows * scols; t += ln.n) {
        int cov = 0;
        const int* row = state + i * scols;
        if (dir == 0) {
          // state(i, (j + k) % cols): the column walks right and wraps
          int col = j;
          for (int k = 0; k <= pos_now; ++k) {
            if (row[col] == cd[k]) ++cov;
            if (++col == scols) col = 0;
          }
        } else {
          // inverse reading: state read at the flat offset i * cols + (j - k + cols) % cols with C's remainder, i.e. a
          // contiguous cv::Mat read whose column offset turns negative once k > j + cols (then it walks 0, -1, ...,
          // -(cols - 1), 0, ... into the previous row); reads before the buffer never match.  a = j - k + cols.
          int a = j + scols, col = j;
          for (int k = 0; k <= pos_now; ++k) {
            const int ck = cd[k];
            const int inv = (7 - ck / 8) + (7 - ck % 8) * 8;
            const int idx = i * scols + col;
            if (idx >= 0 && state[idx] == inv) ++cov;
            --a;
            if (a >= 0) col = col == 0 ? scols - 1 : col - 1;
            else col = col == -(scols - 1) ? 0 : col - 1;
          }
        }
        sc.cover[t] = cov;
        j += step_j;
        i += step_i;
        if (j >= scols) j -= scols, ++i;
        while (i >= srows) i -= srows, ++dir;
      }
      w_sync();
      // ---- ... then the max / second bookkeeping of the reference's single sequential scan
      //   if (cov > max) { max = cov; remember t } else if (cov > second) second = cov;
      // which is not a true runner-up (SURVEY C-8): `second` only sees entries that did not raise the maximum.  The
      // scan is cut into one contiguous chunk per lane; a lane replays its chunk starting from the maximum of all
      // earlier chunks, which reproduces every comparison of the sequential scan.
      int maxc, second, tbest;
      {
        const int T = 2 * srows * scols, per = (T + ln.n - 1) / ln.n;
        const int t0 = ln.id * per, t1 = t0 + per < T ? t0 + per : T;
        int lmax = -1;
        for (int t = t0; t < t1; ++t) lmax = sc.cover[t] > lmax ? sc.cover[t] : lmax;
        int run = w_prefix_max_excl_i(lmax, ln.id, -1);
        int sec = -1, first = 0x7fffffff;
        for (int t = t0; t < t1; ++t) {
          const int cov = sc.cover[t];
          if (cov > run) {
            run = cov;
            first = t;  // last record of this chunk; the global maximum's first occurrence is the last record overall
          } else if (cov > sec) {
            sec = cov;
          }
        }
        maxc = w_max_i(lmax);
        second = w_max_i(sec);
        // the record that set the global maximum: the lane whose chunk raised the running maximum to maxc
        tbest = w_min_i((first != 0x7fffffff && run == maxc && lmax == maxc) ? first : 0x7fffffff);
      }
      int store_at = -1;
      if (ln.id == 0) {
        int pi = 0, pj = 0, direc = 1;
        if (maxc > -1 && tbest != 0x7fffffff) {
          int rc = tbest % (srows * scols);
          pi = rc / scols;
          pj = rc - pi * scols;
          direc = tbest < srows * scols ? 1 : -1;
        }
        double need = 0.8 * legal < legal - 1.0 ? 0.8 * legal : legal - 1.0;
        if ((double)maxc >= need && maxc > second) {
          mk.marker_id = pi;
          mk.inverse = direc == -1 ? 1 : 0;
          int np = 0;
          for (int k = 0; k <= pos_now; ++k)
            if (code[k] != -1 && np < CTAG_MAX_FEATURES) mk.feature_pos[np++] = c_mod(pj + direc * k + scols, scols);
          if (mk.inverse) {
            for (int k = 0; k < m; ++k)
              for (int q = 0; q < 4; ++q)
                for (int d = 0; d < 2; ++d) {
                  float tswap = mk.corners[k][q][d];
                  mk.corners[k][q][d] = mk.corners[k][q + 4][d];
                  mk.corners[k][q + 4][d] = tswap;
                }
          }
          if (nout < out_cap) store_at = nout;
          ++nout;
        }
      }
      w_sync();
      store_at = w_bcast_i(store_at, 0);
      if (store_at >= 0) {
        const int* src = reinterpret_cast<const int*>(&mk);
        int* dst = reinterpret_cast<int*>(out + store_at);
        for (int q = ln.id; q < kMkWords; q += ln.n) dst[q] = src[q];
      }
      w_sync();
    }
  }
  nout = w_bcast_i(nout, 0);
  *n_groups = ngroups;
  *flagged = w_bcast_i(flag, 0);
  *stale = w_bcast_i(ids.stale, 0);
  return nout;
}

}  // namespace core
}  // namespace ctag
