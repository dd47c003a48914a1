// CPU restatement of the detect path in C++ (TEST / BASELINE INFRASTRUCTURE -- never loaded by the product).
//
// Purpose: a compiled CPU baseline to time beside the GPU path (bench.py `cpu_baseline` / `--impl reference`), because
// the reference itself cannot be built here (no OpenCV/Ceres C++) and the Python oracle spends most of its
// time in interpreter glue.  Single-threaded per frame like the reference; frames are spread over host threads.
//
// Dense stages (BGR->gray main.cpp:54, 2x cubic resize CylinderTag.cpp:79, adaptiveThreshold corner_detector.cpp:28-79,
// connectedComponentsWithStats + area filter :81-107) are plain loops written here.  The sparse stages reuse the
// arithmetic cores of cylindertag_b200/csrc/*_core.cuh compiled for the host with one lane; those cores keep exact
// incremental moments in expand_line instead of refitting from scratch after every point as the reference does
// (:149-163), so this baseline is FASTER than the reference would be -- the GPU/CPU ratio it yields is conservative.
// Results are validated against the Python/cv2 oracle in tests/test_cpu_ref.py.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../cylindertag_b200/csrc/decode_core.cuh"
#include "../../cylindertag_b200/csrc/feature_core.cuh"
#include "../../cylindertag_b200/csrc/fit_core.cuh"
#include "../../cylindertag_b200/csrc/quad_core.cuh"

using namespace ctag::core;

namespace {

struct Frame {
  int w, h, hw, hh;
  std::vector<uint8_t> gray, half, bin;
  std::vector<int> labels;  // per half-res pixel, 0 = background, components numbered in BBDT order from 1
};

void bgr2gray(const uint8_t* bgr, int w, int h, std::vector<uint8_t>& gray) {
  gray.resize((size_t)w * h);
  for (size_t i = 0; i < (size_t)w * h; ++i)
    gray[i] = (uint8_t)((3735 * bgr[3 * i] + 19235 * bgr[3 * i + 1] + 9798 * bgr[3 * i + 2] + 16384) >> 15);
}

void half_resize(const std::vector<uint8_t>& g, int w, int h, std::vector<uint8_t>& out) {
  const int hw = w / 2, hh = h / 2;
  std::vector<int> hp((size_t)h * hw);
  for (int y = 0; y < h; ++y)
    for (int d = 0; d < hw; ++d) {
      auto at = [&](int x) { return (int)g[(size_t)y * w + std::min(std::max(x, 0), w - 1)]; };
      hp[(size_t)y * hw + d] = -3 * at(2 * d - 1) + 19 * at(2 * d) + 19 * at(2 * d + 1) - 3 * at(2 * d + 2);
    }
  out.resize((size_t)hw * hh);
  for (int e = 0; e < hh; ++e)
    for (int x = 0; x < hw; ++x) {
      auto at = [&](int y) { return hp[(size_t)std::min(std::max(y, 0), h - 1) * hw + x]; };
      int v = -3 * at(2 * e - 1) + 19 * at(2 * e) + 19 * at(2 * e + 1) - 3 * at(2 * e + 2);
      v = (v + 511 + ((v >> 10) & 1)) >> 10;
      out[(size_t)e * hw + x] = (uint8_t)std::min(std::max(v, 0), 255);
    }
}

void adaptive_threshold(const std::vector<uint8_t>& half, int hw, int hh, std::vector<uint8_t>& bin) {
  const int W = 5, cn = (hw + W - 1) / W, rn = (hh + W - 1) / W;
  std::vector<uint8_t> tmin((size_t)cn * rn, 255), tmax((size_t)cn * rn, 0);
  for (int y = 0; y < hh; ++y)
    for (int x = 0; x < hw; ++x) {
      uint8_t v = half[(size_t)y * hw + x];
      size_t t = (size_t)(y / W) * cn + x / W;
      tmin[t] = std::min(tmin[t], v);
      tmax[t] = std::max(tmax[t], v);
    }
  std::vector<float> thr((size_t)cn * rn, 0.f);  // border ring frozen to 0 (SURVEY C-1)
  const float k = (float)(1.0 / 255);
  for (int i = 1; i + 1 < rn; ++i)
    for (int j = 1; j + 1 < cn; ++j) {
      int mn = 255, mx = 0;
      for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
          mn = std::min(mn, (int)tmin[(size_t)(i + di) * cn + j + dj]);
          mx = std::max(mx, (int)tmax[(size_t)(i + di) * cn + j + dj]);
        }
      float s = (float)mx * k + (float)mn * k;
      thr[(size_t)i * cn + j] = std::min(0.3f, s / 2);
    }
  bin.resize((size_t)hw * hh);
  for (int y = 0; y < hh; ++y)
    for (int x = 0; x < hw; ++x)
      bin[(size_t)y * hw + x] = ((float)half[(size_t)y * hw + x] * k < thr[(size_t)(y / W) * cn + x / W]) ? 255 : 0;
}

struct Comp {
  int label, area, x0, y0, x1, y1;
};

int uf_find(std::vector<int>& p, int x) {
  while (p[x] != x) x = p[x] = p[p[x]];
  return x;
}

// 8-connected labelling; components ordered like OpenCV's BBDT: ascending smallest 2x2-block raster index
int label_components(const std::vector<uint8_t>& bin, int hw, int hh, std::vector<int>& labels, std::vector<Comp>& comps) {
  std::vector<int> par((size_t)hw * hh);
  labels.assign((size_t)hw * hh, 0);
  for (int y = 0; y < hh; ++y)
    for (int x = 0; x < hw; ++x) {
      size_t i = (size_t)y * hw + x;
      par[i] = (int)i;
      if (!bin[i]) continue;
      auto join = [&](int xx, int yy) {
        if (xx < 0 || xx >= hw || yy < 0) return;
        size_t j = (size_t)yy * hw + xx;
        if (!bin[j]) return;
        int a = uf_find(par, (int)i), b = uf_find(par, (int)j);
        if (a != b) par[std::max(a, b)] = std::min(a, b);
      };
      join(x - 1, y);
      join(x - 1, y - 1);
      join(x, y - 1);
      join(x + 1, y - 1);
    }
  const int bw = (hw + 1) / 2;
  std::vector<std::pair<int, int>> roots;  // (min block index, root)
  std::vector<int> minblk((size_t)hw * hh, 0x7fffffff);
  for (int y = 0; y < hh; ++y)
    for (int x = 0; x < hw; ++x) {
      size_t i = (size_t)y * hw + x;
      if (!bin[i]) continue;
      int r = uf_find(par, (int)i);
      minblk[r] = std::min(minblk[r], (y >> 1) * bw + (x >> 1));
    }
  for (size_t i = 0; i < (size_t)hw * hh; ++i)
    if (bin[i] && par[i] == (int)i) roots.push_back({minblk[i], (int)i});
  std::sort(roots.begin(), roots.end());
  std::vector<int> id((size_t)hw * hh, 0);
  comps.assign(roots.size() + 1, Comp{0, 0, 0x7fffffff, 0x7fffffff, -1, -1});
  for (size_t k = 0; k < roots.size(); ++k) id[roots[k].second] = (int)k + 1;
  for (int y = 0; y < hh; ++y)
    for (int x = 0; x < hw; ++x) {
      size_t i = (size_t)y * hw + x;
      if (!bin[i]) continue;
      int l = id[uf_find(par, (int)i)];
      labels[i] = l;
      Comp& c = comps[l];
      c.label = l;
      c.area++;
      c.x0 = std::min(c.x0, x), c.y0 = std::min(c.y0, y), c.x1 = std::max(c.x1, x), c.y1 = std::max(c.y1, y);
    }
  return (int)roots.size() + 1;
}

struct Result {
  int n_labels = 0, n_legal = 0, n_quads = 0, n_features = 0, n_groups = 0, n_markers = 0, status = 0, flagged = 0;
  std::vector<ctag_marker> markers;
  std::vector<float> quads;
  std::vector<int> quad_comp;
};

void detect_gray(const std::vector<uint8_t>& gray, int w, int h, const int* state, int srows, int scols, int fsz, bool subpix,
                 int dist, Result& R) {
  const int hw = w / 2, hh = h / 2;
  std::vector<uint8_t> half, bin;
  half_resize(gray, w, h, half);
  adaptive_threshold(half, hw, hh, bin);
  std::vector<int> labels;
  std::vector<Comp> comps;
  R.n_labels = label_components(bin, hw, hh, labels, comps);
  // block labels for the shared cores: any foreground pixel's component id
  const int bw = (hw + 1) / 2, bh = (hh + 1) / 2;
  std::vector<int> blk((size_t)bw * bh, -1);
  for (int y = 0; y < hh; ++y)
    for (int x = 0; x < hw; ++x)
      if (labels[(size_t)y * hw + x]) blk[(size_t)(y >> 1) * bw + (x >> 1)] = labels[(size_t)y * hw + x];
  const int area_max = (int)std::round(0.01 * hw * hh);
  const int pmax = 2 * (hw + hh) + 16;
  std::vector<uint32_t> vis((size_t)((hw + 31) / 32) * hh + 4);
  std::vector<int16_t> ct(hw), cb(hw);
  std::vector<int> pa(pmax), pb(pmax), st(pmax), cl(pmax);
  std::vector<uint64_t> rng(80);
  std::vector<WelschIter> iters(80 * 30);
  std::vector<int> nvis(80);
  float lines[16];
  QuadScratch sc{vis.data(), ct.data(), cb.data(), pa.data(), pb.data(), st.data(), cl.data(), rng.data(), iters.data(), nvis.data(), lines};
  int ci = 0;
  for (size_t k = 1; k < comps.size(); ++k) {
    const Comp& c = comps[k];
    if (c.area < 30 || c.area > area_max) continue;
    CompView cv{bin.data(), hw, blk.data(), bw, hw, hh, c.label, c.area, c.x0, c.y0, c.x1, c.y1};
    QuadResult q;
    quad_extract(cv, sc, Lanes{0, 1}, &q);
    if (q.status == Q_OK) {
      for (int i = 0; i < 8; ++i) R.quads.push_back(q.c[i]);
      R.quad_comp.push_back(ci);
    }
    ++ci;
  }
  R.n_legal = ci;
  R.n_quads = (int)R.quad_comp.size();
  if (R.n_quads == 0) {
    R.status = 1;
    return;
  }
  if (R.n_quads > 1000) {
    R.flagged = 1;
    return;
  }
  // featureRecovery
  const int nq = R.n_quads;
  std::vector<QuadGeom> g(nq);
  for (int i = 0; i < nq; ++i) quad_geom(&R.quads[8 * i], &g[i]);
  std::vector<char> used(nq, 0);
  std::vector<FeatureRec> feats;
  for (int i = 0; i + 1 < nq; ++i) {
    if (used[i]) continue;
    for (int j = i + 1; j < nq; ++j) {
      if (used[j]) continue;
      float fa;
      if (pair_test(&R.quads[8 * i], g[i], &R.quads[8 * j], g[j], &fa)) {
        used[i] = used[j] = 1;
        FeatureRec f;
        float cen[2];
        feature_organize(&R.quads[8 * i], &R.quads[8 * j], g[i], g[j], fa, f.c, cen);
        f.cx = cen[0], f.cy = cen[1], f.angle = fa, f.qi = i, f.qj = j;
        feats.push_back(f);
        break;
      }
    }
  }
  R.n_features = (int)feats.size();
  if (R.n_features < fsz) {
    R.status = 2;
    return;
  }
  if (R.n_features > 100) {
    R.flagged = 1;
    return;
  }
  for (auto& f : feats) {
    float cen[2];
    corner_obtain(f.c, cen);
    f.cx = cen[0], f.cy = cen[1];
  }
  if (subpix) {
    for (auto& f : feats)
      for (int base = 0; base < 8; base += 4) {
        double nxt[4][4], lst[4][4];
        for (int e = 0; e < 4; ++e) {
          int a = base + e, b = base + ((e + 1) & 3);
          double nx, ny;
          int ns;
          edge_setup(f.c[2 * a], f.c[2 * a + 1], f.c[2 * b], f.c[2 * b + 1], &nx, &ny, &ns);
          EdgeMoments mn, ml;
          em_zero(mn), em_zero(ml);
          edge_samples(gray.data(), w, w, h, f.c[2 * a], f.c[2 * a + 1], f.c[2 * b], f.c[2 * b + 1], dist, 0, 1, nx, ny, ns, mn, ml);
          edge_line(mn, nxt[e]);
          edge_line(ml, lst[e]);
        }
        float nc[4][2];
        bool upd[4];
        for (int it = 0; it < 4; ++it) upd[it] = edge_corner(nxt[it], lst[(it + 1) & 3], &nc[it][0], &nc[it][1]);
        for (int it = 0; it < 4; ++it)
          if (upd[it]) {
            int k = base + ((it + 1) & 3);
            f.c[2 * k] = nc[it][0], f.c[2 * k + 1] = nc[it][1];
          }
      }
  }
  std::vector<int> father(128), group_of(128), order(128), cover(2 * (size_t)srows * scols + 32);
  std::vector<uint8_t> link(128);
  ctag_marker work_mk;
  DecodeScratch ds{father.data(), link.data(), group_of.data(), order.data(), cover.data(), &work_mk};
  R.markers.resize(50);
  int flagged = 0, stale = 0;
  R.n_markers = organize_and_decode(feats.data(), R.n_features, state, srows, scols, fsz, Lanes{0, 1}, ds, R.markers.data(), 50, 0,
                                    &R.n_groups, &flagged, &stale);
  R.flagged |= flagged;
  R.markers.resize(std::min(R.n_markers, 50));
}

}  // namespace

extern "C" {

// frames: [n][h][w][channels] u8 (channels 1 or 3).  counts_out: [n][8] = n_labels, n_legal, n_quads, n_features, n_groups,
// n_markers, status, flagged.  markers_out: [n][cap].  Frames are distributed over `threads` host threads.
int cpu_ref_detect_batch(const uint8_t* frames, int n, int w, int h, int channels, const int32_t* state, int srows, int scols,
                         int fsz, int subpix, int dist, int threads, int32_t* counts_out, ctag_marker* markers_out, int cap) {
  std::atomic<int> next(0);
  auto work = [&]() {
    std::vector<uint8_t> gray;
    while (true) {
      int f = next.fetch_add(1);
      if (f >= n) break;
      const uint8_t* src = frames + (size_t)f * w * h * channels;
      if (channels == 3) bgr2gray(src, w, h, gray);
      else gray.assign(src, src + (size_t)w * h);
      Result R;
      detect_gray(gray, w, h, state, srows, scols, fsz, subpix != 0, dist, R);
      if (counts_out) {
        int32_t* c = counts_out + 8 * f;
        c[0] = R.n_labels, c[1] = R.n_legal, c[2] = R.n_quads, c[3] = R.n_features, c[4] = R.n_groups, c[5] = R.n_markers;
        c[6] = R.status, c[7] = R.flagged;
      }
      if (markers_out)
        for (int k = 0; k < (int)R.markers.size() && k < cap; ++k) markers_out[(size_t)f * cap + k] = R.markers[k];
    }
  };
  if (threads <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  return 0;
}
}
