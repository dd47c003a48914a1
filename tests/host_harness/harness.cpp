// Host-side logic harness (TEST INFRASTRUCTURE): compiles the __host__ __device__ cores of the CUDA kernels for the
// CPU so that their sequential logic can be checked against the oracle without a GPU.  Never linked into the product.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../cylindertag_b200/csrc/fit_core.cuh"
#include "../../cylindertag_b200/csrc/quad_core.cuh"
#include "../../cylindertag_b200/csrc/feature_core.cuh"
#include "../../cylindertag_b200/csrc/decode_core.cuh"
#include "../../cylindertag_b200/csrc/jpeg_core.cuh"

using namespace ctag::core;

extern "C" {

// points: int32 [n][2].  DIST_L2 via exact integer moments.
void hh_fit_l2(const int32_t* pts, int n, float* line) {
  IntMoments m;
  im_reset(m);
  for (int i = 0; i < n; ++i) im_add(m, pts[2 * i], pts[2 * i + 1]);
  im_fit(m, line);
}

// DIST_WELSCH, param 0, reps = aeps = 0.01.  mode: see welsch_combine.
void hh_fit_welsch(const int32_t* pts, int n, int mode, float* line, int* total_iters) {
  std::vector<int> packed(n);
  for (int i = 0; i < n; ++i) packed[i] = pt_pack(pts[2 * i], pts[2 * i + 1]);
  std::vector<WelschIter> iters(20 * 30);
  int nvis[20];
  Rng rng{0xFFFFFFFFFFFFFFFFull};
  const int* pp = packed.data();
  int tot = 0;
  for (int k = 0; k < 20; ++k) {
    Rng start = rng;
    nvis[k] = welsch_restart([pp](int j) { return pp[j]; }, n, start, &iters[k * 30], 1, 0.0);
    welsch_skip_restart(rng, n);
    tot += nvis[k];
  }
  welsch_combine(iters.data(), 30, 1, nvis, 1, n, mode, line);
  if (total_iters) *total_iters = tot;
}

// One component through quad_extract with a single lane.  labels: one int per 2x2 block.
void hh_quad_extract(const uint8_t* bin, int bpitch, const int32_t* labels, int bw, int cols, int rows, int root, int area,
                     int x0, int y0, int x1, int y1, float* corners, int32_t* info /* status, n_trace, n_edges */) {
  CompView cv{bin, bpitch, labels, bw, cols, rows, root, area, x0, y0, x1, y1};
  const int pmax = 2 * (cols + rows) + 8;
  const int w = x1 - x0 + 1, h = y1 - y0 + 1;
  std::vector<uint32_t> vis(((w + 31) / 32) * h + 1);
  std::vector<int16_t> ct(cols), cb(cols);
  std::vector<int> pa(pmax), pb(pmax), st(pmax), cl(pmax);
  std::vector<uint64_t> rng(80);
  std::vector<WelschIter> iters(80 * 30);
  std::vector<int> nvis(80);
  float lines[16];
  QuadScratch sc{vis.data(), ct.data(), cb.data(), pa.data(), pb.data(), st.data(), cl.data(),
                 rng.data(), iters.data(), nvis.data(), lines};
  QuadResult r;
  quad_extract(cv, sc, Lanes{0, 1}, &r);
  info[0] = r.status;
  info[1] = r.n_trace;
  info[2] = r.n_edges;
  for (int i = 0; i < 8; ++i) corners[i] = r.status == Q_OK ? r.c[i] : 0.f;
}

// featureRecovery over an ordered quad list (half-res coords).  feats_out: [cap] FeatureRec.  Returns the feature count.
int hh_feature_recovery(const float* quads, int nq, FeatureRec* feats_out, int cap) {
  std::vector<QuadGeom> g(nq);
  for (int i = 0; i < nq; ++i) quad_geom(quads + 8 * i, &g[i]);
  std::vector<char> vis(nq, 0);
  int nf = 0;
  for (int i = 0; i + 1 < nq; ++i) {
    if (vis[i]) continue;
    for (int j = i + 1; j < nq; ++j) {
      if (vis[j]) continue;
      float fa;
      if (pair_test(quads + 8 * i, g[i], quads + 8 * j, g[j], &fa)) {
        vis[i] = vis[j] = 1;
        if (nf < cap) {
          FeatureRec& f = feats_out[nf];
          float cen[2];
          feature_organize(quads + 8 * i, quads + 8 * j, g[i], g[j], fa, f.c, cen);
          f.cx = cen[0], f.cy = cen[1], f.angle = fa, f.qi = i, f.qj = j;
        }
        ++nf;
        break;
      }
    }
  }
  return nf;
}

void hh_corner_obtain(FeatureRec* feats, int nf) {
  for (int i = 0; i < nf; ++i) {
    float cen[2];
    corner_obtain(feats[i].c, cen);
    feats[i].cx = cen[0], feats[i].cy = cen[1];
  }
}

void hh_edge_refine(const uint8_t* gray, int pitch, int cols, int rows, FeatureRec* feats, int nf, int win) {
  for (int i = 0; i < nf; ++i) {
    float* c = feats[i].c;
    for (int base = 0; base < 8; base += 4) {
      double nxt[4][4], lst[4][4];
      for (int e = 0; e < 4; ++e) {
        int a = base + e, b = base + ((e + 1) & 3);
        double nx, ny;
        int ns;
        edge_setup(c[2 * a], c[2 * a + 1], c[2 * b], c[2 * b + 1], &nx, &ny, &ns);
        EdgeMoments mn, ml;
        em_zero(mn), em_zero(ml);
        edge_samples(gray, pitch, cols, rows, c[2 * a], c[2 * a + 1], c[2 * b], c[2 * b + 1], win, 0, 1, nx, ny, ns, mn, ml);
        edge_line(mn, nxt[e]);
        edge_line(ml, lst[e]);
      }
      float nc[4][2];
      bool upd[4];
      for (int it = 0; it < 4; ++it) upd[it] = edge_corner(nxt[it], lst[(it + 1) & 3], &nc[it][0], &nc[it][1]);
      for (int it = 0; it < 4; ++it)
        if (upd[it]) {
          int k = base + ((it + 1) & 3);
          c[2 * k] = nc[it][0], c[2 * k + 1] = nc[it][1];
        }
    }
  }
}

int hh_organize_decode(const FeatureRec* feats, int nf, const int32_t* state, int rows, int cols, int fsz, ctag_marker* out,
                       int cap, int32_t* info /* n_groups, flagged, stale */) {
  std::vector<int> father(128), group_of(128), order(128), cover(2 * rows * cols + 32);
  std::vector<uint8_t> link(128);
  ctag_marker work_mk;
  DecodeScratch sc{father.data(), link.data(), group_of.data(), order.data(), cover.data(), &work_mk};
  return organize_and_decode(feats, nf, state, rows, cols, fsz, Lanes{0, 1}, sc, out, cap, 0, &info[0], &info[1], &info[2]);
}

int hh_sizeof_feature(void) { return (int)sizeof(FeatureRec); }

// expand_line in its literal sequential form and in the speculative form used by the kernels (lanes emulated by a loop)
void hh_expand_both(const int32_t* pts, int n, int init, int end, int32_t* out4) {
  std::vector<int> packed(n);
  for (int i = 0; i < n; ++i) packed[i] = pt_pack(pts[2 * i], pts[2 * i + 1]);
  expand_span_seq(packed.data(), n, init, end, &out4[0], &out4[1]);
  expand_span(packed.data(), n, init, end, Lanes{0, 1}, &out4[2], &out4[3]);
}

// Baseline JPEG through the decoder cores of the compressed-ingest path (jpeg_core.cuh), one restart interval at a time
// like the GPU kernel.  out: h rows of `pitch` bytes, BGR interleaved.  Returns 0, or the parse status (1 bad, 2
// unsupported); *w / *h are set as soon as the header is known.
int hh_jpeg_decode(const uint8_t* data, size_t len, uint8_t* out, size_t pitch, int cap_w, int cap_h, int* w, int* h, int mode) {
  using namespace ctag::jpeg;
  static FrameHeader fh;
  size_t so = 0, sl = 0;
  int rc = parse_header(data, len, fh, &so, &sl);
  if (rc != JP_OK) return rc;
  *w = fh.width;
  *h = fh.height;
  if (!out) return 0;
  if (fh.width > cap_w || fh.height > cap_h) return 3;
  std::vector<uint32_t> off(fh.n_intervals + 1);
  rc = find_intervals_host(data + so, sl, fh.n_intervals, off.data());
  if (rc != JP_OK) return rc;
  std::vector<uint8_t> planes(fh.plane_bytes);
  if (mode == 0) {
    for (int k = 0; k < fh.n_intervals; ++k) decode_interval(fh, data + so + off[k], data + so + off[k + 1], k, planes.data());
  } else {  // the GPU's two-step form: coefficients first (fast bit reader), inverse DCT per block afterwards
    std::vector<uint8_t> padded(data, data + len);
    padded.resize(len + 32, 0);
    std::vector<int16_t> coefs((size_t)fh.n_blocks * 64, 0);
    std::vector<uint8_t> last(fh.n_blocks, 0);
    for (int k = 0; k < fh.n_intervals; ++k)
      decode_interval_coefs(fh, padded.data() + so + off[k], padded.data() + so + off[k + 1], k, coefs.data(), last.data());
    for (int c = 0; c < fh.ncomp; ++c)
      for (int b = 0; b < fh.blocks_x[c] * fh.mcus_y * fh.comp_v[c]; ++b) {
        const int bx = b % fh.blocks_x[c], by = b / fh.blocks_x[c], blk = fh.block_first[c] + b;
        idct_block_from_coefs(coefs.data() + (size_t)blk * 64, last[blk], planes.data() + fh.plane_off[c] + (size_t)by * 8 * fh.plane_pitch[c] + bx * 8,
                              fh.plane_pitch[c]);
      }
  }
  for (int y = 0; y < fh.height; ++y)
    for (int x = 0; x < fh.width; ++x) output_pixel(fh, planes.data(), x, y, out + (size_t)y * pitch + 3 * x);
  return 0;
}
}
