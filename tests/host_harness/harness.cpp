// Host-side logic harness (TEST INFRASTRUCTURE): compiles the __host__ __device__ cores of the CUDA kernels for the
// CPU so that their sequential logic can be checked against the oracle without a GPU.  Never linked into the product.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../cylindertag_b200/csrc/fit_core.cuh"
#include "../../cylindertag_b200/csrc/quad_core.cuh"

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
}
