// K4: quad extraction (reference row a5, corner_detector.cpp:171-405) as three kernels over device work lists:
//   quad_edges_kernel  one warp per legal component: silhouettes, oriented trace, extended RDP -> four clusters
//   quad_fit_kernel    one warp per (component, edge): the 20 Welsch restarts of cv::fitLine on 20 lanes
//   quad_select_kernel one thread per component: six intersections, best 4-subset
// plus the ordered compaction of the surviving quads.  The per-component arithmetic lives in quad_core.cuh /
// fit_core.cuh (shared with the host logic tests).
#include "common.cuh"
#include "kernels.cuh"
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
enum { QC_WORK_EDGES = 0, QC_WORK_FITS = 1, QC_FIT_COUNT = 2, QC_POOL_CURSOR = 3, QC_OVERFLOW = 4, QC_WORDS = 8 };

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
__global__ void quad_prefix_kernel(const int* __restrict__ counters, int n, int* __restrict__ prefix, int* __restrict__ qctl) {
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
constexpr int kSmemPts = 256;       // points per list on the shared-memory fast path
constexpr int kSmemVisWords = 256;  // bit-map words on the shared-memory fast path

__global__ void __launch_bounds__(32 * kEdgeWarps) quad_edges_kernel(
    int n_frames, FrameGeom g, const uint8_t* __restrict__ bin, size_t bin_fstride, const int* __restrict__ labels,
    const int* __restrict__ legal, int legal_cap, const int* __restrict__ prefix, int* __restrict__ qctl,
    uint8_t* __restrict__ scratch, QuadScratchLayout L, int* __restrict__ quad_status, FitRec* __restrict__ fits, int fit_cap,
    int* __restrict__ pool, int pool_cap) {
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
        slot = atomicAdd(&qctl[QC_FIT_COUNT], 1);
        poff = atomicAdd(&qctl[QC_POOL_CURSOR], ed.cl_off[4]);
        if (slot >= fit_cap || poff + ed.cl_off[4] > pool_cap) {
          atomicExch(&qctl[QC_OVERFLOW], 1);
          slot = -2;
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

// ---- K4b: DIST_WELSCH fits, one warp per (component, edge) -------------------------------------------------------------
// Lane 0 replays the generator to get the state at the start of each of the 20 restarts; lanes 0..19 then run one
// restart each.  If no restart sees an error below EPS the library result is the first occurrence of the minimum
// error (warp arg-min, ties to the lower restart).  Otherwise the restarts are replayed with their trajectories
// stored and the library's exact bookkeeping runs on lane 0.
constexpr int kFitWarps = 4;
constexpr int kFitSmemPts = 256;

struct SmemOrGlobalPts {
  const int* p;
  __device__ __forceinline__ int operator()(int j) const { return p[j]; }
};

__global__ void __launch_bounds__(32 * kFitWarps) quad_fit_kernel(int* __restrict__ qctl, const FitRec* __restrict__ fits,
                                                                  int fit_cap, const int* __restrict__ pool,
                                                                  WelschIter* __restrict__ traj /* per warp 20*30 */,
                                                                  float* __restrict__ lines /* [fit][16] */) {
  __shared__ int s_pts[kFitWarps][kFitSmemPts];
  __shared__ uint64_t s_rng[kFitWarps][20];
  __shared__ int s_nvis[kFitWarps][20];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warp_global = blockIdx.x * kFitWarps + warp;
  int nfit = qctl[QC_FIT_COUNT];
  if (nfit > fit_cap) nfit = fit_cap;
  const int total = 4 * nfit;
  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(&qctl[QC_WORK_FITS], 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    const int slot = item >> 2, c = item & 3;
    const FitRec* fr = fits + slot;
    const int off = fr->pool_off + fr->cl_off[c];
    const int count = fr->cl_off[c + 1] - fr->cl_off[c];
    const int* src = pool + off;
    if (count <= kFitSmemPts) {
      for (int i = lane; i < count; i += 32) s_pts[warp][i] = src[i];
      src = s_pts[warp];
    }
    if (lane == 0) {
      Rng rng{0xFFFFFFFFFFFFFFFFull};
      const int nk = count <= 10 ? 1 : 20;  // <= 10 points: every restart picks all points, only restart 0 is needed
      for (int k = 0; k < nk; ++k) {
        s_rng[warp][k] = rng.state;
        welsch_skip_restart(rng, count);
      }
    }
    __syncwarp();
    SmemOrGlobalPts pa{src};
    const bool active = lane < 20 && (count > 10 || lane == 0);
    WelschBest best;
    best.err = 1.7976931348623157e308;
    best.eps = count * 1.1920928955078125e-07;
    best.sub_eps = false;
    best.line[0] = best.line[1] = best.line[2] = best.line[3] = 0.f;
    if (active) welsch_restart_visit(pa, count, Rng{s_rng[warp][lane]}, best);
    float out4[4];
    if (__ballot_sync(0xffffffffu, active && best.sub_eps) == 0u) {
      // first occurrence of the global minimum: min error, ties to the lowest restart index
      double e = active ? best.err : 1.7976931348623157e308;
      int k = lane;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        double e2 = __shfl_xor_sync(0xffffffffu, e, s);
        int k2 = __shfl_xor_sync(0xffffffffu, k, s);
        if (e2 < e || (e2 == e && k2 < k)) e = e2, k = k2;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) out4[q] = __shfl_sync(0xffffffffu, best.line[q], k);
    } else {
      WelschIter* tw = traj + (size_t)warp_global * 600;
      if (lane < 20) s_nvis[warp][lane] = 0;
      __syncwarp();
      if (active) {
        WelschStore st{tw + lane * 30, 1};
        s_nvis[warp][lane] = welsch_restart_visit(pa, count, Rng{s_rng[warp][lane]}, st);
      }
      __syncwarp();
      if (lane == 0) welsch_combine(tw, 30, 1, s_nvis[warp], 1, count, 1, out4);
#pragma unroll
      for (int q = 0; q < 4; ++q) out4[q] = __shfl_sync(0xffffffffu, out4[q], 0);
    }
    if (lane == 0) *reinterpret_cast<float4*>(lines + (size_t)slot * 16 + 4 * c) = make_float4(out4[0], out4[1], out4[2], out4[3]);
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
size_t quad_traj_bytes_per_warp() { return sizeof(WelschIter) * 600; }
int quad_edge_warps(int sms) { return sms * 4 * kEdgeWarps; }   // persistent: 4 CTAs x 4 warps per SM
int quad_fit_warps(int sms) { return sms * 8 * kFitWarps; }     // persistent: 8 CTAs x 4 warps per SM

int launch_quad(int n, const FrameGeom& g, const uint8_t* bin, size_t bin_fstride, const int* labels, const int* legal,
                int legal_cap, const int* counters, int* prefix, int* qctl, uint8_t* scratch, int edge_warps, void* fits,
                int fit_cap, int* pool, int pool_cap, void* traj, int fit_warps, float* lines, int* quad_status,
                float* quad_corners, int quad_cap, float* quads, int* quad_comp, int* n_quads, cudaStream_t stream,
                int* launches) {
  QuadScratchLayout L = make_layout(g);
  quad_prefix_kernel<<<1, 32, 0, stream>>>(counters, n, prefix, qctl);
  quad_edges_kernel<<<edge_warps / kEdgeWarps, 32 * kEdgeWarps, 0, stream>>>(
      n, g, bin, bin_fstride, labels, legal, legal_cap, prefix, qctl, scratch, L, quad_status, static_cast<FitRec*>(fits),
      fit_cap, pool, pool_cap);
  quad_fit_kernel<<<fit_warps / kFitWarps, 32 * kFitWarps, 0, stream>>>(qctl, static_cast<const FitRec*>(fits), fit_cap, pool,
                                                                        static_cast<WelschIter*>(traj), lines);
  quad_select_kernel<<<(fit_cap + 127) / 128, 128, 0, stream>>>(qctl, static_cast<const FitRec*>(fits), fit_cap, lines,
                                                                quad_status, quad_corners);
  quad_compact_kernel<<<n, 32, 0, stream>>>(counters, legal_cap, quad_status, quad_corners, quad_cap, quads, quad_comp,
                                            n_quads);
  CTAG_CUDA_CHECK(cudaGetLastError());
  if (launches) *launches += 5;
  return CTAG_OK;
}

}  // namespace ctag
