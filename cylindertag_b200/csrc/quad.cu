// K4: quad extraction, one 96-thread CTA per legal component (reference row a5, corner_detector.cpp:171-405).
// The per-component algorithm lives in quad_core.cuh (shared with the host logic tests); this file holds the
// persistent-warp scheduler, the per-warp scratch carving and the ordered compaction of the surviving quads.
#include "common.cuh"
#include "kernels.cuh"
#include "quad_core.cuh"

namespace ctag {

using namespace core;

// Per-warp scratch layout (bytes), a function of the half-res geometry only.
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

// prefix[f] = number of legal components in frames < f; prefix[n] = total.  Also resets the work counter.
__global__ void quad_prefix_kernel(const int* __restrict__ counters, int n, int* __restrict__ prefix,
                                   int* __restrict__ work_counter) {
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int f = 0; f < n; ++f) {
      prefix[f] = acc;
      acc += counters[f * 4 + 1];
    }
    prefix[n] = acc;
    *work_counter = 0;
  }
}

// One CTA of 96 threads per component (persistent CTAs pull components from a global counter):
//   warp 0        : stages 1-4 (silhouettes, trace, RDP/expansion) and stage 6 (corner selection)
//   threads 0..79 : the 80 independent Welsch restarts of stage 5, one per thread
// Small components keep the boundary bit map, the point lists and the clusters in shared memory; large ones fall back
// to the per-CTA global scratch (same code, different pointers).
constexpr int kQuadThreads = 96;
constexpr int kSmemPts = 512;    // points per list that fit the shared-memory fast path
constexpr int kSmemVisWords = 512;

__global__ void __launch_bounds__(kQuadThreads) quad_kernel(int n_frames, FrameGeom g, const uint8_t* __restrict__ bin,
                                                            size_t bin_fstride, const int* __restrict__ labels,
                                                            const int* __restrict__ legal, int legal_cap,
                                                            const int* __restrict__ prefix, int* __restrict__ work_counter,
                                                            uint8_t* __restrict__ scratch, QuadScratchLayout L,
                                                            int* __restrict__ quad_status, float* __restrict__ quad_corners) {
  __shared__ int s_item;
  __shared__ QuadEdges s_ed;
  __shared__ uint32_t s_vis[kSmemVisWords];
  __shared__ int s_pts_a[kSmemPts + 8], s_pts_b[kSmemPts + 8], s_stack[kSmemPts + 8], s_cl[kSmemPts + 8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* base = scratch + (size_t)blockIdx.x * L.total;
  QuadScratch gsc;
  gsc.vis = reinterpret_cast<uint32_t*>(base + L.vis);
  gsc.col_top = reinterpret_cast<int16_t*>(base + L.col_top);
  gsc.col_bot = reinterpret_cast<int16_t*>(base + L.col_bot);
  gsc.pts_a = reinterpret_cast<int*>(base + L.pts_a);
  gsc.pts_b = reinterpret_cast<int*>(base + L.pts_b);
  gsc.stack = reinterpret_cast<int*>(base + L.stack);
  gsc.cl = reinterpret_cast<int*>(base + L.cl);
  gsc.rng = reinterpret_cast<uint64_t*>(base + L.rng);
  gsc.iters = reinterpret_cast<WelschIter*>(base + L.iters);
  gsc.nvis = reinterpret_cast<int*>(base + L.nvis);
  gsc.lines = reinterpret_cast<float*>(base + L.lines);
  const int total = prefix[n_frames];
  while (true) {
    if (tid == 0) s_item = atomicAdd(work_counter, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= total) break;
    // frame of this item: largest f with prefix[f] <= item
    int lo = 0, hi = n_frames - 1;
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
    if (((bw_ + 31) >> 5) * bh_ <= kSmemVisWords) sc.vis = s_vis;
    if (2 * (bw_ + bh_) <= kSmemPts) {  // boundary points <= silhouette pixels <= 2 * (w + h)
      sc.pts_a = s_pts_a;
      sc.pts_b = s_pts_b;
      sc.stack = s_stack;
      sc.cl = s_cl;
    }
    if (warp == 0) {
      QuadEdges ed;
      quad_stage_edges(cv, sc, Lanes{lane, 32}, &ed);
      if (lane == 0) s_ed = ed;
    }
    __syncthreads();
    const QuadEdges ed = s_ed;
    const size_t o = (size_t)fr * legal_cap + ci;
    if (ed.cnt == 4) {
      if (tid < 4) quad_welsch_prepare(ed, sc, tid);
      __syncthreads();
      if (tid < 80) quad_welsch_task(ed, sc, tid);
      __syncthreads();
      if (tid < 4) quad_welsch_combine(ed, sc, tid);
      __syncthreads();
      if (tid == 0) {
        QuadResult r;
        quad_stage_select(cv, sc, ed, &r);
        quad_status[o] = r.status;
        if (r.status == Q_OK) {
          float4* dst = reinterpret_cast<float4*>(quad_corners + o * 8);
          dst[0] = make_float4(r.c[0], r.c[1], r.c[2], r.c[3]);
          dst[1] = make_float4(r.c[4], r.c[5], r.c[6], r.c[7]);
        }
      }
    } else if (tid == 0) {
      quad_status[o] = Q_FEW_EDGES;
    }
    __syncthreads();  // shared scratch and s_item are reused by the next component
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

int launch_quad(int n, const FrameGeom& g, const uint8_t* bin, size_t bin_fstride, const int* labels, const int* legal,
                int legal_cap, const int* counters, int* prefix, int* work_counter, uint8_t* scratch, int scratch_warps,
                int* quad_status, float* quad_corners, int quad_cap, float* quads, int* quad_comp, int* n_quads,
                cudaStream_t stream, int* launches) {
  QuadScratchLayout L = make_layout(g);
  quad_prefix_kernel<<<1, 32, 0, stream>>>(counters, n, prefix, work_counter);
  int ctas = scratch_warps;  // one scratch slot per persistent CTA
  quad_kernel<<<ctas, kQuadThreads, 0, stream>>>(n, g, bin, bin_fstride, labels, legal, legal_cap, prefix, work_counter, scratch, L,
                                        quad_status, quad_corners);
  quad_compact_kernel<<<n, 32, 0, stream>>>(counters, legal_cap, quad_status, quad_corners, quad_cap, quads, quad_comp,
                                            n_quads);
  CTAG_CUDA_CHECK(cudaGetLastError());
  if (launches) *launches += 3;
  return CTAG_OK;
}

}  // namespace ctag
