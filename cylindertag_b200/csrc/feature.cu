// K5/K6: quad pairing into features (reference a6, corner_detector.cpp:465-598), coordinate lift (a7, :561-569)
// and sub-pixel edge refinement (a8, :600-951).
//
//   pair_kernel   : one warp per frame.  The reference's greedy O(Q^2) search is order dependent (first j > i that
//                   passes wins, both quads are consumed), so the outer loop stays sequential while the 32 lanes test
//                   32 candidates j at a time and a ballot picks the first hit.
//   refine_kernel : one CTA per feature, one warp per quad edge (8 edges).  Lanes stride over the >= 128 samples of
//                   the edge, each sample sweeps 41 offsets along the normal reading the u8 gray image through the
//                   256-entry float LUT (the float image of the reference is never materialised); the six fp64
//                   moments of both weightings are reduced with warp shuffles; corners are 2x2 solves in registers.
#include "common.cuh"
#include "decode_core.cuh"
#include "feature_core.cuh"
#include "kernels.cuh"

namespace ctag {

using namespace core;

__global__ void __launch_bounds__(32) pair_kernel(const float* __restrict__ quads, const int* __restrict__ n_quads,
                                                  int quad_cap, QuadGeom* __restrict__ geom, FeatureRec* __restrict__ feats,
                                                  int feat_cap, int feature_size, int* __restrict__ fstate /* [frame][4] */) {
  __shared__ uint8_t visited[CTAG_MAX_FRAME_QUADS];
  // the greedy search reads every quad O(Q / 32) times along ONE dependent chain: frames with up to kPairSmem quads (all
  // but pathological ones) keep corners and derived geometry in shared memory, where a read costs tens of cycles
  // instead of an L2 round trip
  constexpr int kPairSmem = 256;
  __shared__ float s_q[kPairSmem * 8];
  __shared__ QuadGeom s_g[kPairSmem];
  const int fr = blockIdx.x, lane = threadIdx.x;
  const int nq_true = n_quads[fr];
  const bool overflow_q = nq_true > quad_cap;  // the reference's isVisited[1000] would overflow (SURVEY C-4)
  const int nq = overflow_q ? 0 : nq_true;
  const float* Q = quads + (size_t)fr * quad_cap * 8;
  QuadGeom* G = geom + (size_t)fr * quad_cap;
  FeatureRec* F = feats + (size_t)fr * feat_cap;
  const bool in_smem = nq <= kPairSmem;
  if (in_smem) {
    for (int i = lane; i < 8 * nq; i += 32) s_q[i] = Q[i];
    __syncwarp();
    Q = s_q;
    G = s_g;
  }
  for (int i = lane; i < nq; i += 32) {
    QuadGeom g;
    quad_geom(Q + 8 * i, &g);
    G[i] = g;
    visited[i] = 0;
  }
  __syncwarp();
  int nf = 0;
  for (int i = 0; i + 1 < nq; ++i) {
    if (visited[i]) continue;
    const QuadGeom gi = G[i];
    int found = -1;
    float fa_found = 0.f;
    for (int j0 = i + 1; j0 < nq; j0 += 32) {
      int j = j0 + lane;
      float fa = 0.f;
      bool ok = false;
      if (j < nq && !visited[j]) ok = pair_test(Q + 8 * i, gi, Q + 8 * j, G[j], &fa);
      unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (bal) {
        int src = __ffs(bal) - 1;
        found = j0 + src;
        fa_found = __shfl_sync(0xffffffffu, fa, src);
        break;
      }
    }
    if (found >= 0) {
      if (lane == 0) {
        visited[i] = 1;
        visited[found] = 1;
        if (nf < feat_cap) {
          FeatureRec f;
          float cen[2];
          feature_organize(Q + 8 * i, Q + 8 * found, gi, G[found], fa_found, f.c, cen);
          f.cx = cen[0];
          f.cy = cen[1];
          f.angle = fa_found;
          f.qi = i;
          f.qj = found;
          F[nf] = f;
        }
      }
      ++nf;
      __syncwarp();
    }
  }
  const bool overflow_f = nf > feat_cap;  // father[100] would overflow
  int status = CTAG_FRAME_OK;
  if (nq_true == 0) status = CTAG_FRAME_NO_CORNER;
  else if (!overflow_q && nf < feature_size) status = CTAG_FRAME_NO_FEATURE;
  const int nf_eff = (status == CTAG_FRAME_OK && !overflow_q && !overflow_f) ? nf : 0;
  // cornerObtain: half-res -> full-res coordinates
  for (int k = lane; k < nf_eff; k += 32) {
    float cen[2];
    corner_obtain(F[k].c, cen);
    F[k].cx = cen[0];
    F[k].cy = cen[1];
  }
  if (lane == 0) {
    fstate[fr * 4 + 0] = status;
    fstate[fr * 4 + 1] = nf;      // true feature count
    fstate[fr * 4 + 2] = nf_eff;  // features that go on to refinement / decoding
    fstate[fr * 4 + 3] = (overflow_q || overflow_f) ? 1 : 0;
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// three CTAs per SM: 80 registers per thread (98 without the bound: two CTAs, +15 % time; 64 with four: spills, +4 %)
__global__ void __launch_bounds__(256, 3) refine_kernel(const uint8_t* __restrict__ gray, size_t gray_pitch,
                                                     size_t gray_fstride, int cols, int rows, FeatureRec* __restrict__ feats,
                                                     int feat_cap, const int* __restrict__ fstate, int win) {
  __shared__ double lines[2][2][4][4];  // [half][0 = next, 1 = last][edge][Ex,Ey,nx,ny]
  const int fr = blockIdx.y, k = blockIdx.x;
  if (k >= fstate[fr * 4 + 2]) return;
  FeatureRec* f = feats + (size_t)fr * feat_cap + k;
  const uint8_t* img = gray + (size_t)fr * gray_fstride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = warp >> 2, e = warp & 3;
  const int a = 4 * half + e, b = 4 * half + ((e + 1) & 3);
  const float ax = f->c[2 * a], ay = f->c[2 * a + 1], bx = f->c[2 * b], by = f->c[2 * b + 1];
  double nx, ny;
  int ns;
  edge_setup(ax, ay, bx, by, &nx, &ny, &ns);
  EdgeMoments mn, ml;
  em_zero(mn);
  em_zero(ml);
  edge_samples(img, (int)gray_pitch, cols, rows, ax, ay, bx, by, win, lane, 32, nx, ny, ns, mn, ml);
  mn.Mx = warp_sum_d(mn.Mx), mn.My = warp_sum_d(mn.My), mn.Mxx = warp_sum_d(mn.Mxx);
  mn.Mxy = warp_sum_d(mn.Mxy), mn.Myy = warp_sum_d(mn.Myy), mn.N = warp_sum_d(mn.N);
  ml.Mx = warp_sum_d(ml.Mx), ml.My = warp_sum_d(ml.My), ml.Mxx = warp_sum_d(ml.Mxx);
  ml.Mxy = warp_sum_d(ml.Mxy), ml.Myy = warp_sum_d(ml.Myy), ml.N = warp_sum_d(ml.N);
  if (lane == 0) {
    edge_line(mn, lines[half][0][e]);
    edge_line(ml, lines[half][1][e]);
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int h2 = threadIdx.x >> 2, it = threadIdx.x & 3;
    float cx, cy;
    if (edge_corner(lines[h2][0][it], lines[h2][1][(it + 1) & 3], &cx, &cy)) {
      const int kc = 4 * h2 + ((it + 1) & 3);
      f->c[2 * kc] = cx;
      f->c[2 * kc + 1] = cy;
    }
  }
}

// One warp per frame: markerOrganization + featureExtraction + markerDecoder (decode_core.cuh).
__global__ void __launch_bounds__(32) decode_kernel(const FeatureRec* __restrict__ feats, int feat_cap,
                                                    const int* __restrict__ fstate, const int* __restrict__ state, int srows,
                                                    int scols, int fsz, ctag_marker* __restrict__ markers, int marker_cap,
                                                    const int* __restrict__ counters, const int* __restrict__ n_quads,
                                                    int quad_cap, const int* __restrict__ batch_overflow,
                                                    const int* __restrict__ frame_overflow,
                                                    ctag_marker* __restrict__ packed, int* __restrict__ packed_count,
                                                    int* __restrict__ summary /* [frame][12] */) {
  extern __shared__ int smem_i[];
  const int fr = blockIdx.x, lane = threadIdx.x;
  DecodeScratch sc;
  sc.father = smem_i;
  sc.group_of = smem_i + 128;
  sc.order = smem_i + 256;
  sc.link = reinterpret_cast<uint8_t*>(smem_i + 384);
  sc.cover = smem_i + 384 + 32;
  __shared__ ctag_marker s_mk;
  sc.mk = &s_mk;
  const int nf = fstate[fr * 4 + 2];
  int ngroups = 0, flagged = 0, stale = 0, nm = 0;
  // the frame's features (at most CTAG_MAX_FRAME_FEATURES records of 84 bytes) move to shared memory first: everything
  // below is one dependent chain on lane 0 and reads them over and over
  __shared__ FeatureRec s_feats[CTAG_MAX_FRAME_FEATURES];
  {
    const int* src = reinterpret_cast<const int*>(feats + (size_t)fr * feat_cap);
    int* dst = reinterpret_cast<int*>(s_feats);
    const int words = (nf < CTAG_MAX_FRAME_FEATURES ? nf : CTAG_MAX_FRAME_FEATURES) * (int)(sizeof(FeatureRec) / 4);
    for (int i = lane; i < words; i += 32) dst[i] = src[i];
    __syncwarp();
  }
  // ... and so does the dictionary (behind the coverage table): the match reads every entry up to 20 times
  int* s_state = sc.cover + 2 * srows * scols + 32;
  if (nf > 0) {
    for (int i = lane; i < srows * scols; i += 32) s_state[i] = state[i];
    __syncwarp();
  }
  if (nf > 0)
    nm = organize_and_decode(nf <= CTAG_MAX_FRAME_FEATURES ? s_feats : feats + (size_t)fr * feat_cap, nf, s_state, srows, scols, fsz, Lanes{lane, 32}, sc,
                             markers + (size_t)fr * marker_cap, marker_cap, fr, &ngroups, &flagged, &stale);
  // pack this frame's markers behind those of the other frames (one D2H copy for the whole batch)
  const int nstore = nm < marker_cap ? nm : marker_cap;
  int off = 0;
  if (lane == 0 && nstore > 0) off = atomicAdd(packed_count, nstore);
  off = __shfl_sync(0xffffffffu, off, 0);
  {
    const int words = nstore * (int)(sizeof(ctag_marker) / 4);
    const int* src = reinterpret_cast<const int*>(markers + (size_t)fr * marker_cap);
    int* dst = reinterpret_cast<int*>(packed + off);
    for (int i = lane; i < words; i += 32) dst[i] = src[i];
  }
  if (lane == 0) {
    int* s = summary + fr * 12;
    const int nq = n_quads[fr];
    s[0] = fstate[fr * 4 + 0];          // status
    s[1] = counters[fr * 4 + 0] + 1;    // n_labels (+ background label 0)
    s[2] = counters[fr * 4 + 1];        // n_legal
    s[3] = nq;                          // n_quads
    s[4] = fstate[fr * 4 + 1];          // n_features
    s[5] = ngroups;
    s[6] = nm;
    s[7] = (flagged || fstate[fr * 4 + 3] || counters[fr * 4 + 2] || nq > quad_cap || nm > marker_cap || *batch_overflow || frame_overflow[fr]) ? 1 : 0;
    s[8] = stale;
    s[9] = off;                         // offset of this frame's markers in the packed list
    s[10] = nstore;
    s[11] = 0;
  }
}

size_t sizeof_quad_geom() { return sizeof(QuadGeom); }
size_t sizeof_feature_rec() { return sizeof(FeatureRec); }
size_t decode_smem_bytes(int srows, int scols) { return sizeof(int) * (384 + 32 + 3 * (size_t)srows * scols + 32); }

int launch_features(int n, const FrameGeom& g, const float* quads, const int* n_quads, int quad_cap, void* geom, void* feats,
                    int feat_cap, int feature_size, int* fstate, const uint8_t* gray, size_t gray_pitch, size_t gray_fstride,
                    int corner_subpix, int subpix_dist, cudaStream_t stream, int* launches) {
  pair_kernel<<<n, 32, 0, stream>>>(quads, n_quads, quad_cap, static_cast<QuadGeom*>(geom), static_cast<FeatureRec*>(feats),
                                    feat_cap, feature_size, fstate);
  if (launches) *launches += 1;
  if (corner_subpix) {
    refine_kernel<<<dim3(feat_cap, n), 256, 0, stream>>>(gray, gray_pitch, gray_fstride, g.w, g.h,
                                                         static_cast<FeatureRec*>(feats), feat_cap, fstate, subpix_dist);
    if (launches) *launches += 1;
  }
  CTAG_CUDA_CHECK(cudaGetLastError());
  return CTAG_OK;
}

int launch_decode(int n, const void* feats, int feat_cap, const int* fstate, const int* state, int srows, int scols, int fsz,
                  ctag_marker* markers, int marker_cap, const int* counters, const int* n_quads, int quad_cap,
                  const int* batch_overflow, const int* frame_overflow, ctag_marker* packed, int* packed_count, int* summary,
                  cudaStream_t stream,
                  int* launches) {
  size_t smem = decode_smem_bytes(srows, scols);
  if (smem > 200 * 1024) return CTAG_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    CTAG_CUDA_CHECK(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CTAG_CUDA_CHECK(cudaMemsetAsync(packed_count, 0, sizeof(int), stream));
  decode_kernel<<<n, 32, smem, stream>>>(static_cast<const FeatureRec*>(feats), feat_cap, fstate, state, srows, scols, fsz,
                                         markers, marker_cap, counters, n_quads, quad_cap, batch_overflow, frame_overflow, packed, packed_count, summary);
  CTAG_CUDA_CHECK(cudaGetLastError());
  if (launches) *launches += 1;
  return CTAG_OK;
}

}  // namespace ctag
