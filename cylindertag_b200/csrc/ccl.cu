// K2/K3: connected-component labelling + stats + ordered compaction  (reference row a4, corner_detector.cpp:81-107)
//
// The reference calls connectedComponentsWithStats(8-connectivity, CCL_BBDT) and then keeps the components with
// 30 <= area <= round(0.01*w*h) in OpenCV label order.  BBDT numbers components in ascending order of the smallest
// 2x2-block raster index they touch, so a union-find over 2x2 blocks whose root is always the minimum
// block index yields the reference order directly: component order == ascending root index.
//
//   ccl_tiles  : list of the 32x32-block tiles (64x64 px) that hold foreground, from the flags the front kernel wrote
//                together with the binary image (TileHint); a frame has a handful of them.  All later kernels are
//                persistent grids over this list, so an empty image region costs one byte read.
//   ccl_local  : one CTA per listed tile.  Union-find in shared memory, partial stats per local root, one label
//                word per 2x2 block written to HBM (not one per pixel); a flag byte per (block row, tile column)
//                = "this 32-block label segment exists".  Labels of unlisted tiles are never written nor read.
//   ccl_merge  : unions across tile borders (global atomicMin union-find over the local roots only).
//   ccl_final  : per listed tile: flattens every block label to its global root, folds the partial stats of merged
//                local roots into the global root, lists the global roots of every row segment in ascending order
//                (prefix over the eight lanes of a row) and stores their number in the segment's flag byte.
//   ccl_list   : one CTA per frame: scan of the per-segment root counts in raster order -> ordered component list,
//                area filter, second scan -> ordered list of legal components {root, area, bbox}.
//
// A pixel (x,y) belongs to component `root` iff binary(x,y) != 0 and label[(y>>1)*bw + (x>>1)] == root, because all
// foreground pixels of one 2x2 block are mutually 8-connected.
#include "common.cuh"
#include "kernels.cuh"

namespace ctag {

constexpr int kTag = 1 << 30;  // label entry of a non-root block: kTag | index of its tile-local root

// pattern bits: a=1 (x,y)  b=2 (x+1,y)  c=4 (x,y+1)  d=8 (x+1,y+1)
__device__ __forceinline__ int block_pattern(const uint8_t* __restrict__ bin, const FrameGeom& g, int bx, int by) {
  int x = 2 * bx, y = 2 * by;
  int p = 0;
  const uint8_t* r0 = bin + (size_t)y * g.bpitch + x;
  // bpitch is a multiple of 16 and columns >= hw inside the pitch are zero (front kernel), so the 2-byte read is safe
  uint16_t v0 = *reinterpret_cast<const uint16_t*>(r0);
  p |= (v0 & 0xFF) ? 1 : 0;
  p |= (v0 >> 8) ? 2 : 0;
  if (y + 1 < g.hh) {
    uint16_t v1 = *reinterpret_cast<const uint16_t*>(r0 + g.bpitch);
    p |= (v1 & 0xFF) ? 4 : 0;
    p |= (v1 >> 8) ? 8 : 0;
  }
  return p;
}

// Find with path halving: every visited node is re-pointed at its grandparent.  Parents always have smaller indices than
// their children (union by minimum index), so the update is an atomicMin like the unions themselves: it can only move a
// node closer to its root, whatever else runs next to it.
__device__ __forceinline__ int uf_find(int* L, int x) {
  volatile int* Lv = L;
  int p;
  while ((p = Lv[x]) != x) {
    const int gp = Lv[p];
    if (gp != p) atomicMin(&L[x], gp);
    x = gp;
  }
  return x;
}

// union by minimum index (root of a set is always its smallest member)
__device__ __forceinline__ void uf_unite(int* L, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a < b) {
      int t = a;
      a = b;
      b = t;
    }
    int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

// 256 threads per 32x32-block tile; thread t owns the four horizontally adjacent blocks (4*(t&7) .. +3, t>>3) so that the
// binary image is read with 8-byte loads.  A tile that turns out to hold no foreground (possible only without the front
// kernel's flags) is left after that read and writes nothing.
__device__ __forceinline__ void ccl_local_tile(const uint8_t* __restrict__ bin, size_t bin_fstride, const FrameGeom& g,
                                               int* __restrict__ labels, int* __restrict__ st_area, int* __restrict__ st_x0,
                                               int* __restrict__ st_y0, int* __restrict__ st_x1, int* __restrict__ st_y1,
                                               uint8_t* __restrict__ seg_flags, int seg_pitch, size_t seg_fstride,
                                               int tile_x, int tile_y, int fr) {
  __shared__ int L[1024];
  __shared__ uint8_t pat[1024];
  __shared__ int sA[1024], sX0[1024], sY0[1024], sX1[1024], sY1[1024];
  const int t = threadIdx.x, tq = t & 7, ty = t >> 3;
  const int bx0 = tile_x * 32 + 4 * tq, by = tile_y * 32 + ty;
  const size_t base = (size_t)fr * g.nblocks;
  int p4[4] = {0, 0, 0, 0};
  uint2 a = make_uint2(0u, 0u), b = make_uint2(0u, 0u);
  if (by < g.bh && bx0 < g.bw) {
    // 8 pixels of two rows; bpitch is a multiple of 16 and columns >= hw inside the pitch are zero (front kernel)
    const uint8_t* r0 = bin + (size_t)fr * bin_fstride + (size_t)(2 * by) * g.bpitch + 2 * bx0;
    a = *reinterpret_cast<const uint2*>(r0);
    if (2 * by + 1 < g.hh) b = *reinterpret_cast<const uint2*>(r0 + g.bpitch);
  }
  // Empty tile: nothing is written at all.  seg_flags[by][tile column] (zeroed per batch) says which 32-block row
  // segments belong to a tile that holds foreground; only those have valid labels, and every reader of the label array
  // either works from the tile list / checks the flag (ccl_merge, ccl_final, ccl_list) or uses a label only where the
  // pixel is set (the quad stage).
  if (!__syncthreads_or((a.x | a.y | b.x | b.y) != 0u)) return;
  {
    const uint32_t aw[2] = {a.x, a.y}, bw_[2] = {b.x, b.y};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t va = aw[k >> 1] >> (16 * (k & 1)), vb = bw_[k >> 1] >> (16 * (k & 1));
      p4[k] = ((va & 0xFF) ? 1 : 0) | ((va & 0xFF00) ? 2 : 0) | ((vb & 0xFF) ? 4 : 0) | ((vb & 0xFF00) ? 8 : 0);
      if (bx0 + k >= g.bw) p4[k] = 0;
    }
  }
  if (tq == 0 && by < g.bh) seg_flags[(size_t)fr * seg_fstride + (size_t)by * seg_pitch + tile_x] = 1;
  const int l0 = ty * 32 + 4 * tq;  // tile-local index of the first owned block
  // Horizontal runs first: block i is glued to block i-1 when the right column of i-1 and the left column of i both
  // hold a pixel (any such pair is 8-adjacent).  Every block starts out pointing at the first block of its run, so
  // the union-find below only links runs of adjacent rows and its chains stay shorter than the tile height instead
  // of growing with the blob area.
  int left3 = __shfl_up_sync(0xffffffffu, p4[3], 1);
  if (tq == 0) left3 = 0;
  uint32_t glue = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int prev = k ? p4[k - 1] : left3;
    if ((p4[k] & 5) && (prev & 10)) glue |= 1u << k;
  }
  uint32_t H = glue << (4 * tq);  // the row's 32 glue bits: OR over the eight lanes of the row
  H |= __shfl_xor_sync(0xffffffffu, H, 1);
  H |= __shfl_xor_sync(0xffffffffu, H, 2);
  H |= __shfl_xor_sync(0xffffffffu, H, 4);
  int head[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = 4 * tq + k;
    const uint32_t z = ~H & ((2u << i) - 1u);  // bit 0 of H is never set, so z != 0
    head[k] = ty * 32 + 31 - __clz(z);
    pat[l0 + k] = (uint8_t)p4[k];
    L[l0 + k] = p4[k] ? head[k] : -1;
  }
  __syncthreads();
  // Links to the row above.  Many blocks of a run touch the same upper run; a link is skipped when a block further
  // left (or the block's own N link) already implies it, so that a solid blob costs one union per row.
  {
    int right0 = __shfl_down_sync(0xffffffffu, p4[0], 1);
    if (tq == 7) right0 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int p = p4[k], l = l0 + k, tx = 4 * tq + k;
      if (p && ty > 0) {
        const int pm1 = k ? p4[k - 1] : left3, pp1 = k < 3 ? p4[k + 1] : right0;
        const int q = pat[l - 32], qm = tx > 0 ? pat[l - 33] : 0, qp = tx < 31 ? pat[l - 31] : 0;
        const bool gl = (H >> tx) & 1u, gr = tx < 31 && ((H >> (tx + 1)) & 1u);
        const bool gu = (q & 5) && (qm & 10), gur = (qp & 5) && (q & 10);
        const bool vn = (p & 3) && (q & 12), vnl = (pm1 & 3) && (qm & 12), vnr = (pp1 & 3) && (qp & 12);
        const bool vnw = (p & 1) && (qm & 8), vne = (p & 2) && (qp & 4);
        if (vn && !(gl && gu && vnl)) uf_unite(L, l, l - 32);
        if (vnw && !(vn && gu) && !(gl && vnl)) uf_unite(L, l, l - 33);
        if (vne && !(vn && gur) && !(gr && vnr)) uf_unite(L, l, l - 31);
      }
    }
  }
  __syncthreads();
  // run heads look their root up and publish it; everybody else reads it from the head
  int r4[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) r4[k] = (p4[k] && head[k] == l0 + k) ? uf_find(L, l0 + k) : -1;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (p4[k] && head[k] == l0 + k) L[l0 + k] = r4[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) r4[k] = p4[k] ? L[head[k]] : -1;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int l = l0 + k;
    if (p4[k]) {
      L[l] = r4[k];
      if (r4[k] == l) {  // stats live at the tile-local roots only
        sA[l] = 0;
        sX0[l] = 0x7fffffff;
        sY0[l] = 0x7fffffff;
        sX1[l] = -1;
        sY1[l] = -1;
      }
    }
  }
  __syncthreads();
  // partial stats per local root: merge the thread's own blocks that share a root, then smem atomics
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = p4[k];
    if (!p) continue;
    const int r = r4[k];
    bool first = true;
    for (int q = 0; q < k; ++q) first &= !(p4[q] && r4[q] == r);
    if (!first) continue;
    int a = 0, xmin = 0x7fffffff, ymin = 0x7fffffff, xmax = -1, ymax = -1;
    for (int q = k; q < 4; ++q) {
      const int pq = p4[q];
      if (!pq || r4[q] != r) continue;
      const int x = 2 * (bx0 + q), y = 2 * by;
      a += __popc(pq);
      xmin = min(xmin, (pq & 5) ? x : x + 1);
      xmax = max(xmax, (pq & 10) ? x + 1 : x);
      ymin = min(ymin, (pq & 3) ? y : y + 1);
      ymax = max(ymax, (pq & 12) ? y + 1 : y);
    }
    atomicAdd(&sA[r], a);
    atomicMin(&sX0[r], xmin);
    atomicMin(&sY0[r], ymin);
    atomicMax(&sX1[r], xmax);
    atomicMax(&sY1[r], ymax);
  }
  __syncthreads();
  if (by < g.bh) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int bx = bx0 + k;
      if (bx >= g.bw) continue;
      const int gi = by * g.bw + bx, l = l0 + k, r = r4[k];
      int e = -1;
      if (p4[k]) {
        if (r == l) {
          e = gi;
          st_area[base + gi] = sA[l];
          st_x0[base + gi] = sX0[l];
          st_y0[base + gi] = sY0[l];
          st_x1[base + gi] = sX1[l];
          st_y1[base + gi] = sY1[l];
        } else {
          e = kTag | ((tile_y * 32 + (r >> 5)) * g.bw + tile_x * 32 + (r & 31));
        }
      }
      labels[base + gi] = e;
    }
  }
}

// tile_list[0 .. *tile_count) = (frame << 16 | tile row << 8 | tile column) of the tiles to look at: those
// the front kernel flagged (TileHint), or all of them when there are no flags.  Order does not matter.
__global__ void __launch_bounds__(256) ccl_tiles_kernel(const uint8_t* __restrict__ tile_any, int ntiles, int tiles_x, int tiles_y,
                                                        int* __restrict__ tile_list, int* __restrict__ tile_count) {
  const int tile = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31;
  const bool live = tile < ntiles && (!tile_any || tile_any[tile]);
  const unsigned m = __ballot_sync(0xffffffffu, live);
  if (!m) return;
  int at = 0;
  if (lane == 0) at = atomicAdd(tile_count, __popc(m));
  at = __shfl_sync(0xffffffffu, at, 0);
  if (live) {
    const int tpf = tiles_x * tiles_y, fr = tile / tpf, rem = tile - fr * tpf, ty = rem / tiles_x;
    tile_list[at + __popc(m & ((1u << lane) - 1u))] = (fr << 16) | (ty << 8) | (rem - ty * tiles_x);  // tiles_x, tiles_y <= 64
  }
}

// Persistent grid: CTA c works on list entries c, c + G, c + 2G, ... (neighbouring tiles go to different CTAs).
__global__ void __launch_bounds__(256, 8) ccl_local_kernel(const uint8_t* __restrict__ bin, size_t bin_fstride, FrameGeom g,
                                                           int* __restrict__ labels, int* __restrict__ st_area,
                                                           int* __restrict__ st_x0, int* __restrict__ st_y0,
                                                           int* __restrict__ st_x1, int* __restrict__ st_y1,
                                                           uint8_t* __restrict__ seg_flags, int seg_pitch, size_t seg_fstride,
                                                           const int* __restrict__ tile_list, const int* __restrict__ tile_count) {
  const int count = *tile_count;
  for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
    const int tile = tile_list[idx];
    ccl_local_tile(bin, bin_fstride, g, labels, st_area, st_x0, st_y0, st_x1, st_y1, seg_flags, seg_pitch, seg_fstride,
                   tile & 255, (tile >> 8) & 255, tile >> 16);
    __syncthreads();  // the tile's shared arrays are reused by the next one
  }
}

__device__ __forceinline__ int gfind(volatile int* lab, int x) {
  int e = lab[x];
  if (e & kTag) x = e & ~kTag;  // hop from a leaf to its tile-local root (e >= 0 here: callers pass foreground blocks)
  int p;
  while ((p = lab[x]) != x) x = p;
  return x;
}

// both roots at once: the two searches advance in lockstep so that their loads overlap
__device__ __forceinline__ void gfind2(volatile int* lab, int& a, int& b) {
  int ea = lab[a], eb = lab[b];
  if (ea & kTag) a = ea & ~kTag;
  if (eb & kTag) b = eb & ~kTag;
  ea = lab[a], eb = lab[b];
  while (ea != a || eb != b) {
    a = ea, b = eb;
    ea = lab[a], eb = lab[b];
  }
}

__device__ __forceinline__ void gunite(int* lab, int a, int b) {
  while (true) {
    gfind2(lab, a, b);
    if (a == b) return;
    if (a < b) {
      int t = a;
      a = b;
      b = t;
    }
    int old = atomicMin(&lab[a], b);
    if (old == a) return;
    a = old;
  }
}

// 64 threads per tile: threads 0..31 = top row of the tile (checks UL,U,UR), 32..63 = left column (checks L,UL,DL)
__device__ __forceinline__ void ccl_merge_tile(const uint8_t* __restrict__ bin, size_t bin_fstride, const FrameGeom& g,
                                               int* __restrict__ labels, int tile_x, int tile_y, int fr, int t) {
  const uint8_t* b = bin + (size_t)fr * bin_fstride;
  int* lab = labels + (size_t)fr * g.nblocks;
  int bx, by;
  const bool toprow = t < 32;
  if (toprow) {
    bx = tile_x * 32 + t;
    by = tile_y * 32;
  } else {
    bx = tile_x * 32;
    by = tile_y * 32 + (t - 32);
  }
  if (bx >= g.bw || by >= g.bh) return;
  // the block's own pattern and the three neighbours across the border, loaded together (one memory latency)
  const int i = by * g.bw + bx;
  int p = block_pattern(b, g, bx, by), q0 = 0, q1 = 0, q2 = 0;
  if (toprow) {
    if (by > 0) {
      q0 = block_pattern(b, g, bx, by - 1);
      if (bx > 0) q1 = block_pattern(b, g, bx - 1, by - 1);
      if (bx < g.bw - 1) q2 = block_pattern(b, g, bx + 1, by - 1);
    }
    if (!p) return;
    if ((p & 3) && (q0 & 12)) gunite(lab, i, i - g.bw);
    if ((p & 1) && (q1 & 8)) gunite(lab, i, i - g.bw - 1);
    if ((p & 2) && (q2 & 4)) gunite(lab, i, i - g.bw + 1);
  } else {
    if (bx > 0) {
      q0 = block_pattern(b, g, bx - 1, by);
      if (by > 0) q1 = block_pattern(b, g, bx - 1, by - 1);
      if (by < g.bh - 1) q2 = block_pattern(b, g, bx - 1, by + 1);
    }
    if (!p) return;
    if ((p & 5) && (q0 & 10)) gunite(lab, i, i - 1);
    if ((p & 1) && (q1 & 8)) gunite(lab, i, i - g.bw - 1);
    if ((p & 4) && (q2 & 2)) gunite(lab, i, i + g.bw - 1);
  }
}

// 64 threads per listed tile.  Without TileHint flags (exact == false) the list holds every tile: one whose first segment
// flag is clear holds no foreground (its first block row is always inside the frame) and has nothing on its border.
__global__ void __launch_bounds__(64) ccl_merge_kernel(const uint8_t* __restrict__ bin, size_t bin_fstride, FrameGeom g,
                                                       int* __restrict__ labels, const uint8_t* __restrict__ seg_flags,
                                                       int seg_pitch, size_t seg_fstride, const int* __restrict__ tile_list,
                                                       const int* __restrict__ tile_count, bool exact) {
  const int count = *tile_count;
  for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
    const int tile = tile_list[idx];
    const int fr = tile >> 16, ty = (tile >> 8) & 255, tx = tile & 255;
    if (!exact && !seg_flags[(size_t)fr * seg_fstride + (size_t)(ty * 32) * seg_pitch + tx]) continue;
    ccl_merge_tile(bin, bin_fstride, g, labels, tx, ty, fr, threadIdx.x);
  }
}

// 256 threads per listed tile, thread t owns the four blocks (4*(t&7) .. +3) of tile row t>>3 like ccl_local.
// roots_tmp[segment * 32 + k] = k-th global root (ascending) of the 32-block row segment (block row, tile column);
// the segment's flag byte becomes 0x80 | number of roots.
__global__ void __launch_bounds__(256) ccl_final_kernel(FrameGeom g, int* __restrict__ labels, int* __restrict__ st_area,
                                                        int* __restrict__ st_x0, int* __restrict__ st_y0,
                                                        int* __restrict__ st_x1, int* __restrict__ st_y1,
                                                        int* __restrict__ roots_tmp, uint8_t* __restrict__ seg_flags,
                                                        int seg_pitch, size_t seg_fstride, const int* __restrict__ tile_list,
                                                        const int* __restrict__ tile_count, bool exact) {
  constexpr unsigned kFull = 0xffffffffu;
  const int tq = threadIdx.x & 7, trow = threadIdx.x >> 3;
  const int count = *tile_count;
  for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
    const int tile = tile_list[idx];
    const int fr = tile >> 16, ty = (tile >> 8) & 255, tx = tile & 255;
    uint8_t* sf = seg_flags + (size_t)fr * seg_fstride;
    if (!exact && !sf[(size_t)(ty * 32) * seg_pitch + tx]) continue;  // no foreground after all
    const int by = ty * 32 + trow, bx0 = tx * 32 + 4 * tq, i0 = by * g.bw + bx0;
    const size_t base = (size_t)fr * g.nblocks;
    int* lab = labels + base;
    volatile int* labv = lab;
    int e4[4] = {-1, -1, -1, -1};
    if (by < g.bh) {
      if (bx0 + 3 < g.bw && ((base + i0) & 3) == 0) {
        const int4 v = *reinterpret_cast<const int4*>(lab + i0);
        e4[0] = v.x, e4[1] = v.y, e4[2] = v.z, e4[3] = v.w;
      } else {
        for (int k = 0; k < 4; ++k)
          if (bx0 + k < g.bw) e4[k] = lab[i0 + k];
      }
    }
    int nroot = 0;
    unsigned rootmask = 0;
    if ((e4[0] & e4[1] & e4[2] & e4[3]) >= 0) {  // at least one foreground block (labels are >= 0, background is -1)
      // The four root searches advance in lockstep, so that their loads overlap: x = current node, nx = lab[x].
      // The block's own entry is already here; a tagged entry names the tile-local root to start from.
      int x[4], nx[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e4[k];
        x[k] = (e >= 0 && (e & kTag)) ? (e & ~kTag) : i0 + k;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) nx[k] = e4[k] < 0 ? x[k] : ((e4[k] & kTag) ? labv[x[k]] : e4[k]);
      while (true) {
        unsigned moved = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (nx[k] != x[k]) {
            x[k] = nx[k];
            moved |= 1u << k;
          }
        if (!moved) break;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (moved & (1u << k)) nx[k] = labv[x[k]];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e4[k], i = i0 + k, r = x[k];
        if (e < 0) continue;
        if (e & kTag) {
          lab[i] = r;
        } else if (r != i) {
          // a tile-local root that was merged into another tree: fold its partial stats into the global root
          atomicAdd(&st_area[base + r], st_area[base + i]);
          atomicMin(&st_x0[base + r], st_x0[base + i]);
          atomicMin(&st_y0[base + r], st_y0[base + i]);
          atomicMax(&st_x1[base + r], st_x1[base + i]);
          atomicMax(&st_y1[base + r], st_y1[base + i]);
          lab[i] = r;
        } else {
          rootmask |= 1u << k;
          ++nroot;
        }
      }
    }
    int inc = nroot;  // inclusive prefix of the root counts over the eight lanes of the row
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const int n = __shfl_up_sync(kFull, inc, o, 8);
      if (tq >= o) inc += n;
    }
    if (by < g.bh) {
      const size_t seg = (size_t)by * seg_pitch + tx;
      int* out = roots_tmp + ((size_t)fr * seg_fstride + seg) * 32;
      int pos = inc - nroot;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (rootmask & (1u << k)) out[pos++] = i0 + k;
      if (tq == 7) sf[seg] = (uint8_t)(0x80 | inc);
    }
  }
}

__device__ __forceinline__ int block_exclusive_scan_1024(int v, int* total, int* sh /*33 ints*/) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();
  if (lane == 31) sh[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = sh[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    sh[lane] = winc - w;
    if (lane == 31) sh[32] = winc;
  }
  __syncthreads();
  *total = sh[32];
  return sh[wid] + inc - v;
}

// One CTA per frame.  legal[frame][k] = {root, area, x0, y0, x1, y1} in ascending root order (= OpenCV label order).
// Thread t owns a run of consecutive row segments (raster order: block row, tile column), so one scan orders everything.
__global__ void __launch_bounds__(1024) ccl_list_kernel(FrameGeom g, const int* __restrict__ st_area,
                                                        const int* __restrict__ st_x0, const int* __restrict__ st_y0,
                                                        const int* __restrict__ st_x1, const int* __restrict__ st_y1,
                                                        const int* __restrict__ roots_tmp, const uint8_t* __restrict__ seg_flags,
                                                        int segs, int* __restrict__ legal, int legal_cap,
                                                        int* __restrict__ counters /* [frame][4]: n_comp, n_legal, overflow */) {
  __shared__ int sh[33];
  const int t = threadIdx.x, fr = blockIdx.x;
  const size_t base = (size_t)fr * g.nblocks;
  const uint8_t* sf = seg_flags + (size_t)fr * segs;
  const int* roots = roots_tmp + (size_t)fr * segs * 32;
  const int per = (segs + 1023) / 1024;
  const int s_lo = min(t * per, segs), s_hi = min(s_lo + per, segs);
  // Almost all segments are empty: the flag bytes are read sixteen at a time (independent loads), then only the
  // segments that hold roots are visited.
  int c = 0, lc = 0;
  for (int s0 = s_lo; s0 < s_hi; s0 += 16) {
    unsigned nz = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int f = s0 + j < s_hi ? sf[s0 + j] : 0;
      if (f & 0x80) nz |= 1u << j;
    }
    while (nz) {
      const int s = s0 + __ffs(nz) - 1;
      nz &= nz - 1;
      const int n = sf[s] & 0x7f;
      c += n;
      for (int k = 0; k < n; ++k) {
        const int a = st_area[base + roots[(size_t)s * 32 + k]];
        lc += (a >= kAreaMin && a <= g.area_max) ? 1 : 0;
      }
    }
  }
  int tot_c, tot_l;
  (void)block_exclusive_scan_1024(c, &tot_c, sh);
  int o = block_exclusive_scan_1024(lc, &tot_l, sh);
  if (lc) {
    for (int s = s_lo; s < s_hi; ++s) {
      const int f = sf[s];
      if (!(f & 0x80)) continue;
      const int n = f & 0x7f;
      for (int k = 0; k < n; ++k) {
        const int r = roots[(size_t)s * 32 + k];
        const int a = st_area[base + r];
        if (a >= kAreaMin && a <= g.area_max) {
          if (o < legal_cap) {
            int* dst = legal + ((size_t)fr * legal_cap + o) * 6;
            dst[0] = r;
            dst[1] = a;
            dst[2] = st_x0[base + r];
            dst[3] = st_y0[base + r];
            dst[4] = st_x1[base + r];
            dst[5] = st_y1[base + r];
          }
          ++o;
        }
      }
    }
  }
  if (t == 0) {
    counters[fr * 4 + 0] = tot_c;
    counters[fr * 4 + 1] = tot_l < legal_cap ? tot_l : legal_cap;
    counters[fr * 4 + 2] = tot_l > legal_cap ? 1 : 0;
    counters[fr * 4 + 3] = 0;
  }
}

size_t ccl_seg_flag_bytes(const FrameGeom& g) { return (size_t)((g.bw + 31) / 32) * g.bh; }
size_t ccl_tile_hint_bytes(const FrameGeom& g) { return (size_t)((g.bw + 31) / 32) * ((g.bh + 31) / 32); }
TileHint ccl_tile_hint(const FrameGeom& g, uint8_t* buf) {
  TileHint h;
  h.any = buf;
  h.pitch = (g.bw + 31) / 32;
  h.fstride = (int)ccl_tile_hint_bytes(g);
  return h;
}

size_t ccl_roots_ints(const FrameGeom& g) { return ccl_seg_flag_bytes(g) * 32; }
size_t ccl_tile_list_ints(const FrameGeom& g, int frames) { return 4 + ccl_tile_hint_bytes(g) * (size_t)frames; }

int launch_ccl(const uint8_t* bin, size_t bin_fstride, int n, const FrameGeom& g, int* labels, int* st_area, int* st_x0,
               int* st_y0, int* st_x1, int* st_y1, int* roots_tmp, int* tile_list, uint8_t* seg_flags, const uint8_t* tile_any,
               int* legal, int legal_cap, int* counters, cudaStream_t stream, int* launches) {
  const int tiles_x = (g.bw + 31) / 32, tiles_y = (g.bh + 31) / 32, ntiles = tiles_x * tiles_y * n;
  const int seg_pitch = tiles_x;
  static int sms_cache[64] = {0};
  int dev = 0;
  CTAG_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!sms_cache[dev]) CTAG_CUDA_CHECK(cudaDeviceGetAttribute(&sms_cache[dev], cudaDevAttrMultiProcessorCount, dev));
  const int sms = sms_cache[dev];
  const size_t seg_fstride = ccl_seg_flag_bytes(g);
  int* tile_count = tile_list;  // [0] = number of entries, the entries start at [4]
  int* tiles = tile_list + 4;
  CTAG_CUDA_CHECK(cudaMemsetAsync(seg_flags, 0, seg_fstride * n, stream));
  CTAG_CUDA_CHECK(cudaMemsetAsync(tile_count, 0, 16, stream));
  ccl_tiles_kernel<<<(ntiles + 255) / 256, 256, 0, stream>>>(tile_any, ntiles, tiles_x, tiles_y, tiles, tile_count);
  // one wave of resident CTAs each (ccl_local: 8 per SM, 26 KB of shared memory each)
  ccl_local_kernel<<<min(ntiles, 8 * sms), 256, 0, stream>>>(bin, bin_fstride, g, labels, st_area, st_x0, st_y0, st_x1, st_y1,
                                                              seg_flags, seg_pitch, seg_fstride, tiles, tile_count);
  ccl_merge_kernel<<<min(ntiles, 16 * sms), 64, 0, stream>>>(bin, bin_fstride, g, labels, seg_flags, seg_pitch, seg_fstride, tiles,
                                                              tile_count, tile_any != nullptr);
  ccl_final_kernel<<<min(ntiles, 8 * sms), 256, 0, stream>>>(g, labels, st_area, st_x0, st_y0, st_x1, st_y1, roots_tmp, seg_flags,
                                                              seg_pitch, seg_fstride, tiles, tile_count, tile_any != nullptr);
  ccl_list_kernel<<<n, 1024, 0, stream>>>(g, st_area, st_x0, st_y0, st_x1, st_y1, roots_tmp, seg_flags, (int)seg_fstride, legal,
                                          legal_cap, counters);
  CTAG_CUDA_CHECK(cudaGetLastError());
  if (launches) *launches += 5;
  return CTAG_OK;
}

}  // namespace ctag
