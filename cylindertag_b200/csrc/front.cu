// K1: fused dense front end  (reference rows a0-a3 of SURVEY 8a)
//
//   a0  BGR -> gray                 main.cpp:36,54          (cvtColor, integer weights 3735/19235/9798 >> 15)
//   a1  2x bicubic decimation       CylinderTag.cpp:79      (separable taps (-3,19,19,-3)/32, RNE, saturate)
//   a2  u8 -> f32 * 1/255           CylinderTag.cpp:80      (never materialised: applied to tile extrema only)
//   a3  adaptiveThreshold           corner_detector.cpp:28-79 (5x5 tile min/max, 3x3 tile dilation, threshold)
//
// One tile = a 80x40 half-resolution patch (16x8 threshold tiles).  It needs the 18x10 surrounding tiles, i.e. a 90x50
// half-res patch, i.e. a 192x102 full-res region which TMA stages in shared memory (cp.async.bulk.tensor.3d, zero fill
// outside the image; replicate borders are patched afterwards).  The half-res gray image, the float image and the
// tile extrema never touch HBM.
//
// Three families of kernels share the stencil / threshold phases of front_phases.cuh:
//   front_bgr_slide_kernel / front_gray_slide_kernel   the production kernels: persistent CTAs walk RUNS of tiles down a
//       tile column and keep what vertically adjacent tiles share in shared memory (see the comments in front of them)
//   front_kernel<C>   the same phases on independent tiles taken with a stride (whole 192x102 region staged per tile;
//       the TMA load of a CTA's next tile is issued as soon as the staging buffer is free); kept for A/B runs
//       (CTAG_FRONT_NOSLIDE=1)
//
// HBM traffic per frame (N = w*h): BGR input 3N read + N gray write + N/4 binary write; gray input N + N/4.
#include "common.cuh"
#include "front_phases.cuh"
#include "kernels.cuh"
#include <cstdlib>

namespace ctag {

namespace front {
template <int C>
struct Layout {
  // C==3: [bgr 3 boxes | P | small]   gray tile aliases box 0 (built through registers once every thread has read its
  //                                   BGR bytes), HT aliases box 1: 67.5 KB per CTA, three CTAs per SM
  // C==1: [g0 | g1 | HT | P | small]  (gray tiles double-buffered)
  static constexpr int bgr = 0;
  static constexpr int g = 0;                        // C==1: g0 at 0, g1 at BOX
  static constexpr int h = (C == 3) ? BOX : 2 * BOX;
  static constexpr int after_h = (C == 3) ? 3 * BOX : h + H_BYTES;
  static constexpr int p = after_h;
  static constexpr int small_ = p + P_BYTES;
  static constexpr int tmin = small_;               // CTY*CTX bytes
  static constexpr int tmax = small_ + 192;         // CTY*CTX bytes
  static constexpr int cmn = small_ + 384;          // CTY x 96 column minima
  static constexpr int cmx = small_ + 384 + 960;    // CTY x 96 column maxima
  static constexpr int vthr = small_ + 2304;        // OTY*OTX bytes
  static constexpr int thr16 = small_ + 2304 + 128; // OTY x OW threshold bytes (one per owned pixel column)
  static constexpr int mbar = small_ + 2304 + 128 + 640;  // 2 x 8 bytes
  static constexpr int total = small_ + 2304 + 128 + 640 + 16;
};
}  // namespace front

struct TileGrid {
  int tiles_x, tiles_y, tiles_per_frame, ntiles;
  uint32_t m_frame, m_row;  // floor(2^32 / tiles_per_frame), floor(2^32 / tiles_x)
};

// tile -> (frame, tile row, tile column) without integer division: the multiply-high quotient with the floored
// reciprocal is at most one too small for any 32-bit dividend, one compare fixes it.
__device__ __forceinline__ void tile_coords(const TileGrid& tg, int tile, int& fr, int& cy, int& cx) {
  uint32_t q = __umulhi((uint32_t)tile, tg.m_frame);
  uint32_t rem = (uint32_t)tile - q * (uint32_t)tg.tiles_per_frame;
  if (rem >= (uint32_t)tg.tiles_per_frame) { ++q; rem -= (uint32_t)tg.tiles_per_frame; }
  uint32_t r = __umulhi(rem, tg.m_row);
  uint32_t c = rem - r * (uint32_t)tg.tiles_x;
  if (c >= (uint32_t)tg.tiles_x) { ++r; c -= (uint32_t)tg.tiles_x; }
  fr = (int)q; cy = (int)r; cx = (int)c;
}

template <int C>
__device__ __forceinline__ void issue_tile_load(const CUtensorMap* tmap, uint32_t mbar, uint32_t dst, const TileGrid& tg,
                                                int tile) {
  using namespace front;
  int fr, cy, cx;
  tile_coords(tg, tile, fr, cy, cx);
  const int x0r = 2 * OW * cx - 16, y0r = 2 * OH * cy - 11;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(C * BOX) : "memory");
#pragma unroll
  for (int b = 0; b < C; ++b) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(dst + b * BOX), "l"(tmap), "r"(mbar), "r"(C * x0r + RW * b), "r"(y0r), "r"(fr)
        : "memory");
  }
}

template <int C>
__global__ void __launch_bounds__(front::NT, 3) front_kernel(const __grid_constant__ CUtensorMap tmap, FrameGeom geo, TileGrid tg,
                                                          uint8_t* __restrict__ gray_out, size_t gray_fstride,
                                                          uint8_t* __restrict__ bin_out, size_t bin_fstride, TileHint hint) {
  using namespace front;
  using L = Layout<C>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const uint32_t mbar0 = smem_u32(smem + L::mbar);
  uint32_t* HT = reinterpret_cast<uint32_t*>(smem + L::h);
  uint8_t* P = smem + L::p;
  uint8_t* tmin = smem + L::tmin;
  uint8_t* tmax = smem + L::tmax;
  uint8_t* cmn = smem + L::cmn;
  uint8_t* cmx = smem + L::cmx;
  uint8_t* thr16 = smem + L::thr16;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  int tile = blockIdx.x;
  if (tid == 0 && tile < tg.ntiles) issue_tile_load<C>(&tmap, mbar0, smem_u32(smem + (C == 3 ? L::bgr : L::g)), tg, tile);

  // C==3: one staging buffer + one barrier, parity flips per tile.  C==1: buffers/barriers alternate, parity flips
  // every second tile.
  for (int it = 0; tile < tg.ntiles; tile += gridDim.x, ++it) {
    int fr, cy, cx;
    tile_coords(tg, tile, fr, cy, cx);
    const int x0r = 2 * OW * cx - 16;  // full-res x of region column 0
    const int y0r = 2 * OH * cy - 11;  // full-res y of region row 0
    const int buf = (C == 1) ? (it & 1) : 0;
    uint8_t* g = smem + L::g + (C == 1 ? buf * BOX : 0);
    const int next = tile + gridDim.x;
    if (C == 1) {
      // prefetch the next tile into the other gray buffer (its last reader, phase B of the previous tile, is behind
      // at least one barrier for every thread)
      if (tid == 0 && next < tg.ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_tile_load<C>(&tmap, mbar0 + 8 * (buf ^ 1), smem_u32(smem + L::g + (buf ^ 1) * BOX), tg, next);
      }
      mbar_wait(mbar0 + 8 * buf, (it >> 1) & 1);
    } else {
      mbar_wait(mbar0, it & 1);
    }

    // ---- phase A: BGR -> gray (whole region) + store of the owned gray pixels --------------------------------
    if (C == 3) {
      {
        // thread -> (16-pixel group gq, row).  Eight consecutive lanes take the four groups of one staging box on two
        // consecutive rows: their 16-byte slots (3*(gq&3) + 4*row) mod 8 are all different, so the 128-bit loads from
        // the box and the 128-bit store into the gray tile are free of bank conflicts.
        const int rest = tid >> 3;
        const int gq = (rest % 3) * 4 + (tid & 3);
        const int row = 2 * (rest / 3) + ((tid >> 2) & 1);  // rows row, row+32, row+64, row+96
        const uint8_t* src = smem + L::bgr + (gq >> 2) * BOX + (gq & 3) * 48 + row * RW;
        uint8_t* dst = g + gq * 16 + row * RW;
        const int x = x0r + gq * 16;
        const bool own_col = gq >= 1 && gq <= 10 && x < geo.w;
        // store address = CTA-uniform base (kept in uniform registers) + a 32-bit per-thread offset
        uint8_t* tile_base = gray_out + (size_t)fr * gray_fstride + (ptrdiff_t)y0r * geo.gpitch + x0r;
        asm volatile("" : "+l"(tile_base));  // keep the pointer live: recomputing the 64-bit products per store costs more
        const uint32_t toff = (uint32_t)row * (uint32_t)geo.gpitch + (uint32_t)gq * 16u;
        constexpr int RS = NT / 12;  // row step
        uint4 o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = row + u * RS;
          if (r < RH) {
            const uint4* s4 = reinterpret_cast<const uint4*>(src + u * RS * RW);
            const uint4 a = s4[0], b = s4[1], c = s4[2];
            o[u].x = gray4(a.x, a.y, a.z);
            o[u].y = gray4(a.w, b.x, b.y);
            o[u].z = gray4(b.z, b.w, c.x);
            o[u].w = gray4(c.y, c.z, c.w);
            if (own_col && r >= 11 && r < 11 + 2 * OH && y0r + r < geo.h)
              *reinterpret_cast<uint4*>(tile_base + (toff + (uint32_t)(u * RS) * (uint32_t)geo.gpitch)) = o[u];
          }
        }
        __syncthreads();  // every BGR byte has been read: the gray tile may overwrite box 0
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (row + u * RS < RH) *reinterpret_cast<uint4*>(dst + u * RS * RW) = o[u];
      }
      __syncthreads();
    }

    phase_border(g, geo, cx, cy, x0r, y0r, tid);
    phase_horizontal(g, HT, tid);
    __syncthreads();
    const bool edge_cta = (cx == 0) || (cy == 0) || (OW * cx + OW + 5 > geo.hw) || (OH * cy + OH + 5 > geo.hh);
    phase_vertical_extrema(HT, P, cmn, cmx, geo, cx, cy, edge_cta, 0, tid);
    __syncthreads();
    if (C == 3) {
      // the gray tile (box 0) and HT (box 1) are dead: load this CTA's next tile behind phases D-F; the other CTAs of
      // the SM cover what remains of the latency
      if (tid == 0 && next < tg.ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_tile_load<C>(&tmap, mbar0, smem_u32(smem + L::bgr), tg, next);
      }
    }

    phase_tile_extrema(cmn, cmx, tmin, tmax, tid);
    __syncthreads();
    phase_threshold(tmin, tmax, thr16, geo, cx, cy, tid);
    __syncthreads();
    phase_compare_store(P, thr16, bin_out, bin_fstride, hint, geo, fr, cx, cy, tid);
    // no barrier needed here: the next iteration touches g/bgr (free since phase B / phase A) and reaches HT, P and
    // the small arrays only after further barriers
  }
}

// =====================================================================================================================
// BGR input, sliding variant.  Vertically adjacent tiles share 22 of their 102 region rows (11 horizontal-pass row pairs,
// 10 of their 50 half-res rows).  A CTA therefore walks RUNS of consecutive tiles of one tile column top down and keeps
// what the tile below needs in shared memory: its 10 shared half-res rows, 2 rows of column extrema and the ONE
// horizontal-pass row pair (50 -> 10) its first new half-res row reads.  Only the 80 new full-res rows are staged (TMA)
// and converted, only the 40 new row pairs go through the horizontal and the vertical pass.  The first tile of a run / of
// a column loads its 22 extra rows with a second, 22-row tensor map; a tile whose run goes on also writes the 11 gray
// output rows that belong to the tile below (it has them, the tile below will not load them).
// Shared memory: [stage 3 x 192 x 80 | g | P | small | keep], HT aliases the staging area
// (74.4 KB per CTA, three CTAs per SM).
namespace front {
constexpr int SROWS = 80;            // rows staged per tile after the first
constexpr int OVR = RH - SROWS;      // 22 rows shared with the tile above
constexpr int SBOX = RW * SROWS;     // bytes of one colour box slot in the staging area
struct SlideLayout {
  static constexpr int stage = 0;
  static constexpr int h = 0;  // HT (RH/2 x HP words = 19584 B) lives in the staging area between phase A and phase C
  static constexpr int g = 3 * SBOX;
  static constexpr int p = g + BOX;
  static constexpr int small_ = p + P_BYTES;
  static constexpr int tmin = small_;
  static constexpr int tmax = small_ + 192;
  static constexpr int cmn = small_ + 384;
  static constexpr int cmx = small_ + 384 + 960;
  static constexpr int thr16 = small_ + 2304 + 128;
  static constexpr int mbar = small_ + 2304 + 128 + 640;
  static constexpr int keep = small_ + 2304 + 128 + 640 + 16;  // 2 x HP words: row pair 50 for the tile below
  static constexpr int total = keep + 2 * HP * 4;
};
static_assert(H_BYTES <= 3 * SBOX, "HT must fit into the staging area");
}  // namespace front

struct RunGrid {
  int tiles_x, tiles_y, tiles_per_frame, ntiles;
  uint32_t m_frame, m_col;  // floor(2^32 / tiles_per_frame), floor(2^32 / tiles_y)
  int wave_runs;            // runs per full wave (= grid size)
  int full_waves, len, tail_len, nruns;
};

// run r -> [t0, t1) in column-major tile order (frame, tile column, tile row)
__device__ __forceinline__ void run_range(const RunGrid& rg, int r, int& t0, int& t1) {
  const int full = rg.full_waves * rg.wave_runs;
  if (r < full) {
    t0 = r * rg.len;
    t1 = t0 + rg.len;
  } else {
    t0 = full * rg.len + (r - full) * rg.tail_len;
    t1 = t0 + rg.tail_len;
  }
  if (t1 > rg.ntiles) t1 = rg.ntiles;
}
__device__ __forceinline__ void run_tile_coords(const RunGrid& rg, int t, int& fr, int& cx, int& cy) {
  uint32_t q = __umulhi((uint32_t)t, rg.m_frame);
  uint32_t rem = (uint32_t)t - q * (uint32_t)rg.tiles_per_frame;
  if (rem >= (uint32_t)rg.tiles_per_frame) { ++q; rem -= (uint32_t)rg.tiles_per_frame; }
  uint32_t c = __umulhi(rem, rg.m_col);
  uint32_t r = rem - c * (uint32_t)rg.tiles_y;
  if (r >= (uint32_t)rg.tiles_y) { ++c; r -= (uint32_t)rg.tiles_y; }
  fr = (int)q; cx = (int)c; cy = (int)r;
}

// The tile a CTA works on after (run, t): one tile down in the same run (`first` = false: the rows shared with the
// tile above are still in shared memory), the top of the next tile column inside the run, or the first tile of the
// CTA's next run (both `first` = true: nothing to reuse).  `valid` = false when the CTA has no more work.
__device__ __forceinline__ void next_tile(const RunGrid& rg, int run, int t, int t_end, int fr, int cx, int cy, int& nrun, int& nt,
                                          int& nt_end, int& nfr, int& ncx, int& ncy, bool& nvalid, bool& nfirst) {
  nt = t + 1, nrun = run, nt_end = t_end;
  nfr = fr, ncx = cx, ncy = cy + 1;
  nvalid = true, nfirst = false;
  if (nt < t_end) {
    if (ncy == rg.tiles_y) {
      ncy = 0;
      nfirst = true;
      if (++ncx == rg.tiles_x) ncx = 0, ++nfr;
    }
  } else {
    nrun = run + gridDim.x;
    nfirst = true;
    nvalid = false;
    if (nrun < rg.nruns) {
      run_range(rg, nrun, nt, nt_end);
      nvalid = nt < nt_end;
      if (nvalid) run_tile_coords(rg, nt, nfr, ncx, ncy);
    }
  }
}

// rows [row0, row0 + nrows) of the region of tile (fr, cx, cy) -> staging area (three colour box slots of SBOX bytes)
__device__ __forceinline__ void issue_rows_load(const CUtensorMap* tmap, uint32_t mbar, uint32_t dst, int fr, int cx, int cy,
                                                int row0, int nrows) {
  using namespace front;
  const int x0r = 2 * OW * cx - 16, y0r = 2 * OH * cy - 11;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(3 * RW * nrows) : "memory");
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(dst + b * SBOX), "l"(tmap), "r"(mbar), "r"(3 * x0r + RW * b), "r"(y0r + row0), "r"(fr)
        : "memory");
  }
}

// the same rows into L2 only: issued a phase ahead of the real load, which then does not wait on DRAM
__device__ __forceinline__ void prefetch_rows_l2(const CUtensorMap* tmap, int fr, int cx, int cy, int row0) {
  using namespace front;
  const int x0r = 2 * OW * cx - 16, y0r = 2 * OH * cy - 11;
#pragma unroll
  for (int b = 0; b < 3; ++b)
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(3 * x0r + RW * b),
                 "r"(y0r + row0), "r"(fr)
                 : "memory");
}

// BGR -> gray for staged rows [0, nrows) -> gray tile rows [row0, row0 + nrows), and the global store of the owned ones.
// thread -> (16-pixel group gq, row): eight consecutive lanes take the four groups of one colour box on two consecutive
// rows; their 16-byte slots (3*(gq&3) + 4*row) mod 8 are all different, so the 128-bit loads from the box and the 128-bit
// store into the gray tile are free of bank conflicts.
__device__ __forceinline__ void convert_rows(const uint8_t* stage, uint8_t* g, int row0, int nrows, int tid, uint8_t* tile_base,
                                             uint32_t gpitch, int x0r, int rlim, int img_w) {
  using namespace front;
  const int rest = tid >> 3;
  const int gq = (rest % 3) * 4 + (tid & 3);
  const int row = 2 * (rest / 3) + ((tid >> 2) & 1);  // staged rows row, row+32, row+64
  const uint8_t* src = stage + (gq >> 2) * SBOX + (gq & 3) * 48 + row * RW;
  uint8_t* dst = g + (row0 + row) * RW + gq * 16;
  const bool own_col = gq >= 1 && gq <= 10 && x0r + gq * 16 < img_w;
  const uint32_t toff = (uint32_t)(row0 + row) * gpitch + (uint32_t)gq * 16u;
  constexpr int RS = NT / 12;
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int sr = row + u * RS;
    if (sr < nrows) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src + u * RS * RW);
      const uint4 a = s4[0], b = s4[1], c = s4[2];
      uint4 o;
      o.x = gray4(a.x, a.y, a.z);
      o.y = gray4(a.w, b.x, b.y);
      o.z = gray4(b.z, b.w, c.x);
      o.w = gray4(c.y, c.z, c.w);
      *reinterpret_cast<uint4*>(dst + u * RS * RW) = o;
      const int r = row0 + sr;  // region row
      if (own_col && r >= 11 && r < rlim) *reinterpret_cast<uint4*>(tile_base + (toff + (uint32_t)(u * RS) * gpitch)) = o;
    }
  }
}

__global__ void __launch_bounds__(front::NT, 3)
    front_bgr_slide_kernel(const __grid_constant__ CUtensorMap tmap_main, const __grid_constant__ CUtensorMap tmap_top, FrameGeom geo,
                           RunGrid rg, uint8_t* __restrict__ gray_out, size_t gray_fstride, uint8_t* __restrict__ bin_out,
                           size_t bin_fstride, TileHint hint) {
  using namespace front;
  using L = SlideLayout;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const uint32_t mbar = smem_u32(smem + L::mbar);
  const uint32_t stage_u32 = smem_u32(smem + L::stage);
  const uint8_t* stage = smem + L::stage;
  uint32_t* HT = reinterpret_cast<uint32_t*>(smem + L::h);
  uint8_t* g = smem + L::g;
  uint8_t* P = smem + L::p;
  uint8_t* tmin = smem + L::tmin;
  uint8_t* tmax = smem + L::tmax;
  uint8_t* cmn = smem + L::cmn;
  uint8_t* cmx = smem + L::cmx;
  uint8_t* thr16 = smem + L::thr16;
  uint32_t* keep = reinterpret_cast<uint32_t*>(smem + L::keep);

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  int run = blockIdx.x, t = 0, t_end = 0;
  if (run < rg.nruns) run_range(rg, run, t, t_end);
  bool valid = run < rg.nruns && t < t_end;
  int fr = 0, cx = 0, cy = 0;
  bool first = true;
  if (valid) {
    run_tile_coords(rg, t, fr, cx, cy);
    if (tid == 0) issue_rows_load(&tmap_top, mbar, stage_u32, fr, cx, cy, 0, OVR);
  }
  uint32_t loads = 0;  // completed-phase counter of the mbarrier (parity = loads & 1)
  uint32_t kbuf = 0;   // keep buffer this tile writes (the other one holds what the tile above left)

  while (valid) {
    // the tile after this one
    int nt, nrun, nt_end, nfr, ncx, ncy;
    bool nvalid, nfirst;
    next_tile(rg, run, t, t_end, fr, cx, cy, nrun, nt, nt_end, nfr, ncx, ncy, nvalid, nfirst);

    const int x0r = 2 * OW * cx - 16;  // full-res x of region column 0
    const int y0r = 2 * OH * cy - 11;  // full-res y of region row 0
    uint8_t* tile_base = gray_out + (size_t)fr * gray_fstride + (ptrdiff_t)y0r * geo.gpitch + x0r;
    asm volatile("" : "+l"(tile_base));  // keep the pointer live: recomputing the 64-bit products per store costs more
    // gray output rows of this tile: region rows 11..90, plus rows 91..101 (rows 11..21 of the tile below) when the
    // run goes on there
    const int rlim = min((nvalid && !nfirst) ? RH : 11 + 2 * OH, geo.h - y0r);

    if (!first) {
      // Reuse from the tile above: its half-res rows 40..49 are rows 0..9 here, its column extrema of tile rows 8,9 are
      // those of tile rows 0,1 (its horizontal-pass row pair 50 waits in the keep buffer).  Everybody has to be done
      // with the previous tile's threshold phases first.
      __syncthreads();
      if (tid < 70) {  // 10 patch rows of PP = 112 bytes = 70 uint4
        reinterpret_cast<uint4*>(P)[tid] = reinterpret_cast<const uint4*>(P + 40 * PP)[tid];
      } else if (tid >= 96 && tid < 96 + 48) {  // 2 x 96 bytes of cmn, 2 x 96 bytes of cmx
        const int q = tid - 96, arr = q / 24, w = q - arr * 24;
        uint32_t* base = reinterpret_cast<uint32_t*>(arr ? cmx : cmn);
        base[w] = base[8 * 24 + w];
        base[24 + w] = base[9 * 24 + w];
      }
    }

    // ---- phase A: BGR -> gray for the staged rows -----------------------------------------------------------------
    mbar_wait(mbar, loads & 1);
    ++loads;
    if (first) {
      convert_rows(stage, g, 0, OVR, tid, tile_base, (uint32_t)geo.gpitch, x0r, rlim, geo.w);
      __syncthreads();  // staging area free again
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_rows_load(&tmap_main, mbar, stage_u32, fr, cx, cy, OVR, SROWS);
      }
      mbar_wait(mbar, loads & 1);
      ++loads;
    }
    convert_rows(stage, g, OVR, SROWS, tid, tile_base, (uint32_t)geo.gpitch, x0r, rlim, geo.w);
    __syncthreads();
    // the staging area is busy (HT) until the vertical pass is over: meanwhile pull the next tile's rows into L2
    if (tid == 0 && nvalid) {
      if (nfirst) {
        prefetch_rows_l2(&tmap_top, nfr, ncx, ncy, 0);
        prefetch_rows_l2(&tmap_main, nfr, ncx, ncy, OVR);
      } else {
        prefetch_rows_l2(&tmap_main, nfr, ncx, ncy, OVR);
      }
    }

    phase_border(g, geo, cx, cy, x0r, y0r, tid, first ? 0 : OVR);
    // HT takes over the staging area; only the new row pairs (11..50) unless this is a first tile
    phase_horizontal(g, HT, tid, first ? 0 : OVR / 2);
    __syncthreads();

    // vertical taps + rounding + column extrema for the NEW tile rows (all ten for a first tile)
    const bool edge_cta = (cx == 0) || (cy == 0) || (OW * cx + OW + 5 > geo.hw) || (OH * cy + OH + 5 > geo.hh);
    phase_vertical_extrema(HT, P, cmn, cmx, geo, cx, cy, edge_cta, first ? 0 : 2, tid, first ? nullptr : keep + (kbuf ^ 1) * HP,
                           keep + kbuf * HP);
    kbuf ^= 1;
    __syncthreads();
    // HT is dead, the staging area is free: load the next tile's rows behind the threshold phases
    if (tid == 0 && nvalid) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (nfirst) issue_rows_load(&tmap_top, mbar, stage_u32, nfr, ncx, ncy, 0, OVR);
      else issue_rows_load(&tmap_main, mbar, stage_u32, nfr, ncx, ncy, OVR, SROWS);
    }

    phase_tile_extrema(cmn, cmx, tmin, tmax, tid);
    __syncthreads();
    phase_threshold(tmin, tmax, thr16, geo, cx, cy, tid);
    __syncthreads();
    phase_compare_store(P, thr16, bin_out, bin_fstride, hint, geo, fr, cx, cy, tid);

    t = nt, t_end = nt_end, run = nrun, valid = nvalid, first = nfirst;
    fr = nfr, cx = ncx, cy = ncy;
  }
}

// =====================================================================================================================
// Gray input, sliding variant (detect()'s own contract: the caller hands over a gray frame).  Same walk as the BGR
// kernel, but nothing is converted: TMA writes the 80 new rows straight into the gray tile, which is dead after the
// horizontal pass, so the next tile's rows are loaded behind everything that follows.
// Shared memory: [g | HT | P | small | keep] = 47.5 KB per CTA, four CTAs per SM.
namespace front {
struct GraySlideLayout {
  static constexpr int g = 0;
  static constexpr int h = BOX;
  static constexpr int p = BOX + H_BYTES;
  static constexpr int small_ = p + P_BYTES;
  static constexpr int tmin = small_ + S_TMIN;
  static constexpr int tmax = small_ + S_TMAX;
  static constexpr int cmn = small_ + S_CMN;
  static constexpr int cmx = small_ + S_CMX;
  static constexpr int thr16 = small_ + S_THR16;
  static constexpr int mbar = small_ + S_MBAR;
  static constexpr int keep = small_ + S_BYTES;  // 2 x HP words: row pair 50 for the tile below
  static constexpr int total = keep + 2 * HP * 4;
};
}  // namespace front

// rows of the gray region of tile (fr, cx, cy) -> gray tile: all 102 (two boxes) for the first tile of a run, else the
// 80 new ones; one mbarrier phase either way
__device__ __forceinline__ void issue_gray_rows(const CUtensorMap* tmap_main, const CUtensorMap* tmap_top, uint32_t mbar,
                                                uint32_t g_u32, int fr, int cx, int cy, bool first) {
  using namespace front;
  const int x0r = 2 * OW * cx - 16, y0r = 2 * OH * cy - 11;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(RW * (first ? RH : SROWS)) : "memory");
  if (first)
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(g_u32), "l"(tmap_top), "r"(mbar), "r"(x0r), "r"(y0r), "r"(fr)
        : "memory");
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(g_u32 + OVR * RW), "l"(tmap_main), "r"(mbar), "r"(x0r), "r"(y0r + OVR), "r"(fr)
      : "memory");
}

__global__ void __launch_bounds__(front::NT, 4)
    front_gray_slide_kernel(const __grid_constant__ CUtensorMap tmap_main, const __grid_constant__ CUtensorMap tmap_top, FrameGeom geo,
                            RunGrid rg, uint8_t* __restrict__ bin_out, size_t bin_fstride, TileHint hint) {
  using namespace front;
  using L = GraySlideLayout;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const uint32_t mbar = smem_u32(smem + L::mbar);
  const uint32_t g_u32 = smem_u32(smem + L::g);
  uint8_t* g = smem + L::g;
  uint32_t* HT = reinterpret_cast<uint32_t*>(smem + L::h);
  uint8_t* P = smem + L::p;
  uint8_t* tmin = smem + L::tmin;
  uint8_t* tmax = smem + L::tmax;
  uint8_t* cmn = smem + L::cmn;
  uint8_t* cmx = smem + L::cmx;
  uint8_t* thr16 = smem + L::thr16;
  uint32_t* keep = reinterpret_cast<uint32_t*>(smem + L::keep);
  uint32_t kbuf = 0;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  int run = blockIdx.x, t = 0, t_end = 0;
  if (run < rg.nruns) run_range(rg, run, t, t_end);
  bool valid = run < rg.nruns && t < t_end;
  int fr = 0, cx = 0, cy = 0;
  bool first = true;
  if (valid) {
    run_tile_coords(rg, t, fr, cx, cy);
    if (tid == 0) issue_gray_rows(&tmap_main, &tmap_top, mbar, g_u32, fr, cx, cy, true);
  }
  uint32_t loads = 0;  // completed-phase counter of the mbarrier (parity = loads & 1)

  while (valid) {
    // the tile after this one
    int nt, nrun, nt_end, nfr, ncx, ncy;
    bool nvalid, nfirst;
    next_tile(rg, run, t, t_end, fr, cx, cy, nrun, nt, nt_end, nfr, ncx, ncy, nvalid, nfirst);
    const int x0r = 2 * OW * cx - 16, y0r = 2 * OH * cy - 11;

    if (!first) {
      // from the tile above: half-res rows 40..49 -> 0..9, column extrema of tile rows 8,9 -> 0,1 (its HT row pair 50
      // waits in the keep buffer).  Its threshold phases have to be over.
      __syncthreads();
      if (tid < 70) {
        reinterpret_cast<uint4*>(P)[tid] = reinterpret_cast<const uint4*>(P + 40 * PP)[tid];
      } else if (tid >= 96 && tid < 96 + 48) {
        const int q = tid - 96, arr = q / 24, w = q - arr * 24;
        uint32_t* base = reinterpret_cast<uint32_t*>(arr ? cmx : cmn);
        base[w] = base[8 * 24 + w];
        base[24 + w] = base[9 * 24 + w];
      }
    }
    mbar_wait(mbar, loads & 1);
    ++loads;

    phase_border(g, geo, cx, cy, x0r, y0r, tid, first ? 0 : OVR);
    phase_horizontal(g, HT, tid, first ? 0 : OVR / 2);
    __syncthreads();
    // the gray tile is dead: the next tile's rows arrive behind the rest of this one
    if (tid == 0 && nvalid) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue_gray_rows(&tmap_main, &tmap_top, mbar, g_u32, nfr, ncx, ncy, nfirst);
    }

    const bool edge_cta = (cx == 0) || (cy == 0) || (OW * cx + OW + 5 > geo.hw) || (OH * cy + OH + 5 > geo.hh);
    phase_vertical_extrema(HT, P, cmn, cmx, geo, cx, cy, edge_cta, first ? 0 : 2, tid, first ? nullptr : keep + (kbuf ^ 1) * HP,
                           keep + kbuf * HP);
    kbuf ^= 1;
    __syncthreads();
    phase_tile_extrema(cmn, cmx, tmin, tmax, tid);
    __syncthreads();
    phase_threshold(tmin, tmax, thr16, geo, cx, cy, tid);
    __syncthreads();
    phase_compare_store(P, thr16, bin_out, bin_fstride, hint, geo, fr, cx, cy, tid);

    t = nt, t_end = nt_end, run = nrun, valid = nvalid, first = nfirst;
    fr = nfr, cx = ncx, cy = ncy;
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int front_smem_bytes(int channels) {
  return channels == 3 ? front::Layout<3>::total : front::Layout<1>::total;
}

// Resident CTAs of `kernel` on the current device (SMs x CTAs per SM).  The shared-memory opt-in and the occupancy query
// are done once per kernel and device, not per batch: they cost host time in front of every launch otherwise.
static int resident_ctas(const void* kernel, int smem_bytes, int slot, int* out) {
  constexpr int kMaxDev = 64, kKernels = 4;
  static int cache[kKernels][kMaxDev];  // 0 = not yet known
  int dev = 0;
  CTAG_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev || !cache[slot][dev]) {
    int sms = 0, per_sm = 0;
    CTAG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CTAG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CTAG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, front::NT, smem_bytes));
    if (per_sm < 1) per_sm = 1;
    static const int cap = getenv("CTAG_FRONT_CTAS") ? atoi(getenv("CTAG_FRONT_CTAS")) : 0;  // A/B knob, read once
    if (cap >= 1 && cap < per_sm) per_sm = cap;
    if (dev < 0 || dev >= kMaxDev) {
      *out = sms * per_sm;
      return CTAG_OK;
    }
    cache[slot][dev] = sms * per_sm;
  }
  *out = cache[slot][dev];
  return CTAG_OK;
}

template <int C>
static int launch_front_t(const CUtensorMap& tmap, int n, const FrameGeom& geo, uint8_t* gray_out, size_t gray_fstride,
                          uint8_t* bin_out, size_t bin_fstride, TileHint hint, cudaStream_t stream, cudaEvent_t ev_start) {
  using namespace front;
  int grid = 0;
  const int rc = resident_ctas((const void*)front_kernel<C>, Layout<C>::total, C == 3 ? 0 : 1, &grid);
  if (rc != CTAG_OK) return rc;
  TileGrid tg;
  tg.tiles_x = (geo.hw + OW - 1) / OW;
  tg.tiles_y = (geo.hh + OH - 1) / OH;
  tg.tiles_per_frame = tg.tiles_x * tg.tiles_y;
  tg.ntiles = tg.tiles_per_frame * n;
  tg.m_frame = tg.tiles_per_frame > 1 ? (uint32_t)((1ull << 32) / (uint64_t)tg.tiles_per_frame) : 0xFFFFFFFFu;
  tg.m_row = tg.tiles_x > 1 ? (uint32_t)((1ull << 32) / (uint64_t)tg.tiles_x) : 0xFFFFFFFFu;
  if (grid > tg.ntiles) grid = tg.ntiles;  // persistent: one wave of resident CTAs
  if (ev_start) CTAG_CUDA_CHECK(cudaEventRecord(ev_start, stream));
  front_kernel<C><<<grid, NT, Layout<C>::total, stream>>>(tmap, geo, tg, gray_out, gray_fstride, bin_out, bin_fstride, hint);
  CTAG_CUDA_CHECK(cudaGetLastError());
  return CTAG_OK;
}

static int make_run_grid(const void* kernel, int smem_bytes, int slot, int n, const FrameGeom& geo, RunGrid* out, int* grid_out) {
  using namespace front;
  int grid = 0;
  const int rc = resident_ctas(kernel, smem_bytes, slot, &grid);
  if (rc != CTAG_OK) return rc;
  RunGrid rg;
  rg.tiles_x = (geo.hw + OW - 1) / OW;
  rg.tiles_y = (geo.hh + OH - 1) / OH;
  rg.tiles_per_frame = rg.tiles_x * rg.tiles_y;
  rg.ntiles = rg.tiles_per_frame * n;
  rg.m_frame = rg.tiles_per_frame > 1 ? (uint32_t)((1ull << 32) / (uint64_t)rg.tiles_per_frame) : 0xFFFFFFFFu;
  rg.m_col = rg.tiles_y > 1 ? (uint32_t)((1ull << 32) / (uint64_t)rg.tiles_y) : 0xFFFFFFFFu;
  // Runs of consecutive tiles (column-major order: a run walks down a tile column) are dealt round-robin to one wave of
  // resident CTAs, so neighbouring CTAs work on neighbouring columns at the same time (their halo columns meet in L2).
  // Full waves use runs of kRun tiles; what is left is split evenly over the CTAs so that everybody finishes together.
  const int kRun = rg.tiles_y;  // whole tile columns: one 22-row top load per column
  if (grid > rg.ntiles) grid = rg.ntiles;
  rg.wave_runs = grid;
  rg.len = kRun;
  rg.full_waves = rg.ntiles / (grid * kRun);
  const int rest = rg.ntiles - rg.full_waves * grid * kRun;
  rg.tail_len = rest > 0 ? (rest + grid - 1) / grid : 1;
  rg.nruns = rg.full_waves * grid + (rest > 0 ? (rest + rg.tail_len - 1) / rg.tail_len : 0);
  *out = rg;
  *grid_out = grid;
  return CTAG_OK;
}

static int launch_front_slide(const CUtensorMap& tmap_main, const CUtensorMap& tmap_top, int n, const FrameGeom& geo, int channels,
                              uint8_t* gray_out, size_t gray_fstride, uint8_t* bin_out, size_t bin_fstride, TileHint hint,
                              cudaStream_t stream, cudaEvent_t ev_start) {
  using namespace front;
  RunGrid rg;
  int grid = 0;
  if (channels == 3) {
    const int rc = make_run_grid((const void*)front_bgr_slide_kernel, SlideLayout::total, 2, n, geo, &rg, &grid);
    if (rc != CTAG_OK) return rc;
    if (ev_start) CTAG_CUDA_CHECK(cudaEventRecord(ev_start, stream));
    front_bgr_slide_kernel<<<grid, NT, SlideLayout::total, stream>>>(tmap_main, tmap_top, geo, rg, gray_out, gray_fstride, bin_out,
                                                                    bin_fstride, hint);
  } else {
    const int rc = make_run_grid((const void*)front_gray_slide_kernel, GraySlideLayout::total, 3, n, geo, &rg, &grid);
    if (rc != CTAG_OK) return rc;
    if (ev_start) CTAG_CUDA_CHECK(cudaEventRecord(ev_start, stream));
    front_gray_slide_kernel<<<grid, NT, GraySlideLayout::total, stream>>>(tmap_main, tmap_top, geo, rg, bin_out, bin_fstride, hint);
  }
  CTAG_CUDA_CHECK(cudaGetLastError());
  return CTAG_OK;
}

int launch_front(const void* frames_dev, int n, const FrameGeom& geo, int channels, size_t pitch, size_t frame_stride,
                 uint8_t* gray_out, size_t gray_fstride, uint8_t* bin_out, size_t bin_fstride, TileHint hint,
                 cudaStream_t stream, cudaEvent_t ev_start) {
  using namespace front;
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_last_error_text("cuTensorMapEncodeTiled entry point not available");
    return CTAG_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(frames_dev) & 15) || (pitch & 15) || (frame_stride & 15)) return CTAG_ERR_ALIGNMENT;
  CUtensorMap tmap;
  cuuint64_t dims[3] = {(cuuint64_t)geo.w * channels, (cuuint64_t)geo.h, (cuuint64_t)n};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
  cuuint32_t box[3] = {(cuuint32_t)RW, (cuuint32_t)RH, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(frames_dev), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error_text("cuTensorMapEncodeTiled failed");
    return CTAG_ERR_CUDA;
  }
  static const bool noslide = getenv("CTAG_FRONT_NOSLIDE") != nullptr;  // A/B switch, read once per process
  if (!noslide) {
    CUtensorMap tmap_main, tmap_top;
    cuuint32_t box_main[3] = {(cuuint32_t)RW, (cuuint32_t)SROWS, 1}, box_top[3] = {(cuuint32_t)RW, (cuuint32_t)OVR, 1};
    if (enc(&tmap_main, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(frames_dev), dims, strides, box_main, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&tmap_top, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(frames_dev), dims, strides, box_top, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_last_error_text("cuTensorMapEncodeTiled failed");
      return CTAG_ERR_CUDA;
    }
    return launch_front_slide(tmap_main, tmap_top, n, geo, channels, gray_out, gray_fstride, bin_out, bin_fstride, hint, stream,
                              ev_start);
  }
  return channels == 3 ? launch_front_t<3>(tmap, n, geo, gray_out, gray_fstride, bin_out, bin_fstride, hint, stream, ev_start)
                       : launch_front_t<1>(tmap, n, geo, gray_out, gray_fstride, bin_out, bin_fstride, hint, stream, ev_start);
}

}  // namespace ctag
