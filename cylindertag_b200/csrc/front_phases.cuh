// Stencil / threshold phases of the fused dense front end (front.cu), shared by the gray-input and the BGR-input kernel.
// Every function is called by all threads of the CTA (NT threads); the caller places the CTA barriers between them
// unless noted.  Buffers (shared memory):
//   g     gray tile, RH rows of RW bytes (full-res region of the tile, replicate borders patched in)
//   HT    horizontal-pass buffer, RH/2 rows of HP words: HT[rp][j'] = (h[2rp][j'], h[2rp+1][j']) as an int16 pair
//   P     half-res patch, 50 rows of PP bytes (column of half-res x = 80cx - 5 + j is byte POFF + j)
//   cmn / cmx    column extrema per threshold-tile row, CTY rows of 96 bytes (column index = j + HOFF)
//   tmin / tmax  5x5 tile extrema, CTY x CTX bytes
//   thr16        integer threshold per owned pixel column, OTY rows of OW bytes
#pragma once
#include "common.cuh"

namespace ctag {

namespace front {
constexpr int OW = 80, OH = 40;      // owned half-res pixels per tile
constexpr int OTX = 16, OTY = 8;     // owned threshold tiles
constexpr int CTX = 18, CTY = 10;    // computed threshold tiles (owned + 1 ring)
constexpr int RW = 192, RH = 102;    // full-res region (pixels) per tile
constexpr int HP = 96;               // pitch of the horizontal-pass buffer (row-pair words), 24 quads of columns
constexpr int PP = 112;              // pitch of the half-res patch (bytes)
constexpr int POFF = 11;             // patch column of half-res column j (x = 80cx - 5 + j); owned pixels start at 16
constexpr int HOFF = 3;              // the stencil passes run on j' = j + HOFF so that their 4-column groups are aligned
constexpr int NT = 384;              // 12 warps per CTA, three CTAs per SM (shared memory bound)
constexpr int BOX = RW * RH;         // bytes of one gray tile
constexpr int H_BYTES = (((RH / 2) * HP * 4 + 127) / 128) * 128;
constexpr int P_BYTES = ((50 * PP + 127) / 128) * 128;
// offsets of the small arrays behind the patch (same in both kernels' layouts)
constexpr int S_TMIN = 0, S_TMAX = 192, S_CMN = 384, S_CMX = 384 + 960, S_THR16 = 2304 + 128, S_MBAR = 2304 + 128 + 640;
constexpr int S_BYTES = 2304 + 128 + 640 + 16;
}  // namespace front

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// dp2a: d = c + a.u16[0] * b.u8[2h] + a.u16[1] * b.u8[2h+1]  (h = 0 for .lo, 1 for .hi)
__device__ __forceinline__ uint32_t dp2a_lo_uu(uint32_t a16, uint32_t b8, uint32_t c) {
  uint32_t d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16), "r"(b8), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t dp2a_hi_uu(uint32_t a16, uint32_t b8, uint32_t c) {
  uint32_t d;
  asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16), "r"(b8), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_lo_ss(uint32_t a16, uint32_t b8, int c) {
  int d;
  asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16), "r"(b8), "r"(c));
  return d;
}

// d = (c << 16) | (sat_u8(a) << 8) | sat_u8(b)
__device__ __forceinline__ uint32_t pack_sat_u8(int a, int b, uint32_t c) {
  uint32_t d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// cvtColor(BGR2GRAY): (3735*B + 19235*G + 9798*R + 16384) >> 15.  With the coefficients doubled the result is byte 2
// of (7470*B + 38470*G + 19596*R + 32768) < 2^24, and 16-bit coefficients fit dp2a: two IDP per pixel, no byte
// extraction (the four pixels of a 12-byte group sit at byte offsets 0,3,6,9; each lands on one .lo and one .hi).
__device__ __forceinline__ uint32_t gray4(uint32_t w0, uint32_t w1, uint32_t w2) {
  const uint32_t cB = 7470u, cG = 38470u, cR = 19596u;
  const uint32_t BG = cB | (cG << 16), R0 = cR, ZB = cB << 16, GR = cG | (cR << 16);
  uint32_t g0 = dp2a_hi_uu(R0, w0, dp2a_lo_uu(BG, w0, 32768u));  // B,G,R = w0.b0 w0.b1 w0.b2
  uint32_t g1 = dp2a_lo_uu(GR, w1, dp2a_hi_uu(ZB, w0, 32768u));  // w0.b3 w1.b0 w1.b1
  uint32_t g2 = dp2a_lo_uu(R0, w2, dp2a_hi_uu(BG, w1, 32768u));  // w1.b2 w1.b3 w2.b0
  uint32_t g3 = dp2a_hi_uu(GR, w2, dp2a_lo_uu(ZB, w2, 32768u));  // w2.b1 w2.b2 w2.b3
  return __byte_perm(__byte_perm(g0, g1, 0x0062), __byte_perm(g2, g3, 0x0062), 0x5410);
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
  } while (!ok);
}

// float image value of a u8 sample: Mat::convertTo(CV_32F, 1.0/255) = float(v) * float(1.0/255)  
__device__ __forceinline__ float lut255(int v) { return __fmul_rn((float)v, (float)(1.0 / 255)); }

// ---- replicate-border patch (TMA zero-fills outside the image; INTER_CUBIC uses BORDER_REPLICATE).  Contains its own
//      CTA barriers (taken by all threads or none: the conditions are uniform). ------------------------------------------
__device__ __forceinline__ void phase_border(uint8_t* g, const FrameGeom& geo, int cx, int cy, int x0r, int y0r, int tid,
                                             int r0 = 0) {  // rows below r0 are already patched (sliding kernels)
  using namespace front;
  const int rW = geo.w - x0r, rH = geo.h - y0r;
  const bool left = (cx == 0), right = (rW < RW), top = (cy == 0), bottom = (rH < RH);
  if (left | right | top | bottom) {
    if (left)
      for (int r = r0 + tid; r < RH; r += NT) g[r * RW + 15] = g[r * RW + 16];
    if (right)
      for (int r = r0 + tid; r < RH; r += NT) g[r * RW + rW] = g[r * RW + rW - 1];
    __syncthreads();
    if (top)
      for (int c = tid; c < RW; c += NT) g[10 * RW + c] = g[11 * RW + c];
    if (bottom)
      for (int c = tid; c < RW; c += NT) g[rH * RW + c] = g[(rH - 1) * RW + c];
    __syncthreads();
  }
}

// ---- phase B: horizontal taps (-3,19,19,-3) on two rows at a time: HT[rp][j'] = (h[2rp][j'], h[2rp+1][j']) as an int16
//      pair, the layout the vertical dp2a wants.  j' = j + 3 (x = 80cx - 8 + j'), so h[row][j'] uses region columns
//      2j'-1..2j'+2 and every group of four j' starts on an 8-byte boundary of the gray row; columns j' = 0..2 and 93..95
//      are never used. ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void phase_horizontal(const uint8_t* g, uint32_t* HT, int tid, int rp0 = 0) {  // row pairs >= rp0
  using namespace front;
  const uint32_t COEF = 0xFD1313FDu;  // (-3, 19, 19, -3) as signed bytes
  const int k = tid % 24;
  int rp = rp0 + tid / 24;  // row pairs rp, rp+16, ...
  const uint8_t* src = g + (2 * rp) * RW + 8 * k;
  uint32_t* dst = HT + rp * HP + 4 * k;
  // the word in front of k == 0 and the word behind k == 23 only feed the unused columns j' = 0 and 95: those two
  // loads are redirected to a word of the thread's own span so that nothing outside the row (and outside the buffer)
  // is read, without a branch
  const uint8_t* s0 = src + (k ? -4 : 0);
  const uint8_t* s3 = src + (k < 23 ? 8 : 4);
  for (; rp < RH / 2; rp += NT / 24, src += 2 * (NT / 24) * RW, s0 += 2 * (NT / 24) * RW, s3 += 2 * (NT / 24) * RW,
                      dst += (NT / 24) * HP) {
    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(s0), a3 = *reinterpret_cast<const uint32_t*>(s3);
    const uint2 a12 = *reinterpret_cast<const uint2*>(src);
    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(s0 + RW), b3 = *reinterpret_cast<const uint32_t*>(s3 + RW);
    const uint2 b12 = *reinterpret_cast<const uint2*>(src + RW);
    const int h0 = dp4a_us(__byte_perm(a0, a12.x, 0x6543), COEF, 0), g0 = dp4a_us(__byte_perm(b0, b12.x, 0x6543), COEF, 0);
    const int h1 = dp4a_us(__byte_perm(a12.x, a12.y, 0x4321), COEF, 0), g1 = dp4a_us(__byte_perm(b12.x, b12.y, 0x4321), COEF, 0);
    const int h2 = dp4a_us(__byte_perm(a12.x, a12.y, 0x6543), COEF, 0), g2 = dp4a_us(__byte_perm(b12.x, b12.y, 0x6543), COEF, 0);
    const int h3 = dp4a_us(__byte_perm(a12.y, a3, 0x4321), COEF, 0), g3 = dp4a_us(__byte_perm(b12.y, b3, 0x4321), COEF, 0);
    uint4 o;
    o.x = __byte_perm((uint32_t)h0, (uint32_t)g0, 0x5410);
    o.y = __byte_perm((uint32_t)h1, (uint32_t)g1, 0x5410);
    o.z = __byte_perm((uint32_t)h2, (uint32_t)g2, 0x5410);
    o.w = __byte_perm((uint32_t)h3, (uint32_t)g3, 0x5410);
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

// ---- phase C+D1: vertical taps + round-half-even + saturate: P[i][8 + j'] from row pairs i and i+1, and in the same
//      registers the column extrema of the five rows of a threshold-tile row (first half of the 5x5 tile min/max,
//      corner_detector.cpp:42-53).  One thread per (tile row ti >= ti0, word of four patch columns): it walks its five
//      half-res rows top down, so every HT row is loaded once and reused for the next output row. ----------------------
//      Sliding kernels: the last row pair a tile reads (50) is the first one the tile below reads (10); it is handed
//      over through `keep_out` / `keep_in` (HP words each, two alternating buffers) because HT itself does not survive. --
__device__ __forceinline__ void phase_vertical_extrema(const uint32_t* HT, uint8_t* P, uint8_t* cmn, uint8_t* cmx,
                                                       const FrameGeom& geo, int cx, int cy, bool edge_cta, int ti0, int tid,
                                                       const uint32_t* keep_in = nullptr, uint32_t* keep_out = nullptr) {
  using namespace front;
  if (tid >= 24 * (CTY - ti0)) return;
  const uint32_t C01 = 0x000013FDu;  // (-3, 19) on bytes 0,1
  const uint32_t C23 = 0x0000FD13u;  // (19, -3)
  const int k = tid % 24, ti = ti0 + tid / 24;
  const uint32_t* src = HT + (5 * ti) * HP + 4 * k;
  uint8_t* dst = P + (5 * ti) * PP + (POFF - HOFF) + 4 * k;  // 4-byte aligned; patch columns 8 + 4k .. (j = 4k - 3 ..)
  uint32_t keepx = 0xFFFFFFFFu;
  if (edge_cta) {
    // pixels outside the image do not take part in the extrema (corner_detector.cpp:44 clips the window)
    keepx = 0u;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int xh = OW * cx - 8 + 4 * k + bb;
      if (xh >= 0 && xh < geo.hw) keepx |= 0xFFu << (8 * bb);
    }
  }
  uint4 a = *reinterpret_cast<const uint4*>((keep_in && ti == ti0) ? keep_in + 4 * k : src);
  // byte-wise min/max through the native 16x2 three-input min/max on the even and odd bytes (the 8x4 video
  // intrinsics are emulated on sm_100)
  uint32_t ve[5], vo[5], xe[5], xo[5];
#pragma unroll
  for (int dy = 0; dy < 5; ++dy) {
    const uint4 b = *reinterpret_cast<const uint4*>(src + (dy + 1) * HP);
    int v0 = dp2a_lo_ss(b.x, C23, dp2a_lo_ss(a.x, C01, 0));
    int v1 = dp2a_lo_ss(b.y, C23, dp2a_lo_ss(a.y, C01, 0));
    int v2 = dp2a_lo_ss(b.z, C23, dp2a_lo_ss(a.z, C01, 0));
    int v3 = dp2a_lo_ss(b.w, C23, dp2a_lo_ss(a.w, C01, 0));
    // v / 1024 rounded half to even (what cv::resize's float path does), then saturate_cast<uchar>
    v0 = (v0 + 511 + ((v0 >> 10) & 1)) >> 10;
    v1 = (v1 + 511 + ((v1 >> 10) & 1)) >> 10;
    v2 = (v2 + 511 + ((v2 >> 10) & 1)) >> 10;
    v3 = (v3 + 511 + ((v3 >> 10) & 1)) >> 10;
    const uint32_t w = pack_sat_u8(v1, v0, pack_sat_u8(v3, v2, 0u));
    *reinterpret_cast<uint32_t*>(dst + dy * PP) = w;
    uint32_t lo = w, hi = w;
    if (edge_cta) {
      const int yh = OH * cy - 5 + 5 * ti + dy;
      const uint32_t keep = (yh >= 0 && yh < geo.hh) ? keepx : 0u;
      lo = w | ~keep;
      hi = w & keep;
    }
    ve[dy] = lo & 0x00FF00FFu;
    vo[dy] = __byte_perm(lo, 0u, 0x4341);
    xe[dy] = hi & 0x00FF00FFu;
    xo[dy] = __byte_perm(hi, 0u, 0x4341);
    a = b;
  }
  if (keep_out && ti == CTY - 1) *reinterpret_cast<uint4*>(keep_out + 4 * k) = a;  // row pair 50
  const uint32_t mne = __vimin3_u16x2(__vimin3_u16x2(ve[0], ve[1], ve[2]), ve[3], ve[4]);
  const uint32_t mno = __vimin3_u16x2(__vimin3_u16x2(vo[0], vo[1], vo[2]), vo[3], vo[4]);
  const uint32_t mxe = __vimax3_u16x2(__vimax3_u16x2(xe[0], xe[1], xe[2]), xe[3], xe[4]);
  const uint32_t mxo = __vimax3_u16x2(__vimax3_u16x2(xo[0], xo[1], xo[2]), xo[3], xo[4]);
  *reinterpret_cast<uint32_t*>(cmn + ti * 96 + 4 * k) = __byte_perm(mne, mno, 0x6240);  // cmn/cmx column index = j + 3
  *reinterpret_cast<uint32_t*>(cmx + ti * 96 + 4 * k) = __byte_perm(mxe, mxo, 0x6240);
}

// ---- phase D2: 5x5 tile min/max from five column extrema --------------------------------------------------------------
__device__ __forceinline__ void phase_tile_extrema(const uint8_t* cmn, const uint8_t* cmx, uint8_t* tmin, uint8_t* tmax, int tid) {
  using namespace front;
  if (tid >= CTX * CTY) return;
  const int ti = tid / CTX, tj = tid - ti * CTX;
  int mn = 255, mx = 0;
#pragma unroll
  for (int dx = 0; dx < 5; ++dx) {
    mn = min(mn, (int)cmn[ti * 96 + HOFF + 5 * tj + dx]);
    mx = max(mx, (int)cmx[ti * 96 + HOFF + 5 * tj + dx]);
  }
  tmin[tid] = (uint8_t)mn;
  tmax[tid] = (uint8_t)mx;
}

// ---- phase E: 3x3 tile dilation -> integer threshold per owned tile (corner_detector.cpp:54-78) -----------------------
__device__ __forceinline__ void phase_threshold(const uint8_t* tmin, const uint8_t* tmax, uint8_t* thr16, const FrameGeom& geo,
                                                int cx, int cy, int tid) {
  using namespace front;
  if (tid >= OTX * OTY) return;
  const int oi = tid / OTX, oj = tid - oi * OTX;
  const int ty = OTY * cy + oi, tx = OTX * cx + oj;
  int t = 0;  // border ring / outside: threshold 0 -> background (SURVEY C-1)
  if (tx >= 1 && tx <= geo.cn - 2 && ty >= 1 && ty <= geo.rn - 2) {
    int mn = 255, mx = 0;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        mn = min(mn, (int)tmin[(oi + dy) * CTX + oj + dx]);
        mx = max(mx, (int)tmax[(oi + dy) * CTX + oj + dx]);
      }
    // dst = 255 iff src < min(0.3f, (max+min)/2) in float; src = lut255(v) is strictly increasing in v, so find
    // the smallest v with lut255(v) >= thr and compare integers per pixel (t <= 77 because thr <= 0.3).
    const float thr = fminf(0.3f, __fmul_rn(__fadd_rn(lut255(mx), lut255(mn)), 0.5f));
    t = min(max((int)(thr * 255.0f), 0), 255);
    while (t > 0 && !(lut255(t - 1) < thr)) --t;
    while (t < 256 && lut255(t) < thr) ++t;
  }
  // one threshold byte per owned pixel column of this tile: lets phase F compare 4 pixels per instruction group
  uint8_t* dst = thr16 + oi * OW + 5 * oj;
#pragma unroll
  for (int q = 0; q < 5; ++q) dst[q] = (uint8_t)t;
}

// ---- phase F: threshold the owned 80x40 pixels, 16 per thread, 128-bit stores -----------------------------------------
__device__ __forceinline__ void phase_compare_store(const uint8_t* P, const uint8_t* thr16, uint8_t* __restrict__ bin_out,
                                                    size_t bin_fstride, const TileHint& hint, const FrameGeom& geo, int fr, int cx,
                                                    int cy, int tid) {
  using namespace front;
  if (tid >= OH * 5) return;
  const int i = tid / 5, q = tid - i * 5;
  const int yh = OH * cy + i, xh0 = OW * cx + 16 * q;
  if (yh < geo.hh && xh0 < geo.bpitch) {
    const uint4 v = *reinterpret_cast<const uint4*>(P + (i + 5) * PP + 16 + 16 * q);
    const uint4 t = *reinterpret_cast<const uint4*>(thr16 + (i / 5) * OW + 16 * q);
    const uint32_t vw[4] = {v.x, v.y, v.z, v.w}, tw[4] = {t.x, t.y, t.z, t.w};
    uint32_t o[4];
#pragma unroll
    for (int ww = 0; ww < 4; ++ww) {
      // per-byte v < t for t <= 127: ((v | 0x80) - t) keeps bit 7 iff (v & 0x7f) >= t; v >= 128 is never below t
      const uint32_t d = (vw[ww] | 0x80808080u) - tw[ww];
      const uint32_t lt = ~(d | vw[ww]) & 0x80808080u;
      o[ww] = (lt >> 7) * 255u;
    }
    if (xh0 + 16 > geo.hw) {  // right image edge inside this group: columns >= hw stay background
#pragma unroll
      for (int ww = 0; ww < 4; ++ww)
#pragma unroll
        for (int bb = 0; bb < 4; ++bb)
          if (xh0 + 4 * ww + bb >= geo.hw) o[ww] &= ~(255u << (8 * bb));
    }
    *reinterpret_cast<uint4*>(bin_out + (size_t)fr * bin_fstride + (size_t)yh * geo.bpitch + xh0) =
        make_uint4(o[0], o[1], o[2], o[3]);
    // the 16 pixels lie in one 64x64 tile; foreground is rare, so this store almost never happens
    if ((o[0] | o[1] | o[2] | o[3]) && hint.any) hint.any[(size_t)fr * hint.fstride + (yh >> 6) * hint.pitch + (xh0 >> 6)] = 1;
  }
}

}  // namespace ctag
