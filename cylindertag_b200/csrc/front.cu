// K1: fused dense front end  (reference rows a0-a3 of SURVEY 8a)
//
//   a0  BGR -> gray                 main.cpp:36,54          (cvtColor, integer weights 3735/19235/9798 >> 15)
//   a1  2x bicubic decimation       CylinderTag.cpp:79      (separable taps (-3,19,19,-3)/32, RNE, saturate)
//   a2  u8 -> f32 * 1/255           CylinderTag.cpp:80      (never materialised: applied to tile extrema only)
//   a3  adaptiveThreshold           corner_detector.cpp:28-79 (5x5 tile min/max, 3x3 tile dilation, threshold)
//
// One CTA produces a 80x40 half-resolution patch (16x8 threshold tiles).  It needs the 18x10 surrounding tiles,
// i.e. a 90x50 half-res patch, i.e. a 192x102 full-res region which is staged in shared memory by TMA
// (cp.async.bulk.tensor.3d, zero fill outside the image; replicate borders are patched afterwards).
// The half-res gray image, the float image and the tile extrema never touch HBM.
//
// HBM traffic per frame (N = w*h): BGR input 3N read + N gray write + N/4 binary write; gray input N + N/4.
#include "common.cuh"
#include "kernels.cuh"

namespace ctag {

namespace front {
constexpr int OW = 80, OH = 40;      // owned half-res pixels per CTA
constexpr int OTX = 16, OTY = 8;     // owned tiles
constexpr int CTX = 18, CTY = 10;    // computed tiles (owned + 1 ring)
constexpr int RW = 192, RH = 102;    // full-res region (pixels) staged per CTA
constexpr int HP = 92;               // pitch of the horizontal-pass buffer (int16 elements)
constexpr int PP = 112;              // pitch of the half-res patch (bytes)
constexpr int POFF = 11;             // column shift of the half-res patch so that owned pixels start 16B aligned
constexpr int NT = 256;
constexpr int BOX = RW * RH;         // bytes of one TMA box
constexpr int H_BYTES = ((RH * HP * 2 + 127) / 128) * 128;
constexpr int P_BYTES = ((50 * PP + 127) / 128) * 128;

template <int C>
struct Layout {
  // C==3: [bgr 3 boxes | g | small]; H and P alias the bgr boxes once the gray conversion is done.
  // C==1: [g | H | P | small]
  static constexpr int bgr = 0;
  static constexpr int g = (C == 3) ? 3 * BOX : 0;
  static constexpr int h = (C == 3) ? 0 : BOX;
  static constexpr int p = (C == 3) ? H_BYTES : BOX + H_BYTES;
  static constexpr int small_ = (C == 3) ? 4 * BOX : BOX + H_BYTES + P_BYTES;
  static constexpr int tmin = small_;             // CTY*CTX bytes
  static constexpr int tmax = small_ + 192;       // CTY*CTX bytes
  static constexpr int vthr = small_ + 384;       // OTY*OTX int16
  static constexpr int mbar = small_ + 384 + 256; // 8 bytes
  static constexpr int total = small_ + 384 + 256 + 16;
};
}  // namespace front

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// cvtColor(BGR2GRAY): (3735*B + 19235*G + 9798*R + 16384) >> 15, coefficients split into hi/lo bytes for dp4a.
__device__ __forceinline__ uint32_t gray_of(uint32_t bgrx) {
  const uint32_t LO = 151u | (35u << 8) | (70u << 16);
  const uint32_t HI = 14u | (75u << 8) | (38u << 16);
  uint32_t lo = dp4a_uu(bgrx, LO, 16384u);
  uint32_t hi = dp4a_uu(bgrx, HI, 0u);
  return (lo + (hi << 8)) >> 15;
}

__device__ __forceinline__ uint32_t gray4(uint32_t w0, uint32_t w1, uint32_t w2) {
  uint32_t p0 = w0;
  uint32_t p1 = __byte_perm(w0, w1, 0x0543);
  uint32_t p2 = __byte_perm(w1, w2, 0x0432);
  uint32_t p3 = w2 >> 8;
  uint32_t g0 = gray_of(p0), g1 = gray_of(p1), g2 = gray_of(p2), g3 = gray_of(p3);
  return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
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

template <int C>
__global__ void __launch_bounds__(front::NT) front_kernel(const __grid_constant__ CUtensorMap tmap, FrameGeom geo,
                                                          uint8_t* __restrict__ gray_out, size_t gray_fstride,
                                                          uint8_t* __restrict__ bin_out, size_t bin_fstride) {
  using namespace front;
  using L = Layout<C>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int cx = blockIdx.x, cy = blockIdx.y, fr = blockIdx.z;
  const int x0r = 2 * OW * cx - 16;  // full-res x of region column 0
  const int y0r = 2 * OH * cy - 11;  // full-res y of region row 0
  uint8_t* g = smem + L::g;
  const uint32_t mbar = smem_u32(smem + L::mbar);

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(C * BOX) : "memory");
#pragma unroll
    for (int b = 0; b < C; ++b) {
      uint32_t dst = smem_u32(smem + (C == 3 ? L::bgr + b * BOX : L::g));
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
          :
          : "r"(dst), "l"(&tmap), "r"(mbar), "r"(C * x0r + RW * b), "r"(y0r), "r"(fr)
          : "memory");
    }
  }
  mbar_wait(mbar, 0);

  // ---- phase A: BGR -> gray (whole region) + store of the owned gray pixels --------------------------------
  if (C == 3) {
    uint8_t* gray_f = gray_out + (size_t)fr * gray_fstride;
    for (int item = tid; item < RH * 12; item += NT) {
      int row = item / 12, gq = item - row * 12;
      const uint4* src = reinterpret_cast<const uint4*>(smem + L::bgr + (gq >> 2) * BOX + row * RW + (gq & 3) * 48);
      uint4 a = src[0], b = src[1], c = src[2];
      uint4 o;
      o.x = gray4(a.x, a.y, a.z);
      o.y = gray4(a.w, b.x, b.y);
      o.z = gray4(b.z, b.w, c.x);
      o.w = gray4(c.y, c.z, c.w);
      *reinterpret_cast<uint4*>(g + row * RW + gq * 16) = o;
      int y = y0r + row, x = x0r + gq * 16;
      if (row >= 11 && row < 11 + 2 * OH && gq >= 1 && gq <= 10 && y < geo.h && x < geo.w)
        *reinterpret_cast<uint4*>(gray_f + (size_t)y * geo.gpitch + x) = o;
    }
    __syncthreads();
  }

  // ---- replicate-border patch (TMA zero-fills outside the image; INTER_CUBIC uses BORDER_REPLICATE) ---------
  {
    const int rW = geo.w - x0r, rH = geo.h - y0r;
    const bool left = (cx == 0), right = (rW < RW), top = (cy == 0), bottom = (rH < RH);
    if (left | right | top | bottom) {
      if (left)
        for (int r = tid; r < RH; r += NT) g[r * RW + 15] = g[r * RW + 16];
      if (right)
        for (int r = tid; r < RH; r += NT) g[r * RW + rW] = g[r * RW + rW - 1];
      __syncthreads();
      if (top)
        for (int c = tid; c < RW; c += NT) g[10 * RW + c] = g[11 * RW + c];
      if (bottom)
        for (int c = tid; c < RW; c += NT) g[rH * RW + c] = g[(rH - 1) * RW + c];
      __syncthreads();
    }
  }

  // ---- phase B: horizontal taps (-3,19,19,-3): H[row][j] from region columns 2j+5..2j+8 ---------------------
  int16_t* H = reinterpret_cast<int16_t*>(smem + L::h);
  {
    const uint32_t COEF = 0xFD1313FDu;  // (-3, 19, 19, -3) as signed bytes
    for (int item = tid; item < RH * 23; item += NT) {
      int row = item / 23, k = item - row * 23;
      const uint8_t* src = g + row * RW + 8 * k;
      uint32_t w1 = *reinterpret_cast<const uint32_t*>(src + 4);
      uint2 w23 = *reinterpret_cast<const uint2*>(src + 8);
      int h0 = dp4a_us(__byte_perm(w1, w23.x, 0x4321), COEF, 0);
      int h1 = dp4a_us(__byte_perm(w1, w23.x, 0x6543), COEF, 0);
      int h2 = dp4a_us(__byte_perm(w23.x, w23.y, 0x4321), COEF, 0);
      int h3 = dp4a_us(__byte_perm(w23.x, w23.y, 0x6543), COEF, 0);
      uint2 o;
      o.x = (uint32_t)(h0 & 0xFFFF) | ((uint32_t)h1 << 16);
      o.y = (uint32_t)(h2 & 0xFFFF) | ((uint32_t)h3 << 16);
      *reinterpret_cast<uint2*>(H + row * HP + 4 * k) = o;
    }
  }
  __syncthreads();

  // ---- phase C: vertical taps + round-half-even + saturate: P[i][j] from H rows 2i..2i+3 --------------------
  uint8_t* P = smem + L::p;
  for (int item = tid; item < 50 * 45; item += NT) {
    int i = item / 45, jp = item - i * 45;
    const int16_t* hp = H + (2 * i) * HP + 2 * jp;
    uint32_t r0 = *reinterpret_cast<const uint32_t*>(hp);
    uint32_t r1 = *reinterpret_cast<const uint32_t*>(hp + HP);
    uint32_t r2 = *reinterpret_cast<const uint32_t*>(hp + 2 * HP);
    uint32_t r3 = *reinterpret_cast<const uint32_t*>(hp + 3 * HP);
    int va = 19 * ((int)(int16_t)r1 + (int)(int16_t)r2) - 3 * ((int)(int16_t)r0 + (int)(int16_t)r3);
    int vb = 19 * (((int)r1 >> 16) + ((int)r2 >> 16)) - 3 * (((int)r0 >> 16) + ((int)r3 >> 16));
    va = (va + 511 + ((va >> 10) & 1)) >> 10;
    vb = (vb + 511 + ((vb >> 10) & 1)) >> 10;
    va = min(max(va, 0), 255);
    vb = min(max(vb, 0), 255);
    P[i * PP + POFF + 2 * jp] = (uint8_t)va;  // POFF is odd: two byte stores
    P[i * PP + POFF + 2 * jp + 1] = (uint8_t)vb;
  }
  __syncthreads();

  // ---- phase D: 5x5 tile min/max over valid pixels (corner_detector.cpp:42-53) ------------------------------
  uint8_t* tmin = smem + L::tmin;
  uint8_t* tmax = smem + L::tmax;
  if (tid < CTX * CTY) {
    int ti = tid / CTX, tj = tid - ti * CTX;
    int mn = 255, mx = 0;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
      int i = 5 * ti + dy;
      int yh = OH * cy - 5 + i;
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) {
        int j = 5 * tj + dx;
        int xh = OW * cx - 5 + j;
        if (yh >= 0 && yh < geo.hh && xh >= 0 && xh < geo.hw) {
          int v = P[i * PP + POFF + j];
          mn = min(mn, v);
          mx = max(mx, v);
        }
      }
    }
    tmin[tid] = (uint8_t)mn;
    tmax[tid] = (uint8_t)mx;
  }
  __syncthreads();

  // ---- phase E: 3x3 tile dilation -> integer threshold per owned tile (corner_detector.cpp:54-78) -----------
  int16_t* vthr = reinterpret_cast<int16_t*>(smem + L::vthr);
  if (tid < OTX * OTY) {
    int oi = tid / OTX, oj = tid - oi * OTX;
    int ty = OTY * cy + oi, tx = OTX * cx + oj;
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
      // dst = 255 iff src < min(0.3f, (max+min)/2) in float; src = lut255(v) is strictly increasing in v,
      // so find the smallest v with lut255(v) >= thr and compare integers per pixel.
      float thr = fminf(0.3f, __fmul_rn(__fadd_rn(lut255(mx), lut255(mn)), 0.5f));
      t = min(max((int)(thr * 255.0f), 0), 255);
      while (t > 0 && !(lut255(t - 1) < thr)) --t;
      while (t < 256 && lut255(t) < thr) ++t;
    }
    vthr[tid] = (int16_t)t;
  }
  __syncthreads();

  // ---- phase F: threshold the owned 80x40 pixels, 16 per thread, 128-bit stores -----------------------------
  {
    uint8_t* bin_f = bin_out + (size_t)fr * bin_fstride;
    for (int item = tid; item < OH * 5; item += NT) {
      int i = item / 5, q = item - i * 5;
      int yh = OH * cy + i, xh0 = OW * cx + 16 * q;
      if (yh >= geo.hh || xh0 >= geo.bpitch) continue;
      uint4 v = *reinterpret_cast<const uint4*>(P + (i + 5) * PP + 16 + 16 * q);
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t o[4];
      const int16_t* trow = vthr + (i / 5) * OTX;
#pragma unroll
      for (int ww = 0; ww < 4; ++ww) {
        uint32_t acc = 0;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          int k = 4 * ww + bb;
          int jo = 16 * q + k;
          int val = (w[ww] >> (8 * bb)) & 255;
          bool fg = (val < (int)trow[jo / 5]) && (xh0 + k < geo.hw);
          acc |= (fg ? 255u : 0u) << (8 * bb);
        }
        o[ww] = acc;
      }
      *reinterpret_cast<uint4*>(bin_f + (size_t)yh * geo.bpitch + xh0) = make_uint4(o[0], o[1], o[2], o[3]);
    }
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

int launch_front(const void* frames_dev, int n, const FrameGeom& geo, int channels, size_t pitch, size_t frame_stride,
                 uint8_t* gray_out, size_t gray_fstride, uint8_t* bin_out, size_t bin_fstride, cudaStream_t stream) {
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
  dim3 grid((geo.hw + OW - 1) / OW, (geo.hh + OH - 1) / OH, n);
  if (channels == 3) {
    CTAG_CUDA_CHECK(cudaFuncSetAttribute(front_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Layout<3>::total));
    front_kernel<3><<<grid, NT, Layout<3>::total, stream>>>(tmap, geo, gray_out, gray_fstride, bin_out, bin_fstride);
  } else {
    CTAG_CUDA_CHECK(cudaFuncSetAttribute(front_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Layout<1>::total));
    front_kernel<1><<<grid, NT, Layout<1>::total, stream>>>(tmap, geo, gray_out, gray_fstride, bin_out, bin_fstride);
  }
  CTAG_CUDA_CHECK(cudaGetLastError());
  return CTAG_OK;
}

}  // namespace ctag
