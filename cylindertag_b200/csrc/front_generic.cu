// Generic-window front end: the same rows a0-a3 as front.cu for an arbitrary adaptiveThresh window W (reference
// default and every shipped config use W = 5, which front.cu fuses into one TMA kernel).  Three plain kernels with the
// half-resolution image and the per-tile extrema in HBM; meant for completeness of the detect() signature
// (CylinderTag.cpp:67, corner_detector.cpp:28-79), not for speed.
#include "common.cuh"
#include "kernels.cuh"

namespace ctag {

__device__ __forceinline__ int gen_gray_at(const uint8_t* __restrict__ f, size_t pitch, int channels, int w, int h, int x, int y) {
  x = min(max(x, 0), w - 1);  // INTER_CUBIC replicates the border
  y = min(max(y, 0), h - 1);
  const uint8_t* p = f + (size_t)y * pitch + (size_t)x * channels;
  if (channels == 1) return p[0];
  return (3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15;  // cvtColor(BGR2GRAY), main.cpp:54
}

// one thread per half-res pixel: 4x4 taps (-3,19,19,-3)/32 per axis, one round-half-even, saturate;
// BGR input additionally writes the 2x2 full-res gray pixels it owns
__global__ void __launch_bounds__(256) gen_decimate_kernel(const uint8_t* __restrict__ frames, size_t pitch, size_t fstride,
                                                           int channels, FrameGeom g, uint8_t* __restrict__ gray_out,
                                                           size_t gray_fstride, uint8_t* __restrict__ half_out,
                                                           size_t half_fstride) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), fr = blockIdx.z;
  if (x >= g.hw || y >= g.hh) return;
  const uint8_t* f = frames + (size_t)fr * fstride;
  const int c[4] = {-3, 19, 19, -3};
  int v = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int hsum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) hsum += c[k] * gen_gray_at(f, pitch, channels, g.w, g.h, 2 * x - 1 + k, 2 * y - 1 + r);
    v += c[r] * hsum;
  }
  v = (v + 511 + ((v >> 10) & 1)) >> 10;
  half_out[(size_t)fr * half_fstride + (size_t)y * g.bpitch + x] = (uint8_t)min(max(v, 0), 255);
  if (channels == 3) {
    uint8_t* go = gray_out + (size_t)fr * gray_fstride;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx)
        go[(size_t)(2 * y + dy) * g.gpitch + 2 * x + dx] = (uint8_t)gen_gray_at(f, pitch, channels, g.w, g.h, 2 * x + dx, 2 * y + dy);
  }
}

// one thread per tile: min/max over the (clipped) W x W window (corner_detector.cpp:42-53)
__global__ void __launch_bounds__(128) gen_tile_kernel(const uint8_t* __restrict__ half, size_t half_fstride, FrameGeom g,
                                                       int W, int cn, int rn, uint8_t* __restrict__ tmin,
                                                       uint8_t* __restrict__ tmax) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, fr = blockIdx.y;
  if (t >= cn * rn) return;
  const int ti = t / cn, tj = t - ti * cn;
  const uint8_t* hp = half + (size_t)fr * half_fstride;
  int mn = 255, mx = 0;
  for (int y = ti * W; y < min(ti * W + W, g.hh); ++y)
    for (int x = tj * W; x < min(tj * W + W, g.hw); ++x) {
      int v = hp[(size_t)y * g.bpitch + x];
      mn = min(mn, v);
      mx = max(mx, v);
    }
  tmin[(size_t)fr * cn * rn + t] = (uint8_t)mn;
  tmax[(size_t)fr * cn * rn + t] = (uint8_t)mx;
}

// one thread per half-res pixel: 3x3 tile dilation, threshold in float exactly like the reference (:54-78);
// the border ring of tiles is background (SURVEY C-1)
__global__ void __launch_bounds__(256) gen_threshold_kernel(const uint8_t* __restrict__ half, size_t half_fstride, FrameGeom g,
                                                            int W, int cn, int rn, const uint8_t* __restrict__ tmin,
                                                            const uint8_t* __restrict__ tmax, uint8_t* __restrict__ bin,
                                                            size_t bin_fstride) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), fr = blockIdx.z;
  if (x >= g.bpitch || y >= g.hh) return;
  uint8_t out = 0;
  if (x < g.hw) {
    const int ti = y / W, tj = x / W;
    if (ti >= 1 && ti <= rn - 2 && tj >= 1 && tj <= cn - 2) {
      const uint8_t* a = tmin + (size_t)fr * cn * rn;
      const uint8_t* b = tmax + (size_t)fr * cn * rn;
      int mn = 255, mx = 0;
      for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
          mn = min(mn, (int)a[(ti + di) * cn + tj + dj]);
          mx = max(mx, (int)b[(ti + di) * cn + tj + dj]);
        }
      const float k = (float)(1.0 / 255);
      const float thr = fminf(0.3f, __fmul_rn(__fadd_rn(__fmul_rn((float)mx, k), __fmul_rn((float)mn, k)), 0.5f));
      const float v = __fmul_rn((float)half[(size_t)fr * half_fstride + (size_t)y * g.bpitch + x], k);
      out = v < thr ? 255 : 0;
    }
  }
  bin[(size_t)fr * bin_fstride + (size_t)y * g.bpitch + x] = out;
}

int launch_front_generic(const void* frames_dev, int n, const FrameGeom& g, int channels, size_t pitch, size_t frame_stride,
                         int window, uint8_t* gray_out, size_t gray_fstride, uint8_t* half_buf, uint8_t* tile_buf,
                         uint8_t* bin_out, size_t bin_fstride, cudaStream_t stream, int* launches) {
  const int cn = (g.hw + window - 1) / window, rn = (g.hh + window - 1) / window;
  const size_t half_fstride = bin_fstride;  // same geometry as the binary image
  dim3 grid((g.bpitch + 31) / 32, (g.hh + 7) / 8, n);
  gen_decimate_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(frames_dev), pitch, frame_stride, channels, g,
                                                gray_out, gray_fstride, half_buf, half_fstride);
  uint8_t* tmin = tile_buf;
  uint8_t* tmax = tile_buf + (size_t)n * cn * rn;
  gen_tile_kernel<<<dim3((cn * rn + 127) / 128, n), 128, 0, stream>>>(half_buf, half_fstride, g, window, cn, rn, tmin, tmax);
  gen_threshold_kernel<<<grid, 256, 0, stream>>>(half_buf, half_fstride, g, window, cn, rn, tmin, tmax, bin_out, bin_fstride);
  CTAG_CUDA_CHECK(cudaGetLastError());
  if (launches) *launches += 3;
  return CTAG_OK;
}

}  // namespace ctag
