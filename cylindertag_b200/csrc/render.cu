// Synthetic CylinderTag frames rendered on the GPU (the reference's MATLAB generator only draws flat marker
// bitmaps, CylinderTag_generator.m:206-245; BASELINE.json's synthetic configs need whole camera frames).  Same scene
// model as the host renderer cylindertag_b200/synth.py: every marker is its flat texture --
// columns of width W at pitch 1.5 W, two black quadrilaterals per column separated by a white band of height 0.2 L whose
// centre runs from pl * L on the left edge to pr * L on the right edge (the cross-ratio digits of the column's state) --
// wrapped once around a cylinder, seen by a pinhole camera.  Per pixel: 3 x 3 rays, ray / cylinder intersection (near
// root) per marker, far markers first; then a Gaussian blur, additive Gaussian noise, rounding to 8 bits and, for BGR
// output, small per-channel offsets around the luminance.  The frames are written straight into device memory in the
// layout ctag_detect_batch_enqueue takes, so a benchmark over thousands of distinct frames needs no host rendering and no
// PCIe transfer.  (Not bit-compatible with synth.py -- different background texture and noise stream; the golden fixtures
// use synth.py.)
#include "common.cuh"
#include "kernels.cuh"

namespace ctag {

struct RenderMarkerDev {
  float R[9];       // rotation object -> camera
  float o[3];       // camera centre in the object frame: -R^T t
  float radius, W, L, black, white;
  int cols, row_off;  // states of the dictionary row: states[row_off .. row_off + cols)
  int bx0, by0, bx1, by1;  // projected bounding box of the cylinder surface (pixels outside cannot hit it)
};

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float hash_unit(uint32_t x) { return (hash_u32(x) >> 8) * (1.0f / 16777216.0f); }

// band centre of a cross-ratio digit, fraction of L (CylinderTag_generator.m:223-242: root of -p^2 + p + (0.11 - 0.2 cr))
__device__ __forceinline__ float band_centre(int digit) {
  const float cr[4] = {1.47f, 1.54f, 1.61f, 1.68f};
  const float c = cr[digit <= 3 ? digit : 7 - digit];
  const float disc = sqrtf(1.0f + 4.0f * (0.11f - 0.2f * c));
  return digit <= 3 ? (1.0f - disc) * 0.5f : (1.0f + disc) * 0.5f;
}

// frame parameters: [0] fx [1] fy [2] cx [3] cy [4] blur sigma [5] noise sigma [6] seed (bits) [7] background phase seed (bits)
__global__ void __launch_bounds__(256) render_scene_kernel(float* __restrict__ img, int w, int h, const float* __restrict__ fparams,
                                                          const RenderMarkerDev* __restrict__ markers, const int* __restrict__ marker_start,
                                                          const int* __restrict__ states) {
  const int fr = blockIdx.z, x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const float* fp = fparams + 8 * fr;
  const float fx = fp[0], fy = fp[1], cx = fp[2], cy = fp[3];
  const uint32_t bseed = __float_as_uint(fp[7]);
  // smooth background in [140, 220]: a few low-frequency waves with seeded phases and directions
  float bg = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float ang = 6.2831853f * hash_unit(bseed + 11u * k), ph = 6.2831853f * hash_unit(bseed + 11u * k + 1u);
    const float freq = (0.6f + 1.9f * hash_unit(bseed + 11u * k + 2u)) * 6.2831853f / (float)w;
    bg += __sinf(freq * (x * __cosf(ang) + y * __sinf(ang)) + ph);
  }
  float v = 180.f + 40.f * bg * 0.3f;
  v = fminf(fmaxf(v, 140.f), 220.f);
  for (int m = marker_start[fr]; m < marker_start[fr + 1]; ++m) {
    if (x < markers[m].bx0 || x > markers[m].bx1 || y < markers[m].by0 || y > markers[m].by1) continue;
    const RenderMarkerDev mk = markers[m];
    const float margin = 0.15f * mk.L, circ = 1.5f * mk.W * mk.cols;
    float acc = 0.f;
    int hits = 0;
#pragma unroll
    for (int sy = 0; sy < 3; ++sy)
#pragma unroll
      for (int sx = 0; sx < 3; ++sx) {
        const float px = x + (sx + 0.5f) * (1.f / 3.f) - 0.5f, py = y + (sy + 0.5f) * (1.f / 3.f) - 0.5f;
        const float d0 = (px - cx) / fx, d1 = (py - cy) / fy, d2 = 1.f;
        // ray in the object frame: direction R^T d, origin o
        const float dx = mk.R[0] * d0 + mk.R[3] * d1 + mk.R[6] * d2, dy = mk.R[1] * d0 + mk.R[4] * d1 + mk.R[7] * d2,
                    dz = mk.R[2] * d0 + mk.R[5] * d1 + mk.R[8] * d2;
        const float a = dx * dx + dz * dz, b = 2.f * (mk.o[0] * dx + mk.o[2] * dz), c = mk.o[0] * mk.o[0] + mk.o[2] * mk.o[2] - mk.radius * mk.radius;
        const float disc = b * b - 4.f * a * c;
        if (disc <= 0.f || a <= 0.f) continue;
        const float s = (-b - sqrtf(disc)) / (2.f * a);
        if (s <= 0.f) continue;
        const float X = mk.o[0] + s * dx, Y = mk.o[1] + s * dy, Z = mk.o[2] + s * dz;
        const float vv = Y + 0.5f * mk.L;
        if (vv < -margin || vv > mk.L + margin) continue;
        float theta = atan2f(X, -Z);
        if (theta < 0.f) theta += 6.2831853f;
        const float u = theta * (1.f / 6.2831853f) * circ;
        const float pitch = 1.5f * mk.W;
        const float col_f = floorf(u / pitch);
        int col = (int)col_f % mk.cols;
        if (col < 0) col += mk.cols;
        const float xc = u - col_f * pitch;
        const int st = states[mk.row_off + col];
        const float pl = band_centre(st >> 3) * mk.L, pr = band_centre(st & 7) * mk.L;
        const float centre = pl + (pr - pl) * fminf(fmaxf(xc / mk.W, 0.f), 1.f);
        const bool black = xc <= mk.W && vv >= 0.f && vv <= mk.L && (vv < centre - 0.1f * mk.L || vv > centre + 0.1f * mk.L);
        acc += black ? mk.black : mk.white;
        ++hits;
      }
    if (hits) {
      const float frac = hits * (1.f / 9.f);
      v = (acc / hits) * frac + v * (1.f - frac);
    }
  }
  img[((size_t)fr * h + y) * w + x] = v;
}

// Gaussian blur (radius ceil(3 sigma) <= 5, direct 2-D sum) + noise + rounding; gray or BGR output
__global__ void __launch_bounds__(256) render_finish_kernel(const float* __restrict__ img, int w, int h, const float* __restrict__ fparams,
                                                           uint8_t* __restrict__ out, size_t pitch, size_t frame_stride, int channels) {
  const int fr = blockIdx.z, x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const float* fp = fparams + 8 * fr;
  const float sigma = fp[4], nsig = fp[5];
  const uint32_t seed = __float_as_uint(fp[6]);
  const float* src = img + (size_t)fr * h * w;
  float v;
  if (sigma > 0.05f) {
    const int rad = min(5, (int)ceilf(3.f * sigma));
    const float inv = -0.5f / (sigma * sigma);
    float acc = 0.f, wsum = 0.f;
    for (int dy = -rad; dy <= rad; ++dy) {
      const int yy = min(max(y + dy, 0), h - 1);
      const float wy = __expf(inv * dy * dy);
      for (int dx = -rad; dx <= rad; ++dx) {
        const int xx = min(max(x + dx, 0), w - 1);
        const float ww = wy * __expf(inv * dx * dx);
        acc += ww * src[(size_t)yy * w + xx];
        wsum += ww;
      }
    }
    v = acc / wsum;
  } else {
    v = src[(size_t)y * w + x];
  }
  const uint32_t pix = (uint32_t)(y * w + x);
  if (nsig > 0.f) {  // Box-Muller on two hashed uniforms
    const float u1 = fmaxf(hash_unit(seed ^ (pix * 2u + 1u)), 1e-7f), u2 = hash_unit(seed + 0x9E3779B9u + pix * 2u);
    v += nsig * sqrtf(-2.f * __logf(u1)) * __cosf(6.2831853f * u2);
  }
  const int g = min(max(__float2int_rn(v), 0), 255);
  uint8_t* dst = out + frame_stride * fr + pitch * y + (size_t)x * channels;
  if (channels == 1) {
    dst[0] = (uint8_t)g;
  } else {
    const uint32_t hsh = hash_u32(seed * 31u + pix);
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[c] = (uint8_t)min(max(g + (int)((hsh >> (8 * c)) % 13u) - 6, 0), 255);
  }
}

size_t render_marker_dev_bytes() { return sizeof(RenderMarkerDev); }

// markers_host: n_markers x 16 floats: rvec(3) tvec(3) radius ratio black white, dictionary row, [11..15] unused
void render_pack_marker(const float* spec, int cols, int row_off, const float* K /*fx fy cx cy*/, int w, int h, void* out) {
  RenderMarkerDev m;
  const double rx = spec[0], ry = spec[1], rz = spec[2];
  const double th = sqrt(rx * rx + ry * ry + rz * rz);
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (th > 1e-12) {
    const double x = rx / th, y = ry / th, z = rz / th, c = cos(th), s = sin(th), c1 = 1 - c;
    R[0] = c + c1 * x * x, R[1] = c1 * x * y - s * z, R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z, R[4] = c + c1 * y * y, R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y, R[7] = c1 * y * z + s * x, R[8] = c + c1 * z * z;
  }
  for (int i = 0; i < 9; ++i) m.R[i] = (float)R[i];
  for (int i = 0; i < 3; ++i) m.o[i] = (float)-(R[i] * spec[3] + R[3 + i] * spec[4] + R[6 + i] * spec[5]);  // -R^T t
  m.radius = spec[6];
  m.cols = cols;
  m.W = (float)(2.0 * 3.14159265358979323846 * spec[6] / (1.5 * cols));
  m.L = spec[7] * m.W;
  m.black = spec[8];
  m.white = spec[9];
  m.row_off = row_off;
  // bounding box: project points of the surface (48 angles, both ends incl. the white margin), two pixels of slack
  double x0 = 1e30, y0 = 1e30, x1 = -1e30, y1 = -1e30;
  bool behind = false;
  const double margin = 0.15 * m.L;
  for (int a = 0; a < 48; ++a)
    for (int e = 0; e < 2; ++e) {
      const double ang = 2.0 * 3.14159265358979323846 * a / 48.0;
      const double P[3] = {m.radius * sin(ang), e ? 0.5 * m.L + margin : -0.5 * m.L - margin, -m.radius * cos(ang)};
      const double X = R[0] * P[0] + R[1] * P[1] + R[2] * P[2] + spec[3], Y = R[3] * P[0] + R[4] * P[1] + R[5] * P[2] + spec[4],
                   Z = R[6] * P[0] + R[7] * P[1] + R[8] * P[2] + spec[5];
      if (Z <= 1e-3) {
        behind = true;
        continue;
      }
      const double px = X / Z * K[0] + K[2], py = Y / Z * K[1] + K[3];
      x0 = px < x0 ? px : x0, x1 = px > x1 ? px : x1, y0 = py < y0 ? py : y0, y1 = py > y1 ? py : y1;
    }
  if (behind || x1 < x0) {  // partly behind the camera: no culling
    m.bx0 = 0, m.by0 = 0, m.bx1 = w - 1, m.by1 = h - 1;
  } else {
    m.bx0 = (int)floor(x0) - 2, m.by0 = (int)floor(y0) - 2, m.bx1 = (int)ceil(x1) + 2, m.by1 = (int)ceil(y1) + 2;
  }
  memcpy(out, &m, sizeof(m));
}

int launch_render(float* d_img, int n, int w, int h, const float* d_fparams, const void* d_markers, const int* d_marker_start,
                  const int* d_states, uint8_t* out, size_t pitch, size_t frame_stride, int channels, cudaStream_t stream) {
  const dim3 grid((w + 31) / 32, (h + 7) / 8, n);
  render_scene_kernel<<<grid, 256, 0, stream>>>(d_img, w, h, d_fparams, static_cast<const RenderMarkerDev*>(d_markers), d_marker_start, d_states);
  render_finish_kernel<<<grid, 256, 0, stream>>>(d_img, w, h, d_fparams, out, pitch, frame_stride, channels);
  CTAG_CUDA_CHECK(cudaGetLastError());
  return CTAG_OK;
}

}  // namespace ctag
