// Compressed ingest on the GPU: a batch of baseline JPEG frames -> interleaved BGR frames in the staging
// buffer the fused front kernel reads.  Three kernels per chunk of frames, all on the chunk's stream:
//   jpeg_scan_kernel    one CTA per frame: finds the restart markers of the entropy-coded segment in stream order
//                       (16 bytes per thread, ordered compaction by ballot + block scan) -> start offset of every
//                       restart interval
//   jpeg_huffman_kernel one THREAD per restart interval (the unit that decodes independently: byte aligned, DC
//                       predictors reset): Huffman decode + dequantise, coefficients to a zeroed int16 buffer
//   jpeg_idct_kernel    eight lanes per 8x8 block: the two passes of the integer inverse DCT with a shared-memory
//                       transpose between them, pixels to the component planes
//   jpeg_color420_kernel / jpeg_color_kernel   fancy chroma upsampling + YCbCr -> BGR into the staging buffer
// The arithmetic lives in jpeg_core.cuh (shared with the host tests, which check it byte for byte against cv::imdecode).
#include "common.cuh"
#include "kernels.cuh"
#include "jpeg_core.cuh"

namespace ctag {

using namespace jpeg;

static_assert(sizeof(FrameHeader) % 16 == 0, "FrameHeader is copied to shared memory in 16-byte pieces");

// status[f]: 0 ok, 1 = the number of restart markers found does not match the header
__global__ void __launch_bounds__(1024) jpeg_scan_kernel(const FrameHeader* __restrict__ hdr, const uint8_t* __restrict__ bytes,
                                                        uint32_t* __restrict__ ivl, int* __restrict__ status) {
  __shared__ int wsum[32];
  __shared__ int s_base, s_total;  // restart intervals started so far (interval 0 starts at the segment start)
  const FrameHeader& fh = hdr[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t begin = fh.data_off, end = fh.data_off + fh.data_len;
  uint32_t* out = ivl + fh.interval_first;
  const int expect = fh.n_intervals;
  if (tid == 0) {
    out[0] = begin;
    s_base = 1;
  }
  __syncthreads();
  for (uint32_t chunk = begin & ~15u; chunk < end; chunk += 1024u * 16u) {
    const uint32_t p0 = chunk + 16u * tid;
    uint32_t mask = 0;
    if (p0 < end) {
      const uint4 v = *reinterpret_cast<const uint4*>(bytes + p0);  // frames start 16-byte aligned, the buffer is padded by 32
      const uint32_t w[5] = {v.x, v.y, v.z, v.w, (uint32_t)bytes[p0 + 16]};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t b0 = (w[i >> 2] >> (8 * (i & 3))) & 0xFF, b1 = (w[(i + 1) >> 2] >> (8 * ((i + 1) & 3))) & 0xFF;
        const uint32_t pos = p0 + i;
        // FF D0..D7 inside the segment; a data byte FF is always followed by a stuffed 00, so this is a marker
        if (b0 == 0xFF && (b1 & 0xF8) == 0xD0 && pos >= begin && pos + 1 < end) mask |= 1u << i;
      }
    }
    // ordered compaction: exclusive scan of the per-thread counts over the CTA
    const int cnt = __popc(mask);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int wv = wsum[lane];
      int winc = wv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += u;
      }
      wsum[lane] = winc - wv;
      if (lane == 31) s_total = winc;
    }
    __syncthreads();
    int k = s_base + wsum[warp] + inc - cnt;
    uint32_t m = mask;
    while (m) {
      const int i = __ffs(m) - 1;
      m &= m - 1;
      if (k < expect) out[k] = p0 + i + 2;  // the interval starts behind the two marker bytes
      ++k;
    }
    __syncthreads();
    if (tid == 0) s_base += s_total;
    __syncthreads();
  }
  if (tid == 0) {
    out[expect] = end;
    status[blockIdx.x] = s_base == expect ? 0 : 1;
  }
}

constexpr int kDecodeThreads = 128;

// One thread per restart interval: Huffman decode + dequantise; coefficients (int16, natural order) go to the frame's
// coefficient buffer, which is zero on entry.  The frame's tables sit in shared memory.
__global__ void __launch_bounds__(kDecodeThreads) jpeg_huffman_kernel(const FrameHeader* __restrict__ hdr, const uint8_t* __restrict__ bytes,
                                                                     const uint32_t* __restrict__ ivl, const int* __restrict__ status,
                                                                     int16_t* __restrict__ coefs, size_t coef_stride,
                                                                     uint8_t* __restrict__ last_nz, size_t last_stride) {
  __shared__ __align__(16) FrameHeader fh;
  const int fr = blockIdx.y;
  {
    const uint4* src = reinterpret_cast<const uint4*>(hdr + fr);
    uint4* dst = reinterpret_cast<uint4*>(&fh);
    for (int i = threadIdx.x; i < (int)(sizeof(FrameHeader) / 16); i += kDecodeThreads) dst[i] = src[i];
  }
  __syncthreads();
  const int k = blockIdx.x * kDecodeThreads + threadIdx.x;
  if (k >= fh.n_intervals || status[fr] != 0) return;
  const uint32_t* o = ivl + fh.interval_first;
  decode_interval_coefs(fh, bytes + o[k], bytes + o[k + 1], k, coefs + coef_stride * fr, last_nz + last_stride * fr);
}

// Inverse DCT, eight lanes per 8x8 block (32 blocks per CTA): lane r loads row r of the coefficients (one 16-byte
// load), the block is transposed through shared memory, lane c transforms column c (pass 1 of jidctint.c), the result is
// transposed back, lane r transforms row r (pass 2) and stores its eight pixels with one 8-byte store.
__global__ void __launch_bounds__(256) jpeg_idct_kernel(const FrameHeader* __restrict__ hdr, const int* __restrict__ status,
                                                       const int16_t* __restrict__ coefs, size_t coef_stride,
                                                       const uint8_t* __restrict__ last_nz, size_t last_stride,
                                                       uint8_t* __restrict__ planes, size_t plane_stride) {
  __shared__ int tile[32][72];  // 64 ints per block, padded so that the four blocks of a warp sit in different banks
  const int fr = blockIdx.y;
  const FrameHeader& fh = hdr[fr];
  const int n_blocks = fh.n_blocks;
  const int slot = threadIdx.x >> 3, l = threadIdx.x & 7;
  const int blk = blockIdx.x * 32 + slot;
  const bool live = blk < n_blocks && status[fr] == 0;
  int comp = 0;
  if (fh.ncomp == 3) comp = blk >= fh.block_first[2] ? 2 : (blk >= fh.block_first[1] ? 1 : 0);
  const int local = blk - fh.block_first[comp], bxs = fh.blocks_x[comp];
  const int by = local / bxs, bx = local - by * bxs;
  const int pitch = fh.plane_pitch[comp];
  uint8_t* dst = planes + plane_stride * fr + fh.plane_off[comp] + (size_t)(8 * by + l) * pitch + 8 * bx;
  int last = 0;
  uint4 row = make_uint4(0u, 0u, 0u, 0u);
  if (live) {
    last = last_nz[last_stride * fr + blk];
    row = *reinterpret_cast<const uint4*>(coefs + coef_stride * fr + (size_t)blk * 64 + 8 * l);
  }
  // DC-only blocks (flat areas) are a constant: both passes reduce to shifts
  const int dc = __shfl_sync(0xffffffffu, (int)(int16_t)(row.x & 0xFFFF), threadIdx.x & 24);
  const bool dc_only = last == 0;
  int* t = tile[slot];
  if (!dc_only) {
    const uint32_t w[4] = {row.x, row.y, row.z, row.w};
#pragma unroll
    for (int c = 0; c < 8; ++c) t[8 * l + c] = (int)(int16_t)((w[c >> 1] >> (16 * (c & 1))) & 0xFFFF);
  }
  __syncwarp();  // (the barriers stay outside the divergent parts: flat and detailed blocks share warps)
  int o1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (!dc_only)
    idct_1d(t[l], t[8 + l], t[16 + l], t[24 + l], t[32 + l], t[40 + l], t[48 + l], t[56 + l], 13, &o1[0], &o1[1], &o1[2], &o1[3], &o1[4],
            &o1[5], &o1[6], &o1[7]);
  __syncwarp();
  if (!dc_only) {
#pragma unroll
    for (int r = 0; r < 8; ++r) t[8 * r + l] = jp_descale(o1[r], 13 - 2);
  }
  __syncwarp();
  if (!live) return;
  uint32_t lo, hi;
  if (dc_only) {
    const uint32_t v = jp_clamp(jp_descale(dc << 2, 5) + 128);
    lo = hi = v * 0x01010101u;
  } else {
    int o[8];
    const int* w = t + 8 * l;
    idct_1d(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], 13, &o[0], &o[1], &o[2], &o[3], &o[4], &o[5], &o[6], &o[7]);
    uint32_t px[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) px[c] = jp_clamp(jp_descale(o[c], 13 + 2 + 3) + 128);
    lo = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
    hi = px[4] | (px[5] << 8) | (px[6] << 16) | (px[7] << 24);
  }
  *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
}

// Generic colour kernel (any supported sampling), four pixels per thread.
__global__ void __launch_bounds__(256) jpeg_color_kernel(const FrameHeader* __restrict__ hdr, const uint8_t* __restrict__ planes,
                                                        size_t plane_stride, uint8_t* __restrict__ bgr, size_t pitch, size_t frame_stride) {
  __shared__ FrameHeader fh_s;  // only the leading fields are needed; copy the part in front of the tables
  const int fr = blockIdx.z;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(hdr + fr);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&fh_s);
    for (int i = threadIdx.x; i < (int)(offsetof(FrameHeader, zigzag) / 4); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  const FrameHeader& fh = fh_s;
  const int x0 = 4 * (blockIdx.x * 64 + (threadIdx.x & 63)), y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x0 >= fh.width || y >= fh.height) return;
  const uint8_t* pl = planes + plane_stride * fr;
  uint8_t px[12];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (x0 + i < fh.width) output_pixel(fh, pl, x0 + i, y, px + 3 * i);
    else px[3 * i] = px[3 * i + 1] = px[3 * i + 2] = 0;
  }
  uint8_t* dst = bgr + frame_stride * fr + pitch * y + 3 * (size_t)x0;
  if (x0 + 4 <= fh.width) {
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    d[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
    d[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
    d[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
  } else {
    for (int i = 0; i < 3 * (fh.width - x0); ++i) dst[i] = px[i];
  }
}

// 4:2:0 colour kernel: one thread per 8 x 2 output pixels (4 chroma samples of one chroma row).  The fancy upsampling of
// jdsample.c for both output rows from three chroma rows and six chroma columns, then the fixed-point colour conversion;
// 24-byte row pieces stored as three 8-byte words.
__device__ __forceinline__ void chroma_sums(const uint8_t* __restrict__ near_row, const uint8_t* __restrict__ far_row, int cx0, int cw,
                                            int* cs /*[6]: columns cx0-1 .. cx0+4*/) {
  const uint32_t a = *reinterpret_cast<const uint32_t*>(near_row + cx0), b = *reinterpret_cast<const uint32_t*>(far_row + cx0);
#pragma unroll
  for (int j = 0; j < 4; ++j) cs[j + 1] = 3 * (int)((a >> (8 * j)) & 0xFF) + (int)((b >> (8 * j)) & 0xFF);
  cs[0] = cx0 > 0 ? 3 * near_row[cx0 - 1] + far_row[cx0 - 1] : 0;
  cs[5] = cx0 + 4 < cw ? 3 * near_row[cx0 + 4] + far_row[cx0 + 4] : 0;
}
__device__ __forceinline__ void upsample8(const int* cs, int cx0, int cw, int* out /*[8]*/) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cx = cx0 + j, cur = cs[j + 1];
    out[2 * j] = cx == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + cs[j] + 8) >> 4;
    out[2 * j + 1] = cx + 1 >= cw ? (cur * 4 + 7) >> 4 : (cur * 3 + cs[j + 2] + 7) >> 4;
  }
}
__global__ void __launch_bounds__(256) jpeg_color420_kernel(const FrameHeader* __restrict__ hdr, const uint8_t* __restrict__ planes,
                                                           size_t plane_stride, uint8_t* __restrict__ bgr, size_t pitch,
                                                           size_t frame_stride) {
  const int fr = blockIdx.z;
  const FrameHeader& fh = hdr[fr];
  const int width = fh.width, height = fh.height, cw = fh.chroma_w, ch = fh.chroma_h;
  const int cx0 = 4 * (blockIdx.x * 64 + (threadIdx.x & 63)), cy = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (cx0 >= cw || cy >= ch) return;
  const uint8_t* pl = planes + plane_stride * fr;
  const int ypitch = fh.plane_pitch[0], cpitch = fh.plane_pitch[1];
  const int up = cy > 0 ? cy - 1 : 0, dn = cy + 1 < ch ? cy + 1 : ch - 1;
  int cb[2][8], cr[2][8];
  {
    int cs[6];
    const uint8_t* p = pl + fh.plane_off[1];
    chroma_sums(p + (size_t)cy * cpitch, p + (size_t)up * cpitch, cx0, cw, cs);
    upsample8(cs, cx0, cw, cb[0]);
    chroma_sums(p + (size_t)cy * cpitch, p + (size_t)dn * cpitch, cx0, cw, cs);
    upsample8(cs, cx0, cw, cb[1]);
    p = pl + fh.plane_off[2];
    chroma_sums(p + (size_t)cy * cpitch, p + (size_t)up * cpitch, cx0, cw, cs);
    upsample8(cs, cx0, cw, cr[0]);
    chroma_sums(p + (size_t)cy * cpitch, p + (size_t)dn * cpitch, cx0, cw, cs);
    upsample8(cs, cx0, cw, cr[1]);
  }
  const int x0 = 2 * cx0;
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int y = 2 * cy + v;
    if (y >= height) break;
    const uint2 yy = *reinterpret_cast<const uint2*>(pl + fh.plane_off[0] + (size_t)y * ypitch + x0);
    uint8_t px[24];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int yv = (int)(((i < 4 ? yy.x : yy.y) >> (8 * (i & 3))) & 0xFF);
      ycc_to_bgr(yv, cb[v][i], cr[v][i], px + 3 * i);
    }
    uint8_t* dst = bgr + frame_stride * fr + pitch * (size_t)y + 3 * (size_t)x0;
    if (x0 + 8 <= width) {
      uint2* d = reinterpret_cast<uint2*>(dst);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const uint8_t* s = px + 8 * q;
        d[q] = make_uint2(s[0] | (s[1] << 8) | (s[2] << 16) | ((uint32_t)s[3] << 24), s[4] | (s[5] << 8) | (s[6] << 16) | ((uint32_t)s[7] << 24));
      }
    } else {
      for (int i = 0; i < 3 * (width - x0); ++i) dst[i] = px[i];
    }
  }
}

size_t jpeg_header_bytes() { return sizeof(FrameHeader); }

int jpeg_parse_frame(const uint8_t* data, size_t len, void* header_out, size_t* scan_off, size_t* scan_len, int* w, int* h,
                     int* n_intervals, size_t* plane_bytes) {
  FrameHeader* fh = static_cast<FrameHeader*>(header_out);
  const int rc = parse_header(data, len, *fh, scan_off, scan_len);
  if (rc != JP_OK) return rc;
  *w = fh->width;
  *h = fh->height;
  *n_intervals = fh->n_intervals;
  *plane_bytes = (size_t)fh->plane_bytes;
  return 0;
}

void jpeg_place_frame(void* header, uint32_t data_off, uint32_t data_len, int interval_first) {
  FrameHeader* fh = static_cast<FrameHeader*>(header);
  fh->data_off = data_off;
  fh->data_len = data_len;
  fh->interval_first = interval_first;
}

size_t jpeg_coef_bytes_per_block() { return 64 * sizeof(int16_t); }

int jpeg_frame_blocks(const void* header) { return static_cast<const FrameHeader*>(header)->n_blocks; }
int jpeg_frame_is_420(const void* header) {
  const FrameHeader* fh = static_cast<const FrameHeader*>(header);
  return fh->ncomp == 3 && fh->chroma_hs == 2 && fh->chroma_vs == 2;
}

int launch_jpeg_decode(const void* d_hdr, int n, const uint8_t* d_bytes, uint32_t* d_ivl, int max_intervals, int16_t* d_coefs,
                       size_t coef_stride, uint8_t* d_last, size_t last_stride, int max_blocks, uint8_t* d_planes, size_t plane_stride,
                       uint8_t* d_bgr, size_t pitch, size_t frame_stride, int w, int h, int all_420, int* d_status, cudaStream_t stream,
                       int* launches) {
  const FrameHeader* hdr = static_cast<const FrameHeader*>(d_hdr);
  CTAG_CUDA_CHECK(cudaMemsetAsync(d_coefs, 0, coef_stride * sizeof(int16_t) * n, stream));
  jpeg_scan_kernel<<<n, 1024, 0, stream>>>(hdr, d_bytes, d_ivl, d_status);
  jpeg_huffman_kernel<<<dim3((max_intervals + kDecodeThreads - 1) / kDecodeThreads, n), kDecodeThreads, 0, stream>>>(
      hdr, d_bytes, d_ivl, d_status, d_coefs, coef_stride, d_last, last_stride);
  jpeg_idct_kernel<<<dim3((max_blocks + 31) / 32, n), 256, 0, stream>>>(hdr, d_status, d_coefs, coef_stride, d_last, last_stride, d_planes,
                                                                         plane_stride);
  if (all_420) {
    const int cw = (w + 1) / 2, ch = (h + 1) / 2;
    jpeg_color420_kernel<<<dim3((cw + 255) / 256, (ch + 3) / 4, n), 256, 0, stream>>>(hdr, d_planes, plane_stride, d_bgr, pitch, frame_stride);
  } else {
    jpeg_color_kernel<<<dim3((w + 255) / 256, (h + 3) / 4, n), 256, 0, stream>>>(hdr, d_planes, plane_stride, d_bgr, pitch, frame_stride);
  }
  CTAG_CUDA_CHECK(cudaGetLastError());
  if (launches) *launches += 4;
  return CTAG_OK;
}

}  // namespace ctag
