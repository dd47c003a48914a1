// Baseline JPEG decoding cores for the compressed-ingest path, shared by the CUDA kernels (jpeg.cu) and
// the host logic tests (tests/host_harness).  The reference gets its frames from cv::imread / cv::VideoCapture
// (main.cpp:29,45-52), i.e. decoded on the CPU by OpenCV's codec (libjpeg-turbo for JPEG); here a batch of JPEG byte
// strings is decoded on the GPU straight into the BGR staging buffer of the detect pipeline.
//
// What makes a JPEG bitstream parallel is its restart markers (DRI / RSTn, ITU-T T.81 E.1.4, F.1.2.3): after every
// restart interval the entropy coder is byte aligned and the DC predictors are reset, so the intervals decode
// independently -- one GPU thread per interval.  A stream without restart markers is one sequential chain and is
// refused (the caller falls back to nvJPEG).
//
// Everything downstream of the entropy decoder restates the arithmetic of the IJG / libjpeg-turbo reference decoder
// with its default settings -- the accurate integer inverse DCT (jidctint.c), "fancy" triangle-filter chroma
// upsampling (jdsample.c) and the 16-bit fixed-point YCbCr -> RGB tables (jdcolor.c) -- so that the decoded pixels are
// the ones cv::imdecode would produce (tests/test_jpeg_core.py compares them byte for byte).
//
// Supported: baseline / extended sequential DCT (SOF0, SOF1), 8-bit samples, one interleaved scan, 1 component (gray)
// or 3 components (YCbCr) sampled 4:4:4, 4:2:2 (h2v1) or 4:2:0 (h2v2), 8-bit or 16-bit quantisation tables.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define JP_HD __host__ __device__ __forceinline__
#else
#define JP_HD inline
#endif

namespace ctag {
namespace jpeg {

constexpr int kLutBits = 9;  // codes up to this length decode with one table look-up

struct HuffTable {
  uint16_t lut[1 << kLutBits];  // (length << 8) | symbol for codes of at most kLutBits bits, 0 = longer code
  int32_t maxcode[18];          // largest code of each length (left aligned comparison is done by the caller), -1 = none
  int32_t valoff[17];           // huffval index of the first code of each length minus that code
  uint8_t vals[256];
};

struct FrameHeader {
  uint32_t data_off;   // entropy-coded segment: offset into the batch's byte buffer
  uint32_t data_len;   // up to (not including) the EOI marker
  int32_t width, height, ncomp;
  int32_t hmax, vmax;  // largest sampling factors
  int32_t comp_h[3], comp_v[3];
  int32_t comp_tq[3], comp_td[3], comp_ta[3];
  int32_t restart_interval;  // MCUs per restart interval (> 0)
  int32_t mcus_x, mcus_y;
  int32_t n_intervals;         // ceil(mcus_x * mcus_y / restart_interval)
  int32_t interval_first;      // index of this frame's first interval in the batch's interval-offset array
  int32_t plane_off[3];        // byte offset of each component plane inside the frame's plane buffer
  int32_t plane_pitch[3];      // = mcus_x * 8 * comp_h
  int32_t plane_rows[3];       // = mcus_y * 8 * comp_v
  int32_t plane_bytes;         // all planes of the frame
  int32_t blocks_x[3];         // 8x8 blocks per row of each component (= plane_pitch / 8)
  int32_t block_first[3];      // index of each component's first block in the frame's coefficient buffer
  int32_t n_blocks;            // blocks of all components
  int32_t chroma_hs, chroma_vs;  // luma / chroma sampling ratio per axis (1 or 2)
  int32_t chroma_w, chroma_h;    // real (downsampled) chroma size: ceil(width / hs), ceil(height / vs)
  int32_t pad_[4];             // sizeof(FrameHeader) is a multiple of 16 (copied to shared memory in 16-byte pieces)
  uint8_t zigzag[64];          // natural (row-major) position of the k-th transmitted coefficient
  uint16_t quant[4][64];       // quantiser steps in zig-zag order (as transmitted)
  HuffTable huff[4];           // [0] DC table 0, [1] DC table 1, [2] AC table 0, [3] AC table 1
};

// natural (row-major) position of the k-th zig-zag coefficient
JP_HD int zigzag_natural(int k) {
  const uint8_t z[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                         41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                         30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  return z[k];
}

// ---- bit reader over an entropy-coded segment (byte stuffing removed on the fly; T.81 F.2.2.5) ----------------------
struct BitReader {
  const uint8_t* p;    // next byte
  const uint8_t* end;  // end of the interval's bytes (a marker or the end of the segment)
  uint64_t acc;        // bits, left aligned
  int nbits;
};

JP_HD void br_init(BitReader& br, const uint8_t* p, const uint8_t* end) {
  br.p = p;
  br.end = end;
  br.acc = 0;
  br.nbits = 0;
}

// keeps at least 32 valid bits in the accumulator (zero bits once the interval's data are exhausted)
JP_HD void br_fill(BitReader& br) {
  while (br.nbits <= 56) {
    uint32_t b = 0;
    if (br.p < br.end) {
      b = *br.p++;
      if (b == 0xFF) {
        if (br.p < br.end && *br.p == 0x00) ++br.p;  // stuffed zero
        else {  // a marker: the interval ends here
          br.p = br.end;
          b = 0;
        }
      }
    }
    br.acc |= (uint64_t)b << (56 - br.nbits);
    br.nbits += 8;
  }
}
// four bytes starting at p, first byte in the most significant position
JP_HD uint32_t load_be32(const uint8_t* p) {
#ifdef __CUDA_ARCH__
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t lo = w[0], hi = w[1];
  const uint32_t v = __funnelshift_r(lo, hi, 8 * (uint32_t)(a & 3));  // little-endian bytes p[0..3]
  return __byte_perm(v, 0, 0x0123);
#else
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
#endif
}

// The same contract as br_fill, four bytes at a time: entropy-coded data hold an FF byte (stuffing or the marker that ends
// the interval) in about 1.5 % of their words; every other word goes into the accumulator with one shift.  Words are read
// up to 7 bytes beyond `end` (never consumed: an interval ends with a marker); the caller pads its buffer.
JP_HD void br_fill_fast(BitReader& br) {
  if (br.nbits > 32) return;
  if (br.p < br.end) {
    const uint32_t w = load_be32(br.p);
    const uint32_t n = ~w;
    if (((n - 0x01010101u) & ~n & 0x80808080u) == 0) {  // no FF byte among the four
      br.acc |= (uint64_t)w << (32 - br.nbits);
      br.nbits += 32;
      br.p += 4;
      return;
    }
  }
  br_fill(br);
}
JP_HD uint32_t br_peek(const BitReader& br, int n) { return (uint32_t)(br.acc >> (64 - n)); }
JP_HD void br_skip(BitReader& br, int n) {
  br.acc <<= n;
  br.nbits -= n;
}

// one Huffman symbol (T.81 F.2.2.3 with a kLutBits look-ahead table in front)
JP_HD int huff_decode(BitReader& br, const HuffTable& t) {
  const uint32_t look = br_peek(br, kLutBits);
  const uint32_t e = t.lut[look];
  if (e) {
    br_skip(br, (int)(e >> 8));
    return (int)(e & 0xFF);
  }
  const uint32_t w = br_peek(br, 16);
  for (int l = kLutBits + 1; l <= 16; ++l) {
    const int32_t code = (int32_t)(w >> (16 - l));
    if (code <= t.maxcode[l]) {
      br_skip(br, l);
      return t.vals[(code + t.valoff[l]) & 0xFF];
    }
  }
  br_skip(br, 16);  // corrupt stream: consume and go on
  return 0;
}

// T.81 F.2.2.1 EXTEND: the s-bit magnitude category value
JP_HD int br_receive_extend(BitReader& br, int s) {
  if (s == 0) return 0;
  const int v = (int)br_peek(br, s);
  br_skip(br, s);
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

// One 8x8 block: entropy decode + dequantise into natural order.  `coef` must be zero on entry.  Returns the index of
// the last non-zero zig-zag coefficient (0 when only DC).
JP_HD int decode_block(BitReader& br, const HuffTable& dc, const HuffTable& ac, const uint16_t* q, int& dc_pred, int* coef) {
  br_fill(br);
  const int s = huff_decode(br, dc);
  br_fill(br);
  dc_pred += br_receive_extend(br, s & 15);
  coef[0] = dc_pred * q[0];
  int k = 1, last = 0;
  while (k < 64) {
    br_fill(br);
    const int rs = huff_decode(br, ac);
    const int r = rs >> 4, sz = rs & 15;
    if (sz == 0) {
      if (r != 15) break;  // EOB
      k += 16;
      continue;
    }
    k += r;
    if (k > 63) break;  // corrupt stream
    const int v = br_receive_extend(br, sz);
    coef[zigzag_natural(k)] = v * q[k];
    last = k;
    ++k;
  }
  return last;
}

// The same block, for the two-kernel GPU path: dequantised coefficients go to `out` (64 int16 in natural order, zero on
// entry) as they are decoded; `zz` is the zig-zag table (shared memory on the device).
JP_HD int decode_block_coefs(BitReader& br, const HuffTable& dc, const HuffTable& ac, const uint16_t* q, const uint8_t* zz, int& dc_pred,
                             int16_t* out) {
  br_fill_fast(br);
  const int s = huff_decode(br, dc);
  dc_pred += br_receive_extend(br, s & 15);
  out[0] = (int16_t)(dc_pred * q[0]);
  int k = 1, last = 0;
  while (k < 64) {
    br_fill_fast(br);
    const int rs = huff_decode(br, ac);
    const int r = rs >> 4, sz = rs & 15;
    if (sz == 0) {
      if (r != 15) break;  // EOB
      k += 16;
      continue;
    }
    k += r;
    if (k > 63) break;  // corrupt stream
    const int v = br_receive_extend(br, sz);
    out[zz[k]] = (int16_t)(v * q[k]);
    last = k;
    ++k;
  }
  return last;
}

// ---- accurate integer inverse DCT (IJG jidctint.c, "islow"): 13-bit constants, 2 extra bits between the passes -------
JP_HD int jp_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
JP_HD uint8_t jp_clamp(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

JP_HD void idct_1d(const int in0, const int in1, const int in2, const int in3, const int in4, const int in5, const int in6,
                   const int in7, int shift_even, int* o0, int* o1, int* o2, int* o3, int* o4, int* o5, int* o6, int* o7) {
  // even part
  int z2 = in2, z3 = in6;
  int z1 = (z2 + z3) * 4433;
  int tmp2 = z1 + z3 * (-15137);
  int tmp3 = z1 + z2 * 6270;
  z2 = in0;
  z3 = in4;
  int tmp0 = (z2 + z3) << shift_even;
  int tmp1 = (z2 - z3) << shift_even;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  // odd part
  tmp0 = in7;
  tmp1 = in5;
  tmp2 = in3;
  tmp3 = in1;
  z1 = tmp0 + tmp3;
  z2 = tmp1 + tmp2;
  z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * 9633;
  tmp0 *= 2446;
  tmp1 *= 16819;
  tmp2 *= 25172;
  tmp3 *= 12299;
  z1 *= -7373;
  z2 *= -20995;
  z3 *= -16069;
  z4 *= -3196;
  z3 += z5;
  z4 += z5;
  tmp0 += z1 + z3;
  tmp1 += z2 + z4;
  tmp2 += z2 + z3;
  tmp3 += z1 + z4;
  *o0 = tmp10 + tmp3;
  *o7 = tmp10 - tmp3;
  *o1 = tmp11 + tmp2;
  *o6 = tmp11 - tmp2;
  *o2 = tmp12 + tmp1;
  *o5 = tmp12 - tmp1;
  *o3 = tmp13 + tmp0;
  *o4 = tmp13 - tmp0;
}

// coef: 64 dequantised coefficients in natural order; out: 8 rows of `pitch` bytes.  `last_zz`: index of the last
// non-zero zig-zag coefficient (a DC-only block is a constant).
JP_HD void idct_islow_store(const int* coef, int last_zz, uint8_t* out, int pitch) {
  if (last_zz == 0) {
    // both passes reduce to shifts: ws = dc << 2, pixel = clamp(descale(ws, 5) + 128)
    const uint8_t v = jp_clamp(jp_descale(coef[0] << 2, 5) + 128);
    const uint32_t w = v * 0x01010101u;
    for (int r = 0; r < 8; ++r) {
      uint32_t* d = reinterpret_cast<uint32_t*>(out + (size_t)r * pitch);
      d[0] = w;
      d[1] = w;
    }
    return;
  }
  int ws[64];
  for (int c = 0; c < 8; ++c) {  // pass 1: columns, results scaled up by 2^PASS1_BITS
    int o[8];
    idct_1d(coef[c], coef[8 + c], coef[16 + c], coef[24 + c], coef[32 + c], coef[40 + c], coef[48 + c], coef[56 + c], 13, &o[0],
            &o[1], &o[2], &o[3], &o[4], &o[5], &o[6], &o[7]);
    for (int r = 0; r < 8; ++r) ws[8 * r + c] = jp_descale(o[r], 13 - 2);
  }
  for (int r = 0; r < 8; ++r) {  // pass 2: rows, remove the 2^PASS1_BITS and the factor 8 of the 2-D transform
    int o[8];
    const int* w = ws + 8 * r;
    idct_1d(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], 13, &o[0], &o[1], &o[2], &o[3], &o[4], &o[5], &o[6], &o[7]);
    uint8_t px[8];
    for (int c = 0; c < 8; ++c) px[c] = jp_clamp(jp_descale(o[c], 13 + 2 + 3) + 128);
    uint32_t* d = reinterpret_cast<uint32_t*>(out + (size_t)r * pitch);
    d[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
    d[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
  }
}

// ---- one restart interval: `restart_interval` MCUs starting at MCU index interval * restart_interval -----------------
// bytes: the batch's byte buffer; [begin, end) the interval's entropy-coded bytes; planes: the frame's plane buffer.
JP_HD void decode_interval(const FrameHeader& fh, const uint8_t* begin, const uint8_t* end, int interval, uint8_t* planes) {
  BitReader br;
  br_init(br, begin, end);
  int dc_pred[3] = {0, 0, 0};
  const int total = fh.mcus_x * fh.mcus_y;
  int m = interval * fh.restart_interval;
  const int m_end = m + fh.restart_interval < total ? m + fh.restart_interval : total;
  int mx = m % fh.mcus_x, my = m / fh.mcus_x;
  for (; m < m_end; ++m) {
    for (int c = 0; c < fh.ncomp; ++c) {
      const HuffTable& dct = fh.huff[fh.comp_td[c] & 1];
      const HuffTable& act = fh.huff[2 + (fh.comp_ta[c] & 1)];
      const uint16_t* q = fh.quant[fh.comp_tq[c] & 3];
      const int pitch = fh.plane_pitch[c];
      uint8_t* plane = planes + fh.plane_off[c];
      for (int by = 0; by < fh.comp_v[c]; ++by)
        for (int bx = 0; bx < fh.comp_h[c]; ++bx) {
          int coef[64];
          for (int i = 0; i < 64; ++i) coef[i] = 0;
          const int last = decode_block(br, dct, act, q, dc_pred[c], coef);
          const int px = (mx * fh.comp_h[c] + bx) * 8, py = (my * fh.comp_v[c] + by) * 8;
          idct_islow_store(coef, last, plane + (size_t)py * pitch + px, pitch);
        }
    }
    if (++mx == fh.mcus_x) mx = 0, ++my;
  }
}

// The interval's blocks as coefficients: coefs = the frame's coefficient buffer (n_blocks x 64 int16, zeroed), last_nz = one
// byte per block (index of its last non-zero zig-zag coefficient).
JP_HD void decode_interval_coefs(const FrameHeader& fh, const uint8_t* begin, const uint8_t* end, int interval, int16_t* coefs,
                                 uint8_t* last_nz) {
  BitReader br;
  br_init(br, begin, end);
  int dc_pred[3] = {0, 0, 0};
  const int total = fh.mcus_x * fh.mcus_y;
  int m = interval * fh.restart_interval;
  const int m_end = m + fh.restart_interval < total ? m + fh.restart_interval : total;
  int mx = m % fh.mcus_x, my = m / fh.mcus_x;
  for (; m < m_end; ++m) {
    for (int c = 0; c < fh.ncomp; ++c) {
      const HuffTable& dct = fh.huff[fh.comp_td[c] & 1];
      const HuffTable& act = fh.huff[2 + (fh.comp_ta[c] & 1)];
      const uint16_t* q = fh.quant[fh.comp_tq[c] & 3];
      for (int by = 0; by < fh.comp_v[c]; ++by)
        for (int bx = 0; bx < fh.comp_h[c]; ++bx) {
          const int blk = fh.block_first[c] + (my * fh.comp_v[c] + by) * fh.blocks_x[c] + mx * fh.comp_h[c] + bx;
          // 12-bit precision would overflow int16 after dequantisation; 8-bit samples cannot (|coef * q| < 2^15)
          last_nz[blk] = (uint8_t)decode_block_coefs(br, dct, act, q, fh.zigzag, dc_pred[c], coefs + (size_t)blk * 64);
        }
    }
    if (++mx == fh.mcus_x) mx = 0, ++my;
  }
}

// one block of the coefficient buffer -> pixels (host composition of what the GPU's idct kernel does with 8 lanes)
JP_HD void idct_block_from_coefs(const int16_t* c16, int last_zz, uint8_t* out, int pitch) {
  int coef[64];
  for (int i = 0; i < 64; ++i) coef[i] = c16[i];
  idct_islow_store(coef, last_zz, out, pitch);
}

// ---- chroma upsampling + colour conversion for one output pixel ----------------------------------------------------
// "Fancy" upsampling of jdsample.c: each output sample is 3/4 of the nearer and 1/4 of the further input sample per
// axis; h2v2 rounds with the alternating biases 8 / 7 after both axes, h2v1 with 1 / 2.  Rows and columns beyond the
// component's real (downsampled) size replicate the edge sample, the first / last output column of a row uses the edge
// sample alone, as the IJG code does.
JP_HD int chroma_at(const uint8_t* plane, int pitch, int cw, int ch, int hs, int vs, int x, int y) {
  if (hs == 1 && vs == 1) return plane[(size_t)y * pitch + x];
  const int cx = x >> 1;
  if (vs == 1) {  // h2v1
    const int v = plane[(size_t)y * pitch + cx];
    if (x & 1) return cx + 1 < cw ? (3 * v + plane[(size_t)y * pitch + cx + 1] + 2) >> 2 : v;
    return cx > 0 ? (3 * v + plane[(size_t)y * pitch + cx - 1] + 1) >> 2 : v;
  }
  // h2v2
  const int cy = y >> 1;
  int oy = (y & 1) ? cy + 1 : cy - 1;  // the further input row
  oy = oy < 0 ? 0 : (oy >= ch ? ch - 1 : oy);
  const uint8_t* r0 = plane + (size_t)cy * pitch;
  const uint8_t* r1 = plane + (size_t)oy * pitch;
  const int cur = 3 * r0[cx] + r1[cx];
  if (x & 1) {
    if (cx + 1 >= cw) return (cur * 4 + 7) >> 4;
    return (cur * 3 + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
  }
  if (cx == 0) return (cur * 4 + 8) >> 4;
  return (cur * 3 + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
}

// jdcolor.c: 16-bit fixed point, the four terms rounded the way the table-driven code does
JP_HD void ycc_to_bgr(int y, int cb, int cr, uint8_t* bgr) {
  cb -= 128;
  cr -= 128;
  const int r = y + ((91881 * cr + 32768) >> 16);
  const int g = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
  const int b = y + ((116130 * cb + 32768) >> 16);
  bgr[0] = jp_clamp(b);
  bgr[1] = jp_clamp(g);
  bgr[2] = jp_clamp(r);
}

JP_HD void output_pixel(const FrameHeader& fh, const uint8_t* planes, int x, int y, uint8_t* bgr) {
  const int yv = planes[fh.plane_off[0] + (size_t)y * fh.plane_pitch[0] + x];
  if (fh.ncomp == 1) {
    bgr[0] = bgr[1] = bgr[2] = (uint8_t)yv;
    return;
  }
  const int hs = fh.chroma_hs, vs = fh.chroma_vs, cw = fh.chroma_w, ch = fh.chroma_h;
  const int cb = chroma_at(planes + fh.plane_off[1], fh.plane_pitch[1], cw, ch, hs, vs, x, y);
  const int cr = chroma_at(planes + fh.plane_off[2], fh.plane_pitch[2], cw, ch, hs, vs, x, y);
  ycc_to_bgr(yv, cb, cr, bgr);
}

// ---- host side: header parsing (T.81 Annex B) ------------------------------------------------------------------------
enum { JP_OK = 0, JP_BAD = 1, JP_UNSUPPORTED = 2 };

inline void build_huff(HuffTable& t, const uint8_t* counts /*16*/, const uint8_t* vals, int nvals) {
  memset(&t, 0, sizeof(t));
  for (int i = 0; i < nvals && i < 256; ++i) t.vals[i] = vals[i];
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    t.valoff[l] = k - code;
    const int n = counts[l - 1];
    if (n) {
      if (l <= kLutBits)
        for (int i = 0; i < n; ++i) {
          const int c = code + i;
          const uint16_t e = (uint16_t)((l << 8) | vals[k + i]);
          for (int f = 0; f < (1 << (kLutBits - l)); ++f) t.lut[(c << (kLutBits - l)) | f] = e;
        }
      k += n;
      code += n;
      t.maxcode[l] = code - 1;
    } else {
      t.maxcode[l] = -1;
    }
    code <<= 1;
  }
  t.maxcode[17] = 0x7fffffff;
}

// Parses the headers of one JPEG and fills `fh` except data_off / interval_first (relative to the caller's buffers):
// *scan_off / *scan_len receive the entropy-coded segment inside `data`.
inline int parse_header(const uint8_t* data, size_t len, FrameHeader& fh, size_t* scan_off, size_t* scan_len) {
  memset(&fh, 0, sizeof(fh));
  if (len < 4 || data[0] != 0xFF || data[1] != 0xD8) return JP_BAD;
  size_t p = 2;
  bool have_sof = false;
  bool have_dc[2] = {false, false}, have_ac[2] = {false, false};
  int comp_id[3] = {0, 0, 0};
  while (p + 4 <= len) {
    if (data[p] != 0xFF) return JP_BAD;
    const int m = data[p + 1];
    if (m == 0xFF) {  // fill byte
      ++p;
      continue;
    }
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) {
      p += 2;
      continue;
    }
    if (m == 0xD9) return JP_BAD;  // EOI before a scan
    const size_t seg = ((size_t)data[p + 2] << 8) | data[p + 3];
    if (seg < 2 || p + 2 + seg > len) return JP_BAD;
    const uint8_t* s = data + p + 4;
    const size_t n = seg - 2;
    if (m == 0xDB) {  // DQT
      size_t i = 0;
      while (i < n) {
        const int pq = s[i] >> 4, tq = s[i] & 15;
        ++i;
        if (tq > 3 || i + (pq ? 128 : 64) > n) return JP_BAD;
        for (int k = 0; k < 64; ++k) {
          fh.quant[tq][k] = pq ? (uint16_t)((s[i] << 8) | s[i + 1]) : s[i];
          i += pq ? 2 : 1;
        }
      }
    } else if (m == 0xC0 || m == 0xC1) {  // SOF0 / SOF1
      if (n < 6 || s[0] != 8) return JP_UNSUPPORTED;
      fh.height = (s[1] << 8) | s[2];
      fh.width = (s[3] << 8) | s[4];
      fh.ncomp = s[5];
      if ((fh.ncomp != 1 && fh.ncomp != 3) || n < (size_t)(6 + 3 * fh.ncomp) || fh.width <= 0 || fh.height <= 0) return JP_UNSUPPORTED;
      for (int c = 0; c < fh.ncomp; ++c) {
        comp_id[c] = s[6 + 3 * c];
        fh.comp_h[c] = s[7 + 3 * c] >> 4;
        fh.comp_v[c] = s[7 + 3 * c] & 15;
        fh.comp_tq[c] = s[8 + 3 * c] & 3;
      }
      have_sof = true;
    } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return JP_UNSUPPORTED;  // progressive, lossless, arithmetic
    } else if (m == 0xC4) {  // DHT
      size_t i = 0;
      while (i + 17 <= n) {
        const int tc = s[i] >> 4, th = s[i] & 15;
        int total = 0;
        for (int l = 0; l < 16; ++l) total += s[i + 1 + l];
        if (tc > 1 || th > 1 || total > 256 || i + 17 + total > n) return th > 1 ? JP_UNSUPPORTED : JP_BAD;
        build_huff(fh.huff[2 * tc + th], s + i + 1, s + i + 17, total);
        (tc ? have_ac : have_dc)[th] = true;
        i += 17 + total;
      }
    } else if (m == 0xDD) {  // DRI
      if (n < 2) return JP_BAD;
      fh.restart_interval = (s[0] << 8) | s[1];
    } else if (m == 0xDA) {  // SOS
      if (!have_sof || n < 1 || s[0] != fh.ncomp || n < (size_t)(1 + 2 * fh.ncomp + 3)) return JP_UNSUPPORTED;
      for (int c = 0; c < fh.ncomp; ++c) {
        if (s[1 + 2 * c] != comp_id[c]) return JP_UNSUPPORTED;
        fh.comp_td[c] = s[2 + 2 * c] >> 4;
        fh.comp_ta[c] = s[2 + 2 * c] & 15;
        if (fh.comp_td[c] > 1 || fh.comp_ta[c] > 1 || !have_dc[fh.comp_td[c]] || !have_ac[fh.comp_ta[c]]) return JP_UNSUPPORTED;
      }
      if (fh.restart_interval <= 0) return JP_UNSUPPORTED;  // one sequential chain: not for this decoder
      // sampling: luma carries the largest factors, chroma 1x1
      fh.hmax = fh.comp_h[0];
      fh.vmax = fh.comp_v[0];
      if (fh.ncomp == 1) {
        fh.comp_h[0] = fh.comp_v[0] = fh.hmax = fh.vmax = 1;  // a single component is never subsampled (T.81 A.2.2)
      } else {
        const bool ok = fh.comp_h[1] == 1 && fh.comp_v[1] == 1 && fh.comp_h[2] == 1 && fh.comp_v[2] == 1 &&
                        ((fh.hmax == 1 && fh.vmax == 1) || (fh.hmax == 2 && fh.vmax == 1) || (fh.hmax == 2 && fh.vmax == 2));
        if (!ok) return JP_UNSUPPORTED;
      }
      fh.mcus_x = (fh.width + 8 * fh.hmax - 1) / (8 * fh.hmax);
      fh.mcus_y = (fh.height + 8 * fh.vmax - 1) / (8 * fh.vmax);
      fh.n_intervals = (fh.mcus_x * fh.mcus_y + fh.restart_interval - 1) / fh.restart_interval;
      int off = 0;
      for (int c = 0; c < fh.ncomp; ++c) {
        fh.plane_off[c] = off;
        fh.plane_pitch[c] = fh.mcus_x * 8 * fh.comp_h[c];
        fh.plane_rows[c] = fh.mcus_y * 8 * fh.comp_v[c];
        off += fh.plane_pitch[c] * fh.plane_rows[c];
      }
      fh.plane_bytes = (off + 255) & ~255;
      int nb = 0;
      for (int c = 0; c < fh.ncomp; ++c) {
        fh.blocks_x[c] = fh.mcus_x * fh.comp_h[c];
        fh.block_first[c] = nb;
        nb += fh.blocks_x[c] * fh.mcus_y * fh.comp_v[c];
      }
      fh.n_blocks = nb;
      fh.chroma_hs = fh.ncomp == 3 ? fh.hmax / fh.comp_h[1] : 1;
      fh.chroma_vs = fh.ncomp == 3 ? fh.vmax / fh.comp_v[1] : 1;
      fh.chroma_w = (fh.width + fh.chroma_hs - 1) / fh.chroma_hs;
      fh.chroma_h = (fh.height + fh.chroma_vs - 1) / fh.chroma_vs;
      for (int k = 0; k < 64; ++k) fh.zigzag[k] = (uint8_t)zigzag_natural(k);
      // entropy-coded data: up to the EOI marker.  Encoders end the file with it; only when they do not is the segment
      // searched (stuffed bytes and restart markers skipped) -- a megabyte-long byte loop per frame otherwise
      const size_t start = p + 2 + seg;
      size_t e;
      if (len >= start + 2 && data[len - 2] == 0xFF && data[len - 1] == 0xD9) {
        e = len - 2;
      } else {
        e = start;
        while (e + 1 < len) {
          if (data[e] == 0xFF && data[e + 1] != 0x00 && !(data[e + 1] >= 0xD0 && data[e + 1] <= 0xD7) && data[e + 1] != 0xFF) break;
          ++e;
        }
        if (e + 1 >= len) e = len;  // no EOI: take everything
      }
      *scan_off = start;
      *scan_len = e - start;
      return JP_OK;
    }
    p += 2 + seg;
  }
  return JP_BAD;
}

// positions of the interval starts inside the entropy-coded segment (the GPU path finds them with a kernel; the host
// tests use this)
inline int find_intervals_host(const uint8_t* scan, size_t len, int n_intervals, uint32_t* off /*[n_intervals + 1]*/) {
  int k = 0;
  off[k++] = 0;
  for (size_t i = 0; i + 1 < len; ++i)
    if (scan[i] == 0xFF && scan[i + 1] >= 0xD0 && scan[i + 1] <= 0xD7) {
      if (k >= n_intervals) return JP_BAD;
      off[k++] = (uint32_t)(i + 2);
      ++i;
    }
  if (k != n_intervals) return JP_BAD;
  off[n_intervals] = (uint32_t)len;
  return JP_OK;
}

}  // namespace jpeg
}  // namespace ctag
