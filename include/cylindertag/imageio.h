// imageio.h -- minimal frame ingest for programs that use the mirror class without OpenCV:
// uncompressed BMP (8-bit palettised gray, 24-bit BGR, 32-bit BGRA; bottom-up or top-down) and binary PGM/PPM.
// Stands in for cv::imread at main.cpp:29 for the formats the reference's sample data uses (test.bmp).
// With CTAG_WITH_ZLIB defined (link -lz) it also reads non-interlaced 8-bit PNG (gray, gray+alpha, RGB, RGBA, palette).
#pragma once
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#ifdef CTAG_WITH_ZLIB
#include <zlib.h>
#endif

#include "CylinderTag.h"

namespace ctag_api {

struct Image {
  std::vector<uint8_t> data;  // rows x cols x channels, tightly packed; 3-channel data is B,G,R like cv::imread
  int rows = 0, cols = 0, channels = 0;
  bool empty() const { return data.empty(); }
  ImageView view() const { return ImageView{data.data(), rows, cols, (size_t)cols * channels, channels}; }
};

namespace detail {
inline uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
}  // namespace detail

#ifdef CTAG_WITH_ZLIB
namespace detail {
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace detail

// PNG (ISO/IEC 15948): 8-bit, non-interlaced.  Colour images come back as B,G,R like cv::imread; alpha is dropped.
inline Image read_png(const std::vector<uint8_t>& buf) {
  Image img;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (buf.size() < 33 || std::memcmp(buf.data(), sig, 8) != 0) return img;
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8_t> idat, pal;
  for (size_t pos = 8; pos + 12 <= buf.size();) {
    const uint32_t len = detail::be32(&buf[pos]);
    if (pos + 12 + (size_t)len > buf.size()) return img;
    const uint8_t* type = &buf[pos + 4];
    const uint8_t* d = &buf[pos + 8];
    if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
      w = detail::be32(d), h = detail::be32(d + 4), depth = d[8], ctype = d[9], interlace = d[12];
    } else if (!std::memcmp(type, "PLTE", 4)) {
      pal.assign(d, d + len);
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), d, d + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + (size_t)len;
  }
  int spp = 0;  // samples per pixel
  switch (ctype) {
    case 0: spp = 1; break;
    case 2: spp = 3; break;
    case 3: spp = 1; break;
    case 4: spp = 2; break;
    case 6: spp = 4; break;
    default: return img;
  }
  if (depth != 8 || interlace != 0 || w == 0 || h == 0 || w > 65535 || h > 65535 || idat.empty()) return img;
  const size_t stride = (size_t)w * spp;
  std::vector<uint8_t> raw((stride + 1) * h);
  uLongf out_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) return img;
  std::vector<uint8_t> prev(stride, 0), cur(stride);
  const bool colour = ctype == 2 || ctype == 6 || ctype == 3;
  img.rows = (int)h, img.cols = (int)w, img.channels = colour ? 3 : 1;
  img.data.resize((size_t)h * w * img.channels);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t* line = &raw[(stride + 1) * y];
    const int filter = line[0];
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= (size_t)spp ? cur[x - spp] : 0, b = prev[x], c = x >= (size_t)spp ? prev[x - spp] : 0;
      int v = line[1 + x];
      switch (filter) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += detail::paeth(a, b, c); break;
        default: return Image();
      }
      cur[x] = (uint8_t)v;
    }
    uint8_t* dst = &img.data[(size_t)y * w * img.channels];
    for (uint32_t x = 0; x < w; ++x) {
      const uint8_t* s = &cur[(size_t)x * spp];
      if (ctype == 0 || ctype == 4) {
        dst[x] = s[0];
      } else if (ctype == 3) {
        if ((size_t)3 * s[0] + 2 >= pal.size()) return Image();
        dst[3 * x] = pal[3 * s[0] + 2], dst[3 * x + 1] = pal[3 * s[0] + 1], dst[3 * x + 2] = pal[3 * s[0]];
      } else {
        dst[3 * x] = s[2], dst[3 * x + 1] = s[1], dst[3 * x + 2] = s[0];
      }
    }
    prev.swap(cur);
  }
  return img;
}
#endif  // CTAG_WITH_ZLIB

// Returns an empty image on any error (like cv::imread).
inline Image imread(const std::string& path) {
  Image img;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return img;
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t n;
  while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  std::fclose(f);
#ifdef CTAG_WITH_ZLIB
  if (buf.size() >= 8 && buf[0] == 0x89 && buf[1] == 'P' && buf[2] == 'N' && buf[3] == 'G') return read_png(buf);
#endif
  if (buf.size() >= 54 && buf[0] == 'B' && buf[1] == 'M') {
    // every size below comes from the file: all arithmetic in 64 bits, every offset checked before it is used
    const uint64_t off = detail::rd32(&buf[10]), hdr = detail::rd32(&buf[14]);
    if (hdr < 40 || 14 + hdr > buf.size()) return img;
    const int32_t w = (int32_t)detail::rd32(&buf[18]), hs = (int32_t)detail::rd32(&buf[22]);
    const int bpp = detail::rd16(&buf[28]);
    const uint32_t comp = detail::rd32(&buf[30]);
    if (w <= 0 || w > 65535 || hs == 0 || hs == INT32_MIN || hs > 65535 || hs < -65535) return img;
    if ((comp != 0 && !(comp == 3 && bpp == 32)) || (bpp != 8 && bpp != 24 && bpp != 32)) return img;
    const int h = hs < 0 ? -hs : hs;
    const uint64_t stride = (((uint64_t)w * bpp + 31) / 32) * 4;
    if (off < 14 + hdr || off > buf.size() || stride * (uint64_t)h > buf.size() - off) return img;
    const uint8_t* pal = &buf[14 + hdr];
    uint8_t pal_full[256 * 4] = {0};  // indices beyond the stored palette read black instead of whatever follows in the file
    bool gray_palette = bpp == 8;
    if (bpp == 8) {
      uint32_t ncol = detail::rd32(&buf[46]);
      if (ncol == 0 || ncol > 256) ncol = 256;
      if (14 + hdr + 4 * (uint64_t)ncol > buf.size()) return img;
      std::memcpy(pal_full, pal, 4 * (size_t)ncol);
      pal = pal_full;
      for (uint32_t c = 0; c < ncol && gray_palette; ++c) gray_palette = pal[4 * c] == pal[4 * c + 1] && pal[4 * c] == pal[4 * c + 2];
    }
    img.rows = h, img.cols = w, img.channels = (bpp == 8 && gray_palette) ? 1 : 3;
    img.data.resize((size_t)h * w * img.channels);
    for (int y = 0; y < h; ++y) {
      const uint8_t* src = &buf[off + stride * (size_t)(hs < 0 ? y : h - 1 - y)];
      uint8_t* dst = &img.data[(size_t)y * w * img.channels];
      if (bpp == 8 && gray_palette) {
        for (int x = 0; x < w; ++x) dst[x] = pal[4 * src[x]];
      } else if (bpp == 8) {
        for (int x = 0; x < w; ++x) std::memcpy(dst + 3 * x, pal + 4 * src[x], 3);
      } else if (bpp == 24) {
        std::memcpy(dst, src, (size_t)3 * w);
      } else {
        for (int x = 0; x < w; ++x) std::memcpy(dst + 3 * x, src + 4 * x, 3);
      }
    }
    return img;
  }
  if (buf.size() > 2 && buf[0] == 'P' && (buf[1] == '5' || buf[1] == '6')) {
    size_t p = 2;
    int vals[3], got = 0;
    while (got < 3 && p < buf.size()) {
      while (p < buf.size() && (buf[p] == ' ' || buf[p] == '\n' || buf[p] == '\r' || buf[p] == '\t')) ++p;
      if (p < buf.size() && buf[p] == '#') {
        while (p < buf.size() && buf[p] != '\n') ++p;
        continue;
      }
      int v = 0, digits = 0;
      while (p < buf.size() && buf[p] >= '0' && buf[p] <= '9') v = v * 10 + (buf[p++] - '0'), ++digits;
      if (!digits) return img;
      vals[got++] = v;
    }
    ++p;  // single whitespace after maxval
    const int ch = buf[1] == '5' ? 1 : 3;
    if (got < 3 || vals[2] != 255 || p + (size_t)vals[0] * vals[1] * ch > buf.size()) return img;
    img.cols = vals[0], img.rows = vals[1], img.channels = ch;
    img.data.assign(buf.begin() + p, buf.begin() + p + (size_t)vals[0] * vals[1] * ch);
    if (ch == 3)  // PPM is R,G,B: swap to B,G,R
      for (size_t i = 0; i + 2 < img.data.size(); i += 3) {
        const uint8_t t = img.data[i];
        img.data[i] = img.data[i + 2];
        img.data[i + 2] = t;
      }
    return img;
  }
  return img;
}

// binary PPM (P6) / PGM (P5) writer for 3- / 1-channel 8-bit data, e.g. the drawAxis overlay
inline bool imwrite_pnm(const std::string& path, const uint8_t* data, int rows, int cols, int channels) {
  if (!data || rows <= 0 || cols <= 0 || (channels != 1 && channels != 3)) return false;
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  std::fprintf(f, "P%d\n%d %d\n255\n", channels == 3 ? 6 : 5, cols, rows);
  const size_t n = (size_t)rows * cols * channels;
  const bool ok = std::fwrite(data, 1, n, f) == n;
  return std::fclose(f) == 0 && ok;
}

// cv::cvtColor(COLOR_BGR2GRAY) for 8-bit data (main.cpp:36): (3735 B + 19235 G + 9798 R + 16384) >> 15.
// (The batched C entry point ctag_detect_batch(channels = 3) does this on the GPU; this host version is for the
//  single-image flow of the mirror class, whose detect() takes a gray image like the reference's.)
inline Image bgr2gray(const Image& bgr) {
  if (bgr.channels == 1) return bgr;
  Image g;
  g.rows = bgr.rows, g.cols = bgr.cols, g.channels = 1;
  g.data.resize((size_t)bgr.rows * bgr.cols);
  for (size_t i = 0; i < g.data.size(); ++i)
    g.data[i] = (uint8_t)((3735u * bgr.data[3 * i] + 19235u * bgr.data[3 * i + 1] + 9798u * bgr.data[3 * i + 2] + 16384u) >> 15);
  return g;
}

}  // namespace ctag_api
