// Shared definitions for the sm_100a CylinderTag detection kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ctag.h"

namespace ctag {

// ---- frozen constants of the reference (header/corner_detector.h, corner_detector.cpp) ----
constexpr int kWin = 5;               // adaptiveThresh the fused front end is specialised for (main.cpp:39,57)
constexpr int kAreaMin = 30;          // corner_detector.cpp:88
constexpr float kThrExpand = 1.2f;    // corner_detector.h:90 threshold_expand
constexpr float kThrLine = 1.8f;      // corner_detector.h:90 threshold_line
constexpr float kThrRAC = 0.3f;       // corner_detector.h:110
constexpr float kThrAngle = 5.0f;     // corner_detector.h:122
constexpr double kPi = 3.1415926535897932384626433832795;  // CV_PI

// ---- geometry of one frame ----
struct FrameGeom {
  int w, h;        // full resolution
  int hw, hh;      // half resolution (w/2, h/2)
  int gpitch;      // gray pitch (bytes), multiple of 16
  int bpitch;      // binary pitch (bytes), multiple of 16
  int cn, rn;      // threshold tiles (ceil(hw/5), ceil(hh/5))
  int bw, bh;      // 2x2 CCL blocks (ceil(hw/2), ceil(hh/2))
  int nblocks;     // bw*bh
  int area_max;    // round(0.01*hw*hh)
};

// Which 64x64-pixel tiles of the binary image (= 32x32 CCL blocks) hold foreground: any[frame*fstride + ty*pitch + tx],
// zeroed per batch, set by the kernel that writes the binary image.  any == nullptr: not tracked.
struct TileHint {
  uint8_t* any;
  int pitch, fstride;
};

#define CTAG_CUDA_CHECK(expr)                                  \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) {                                   \
      ::ctag::set_last_error(#expr, _e, __FILE__, __LINE__);   \
      return CTAG_ERR_CUDA;                                    \
    }                                                          \
  } while (0)

void set_last_error(const char* what, cudaError_t e, const char* file, int line);
void set_last_error_text(const char* text);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

}  // namespace ctag
