// Host-callable launchers of the detection kernels (one translation unit per stage).
#pragma once
#include "common.cuh"

namespace ctag {

// K1 (front.cu): fused gray + 2x cubic decimation + adaptive threshold. `frames_dev` is u8, `channels` 1 or 3.
int launch_front(const void* frames_dev, int n, const FrameGeom& geo, int channels, size_t pitch, size_t frame_stride,
                 uint8_t* gray_out, size_t gray_fstride, uint8_t* bin_out, size_t bin_fstride, cudaStream_t stream);
int front_smem_bytes(int channels);

}  // namespace ctag
