// Host-callable launchers of the detection kernels (one translation unit per stage).
#pragma once
#include "common.cuh"

namespace ctag {

// K1 (front.cu): fused gray + 2x cubic decimation + adaptive threshold. `frames_dev` is u8, `channels` 1 or 3.
// `ev_start` (optional) is recorded on the stream right in front of the kernel, behind the host-side preparation
// (tensor maps), so that stage timings measure the kernel and not the host.
int launch_front(const void* frames_dev, int n, const FrameGeom& geo, int channels, size_t pitch, size_t frame_stride,
                 uint8_t* gray_out, size_t gray_fstride, uint8_t* bin_out, size_t bin_fstride, TileHint hint,
                 cudaStream_t stream, cudaEvent_t ev_start = nullptr);
int front_smem_bytes(int channels);

// front_generic.cu: the same stages for an adaptiveThresh window other than 5 (three plain kernels, half-res image in HBM).
// tile_buf needs 2 * n * ceil(hw/window) * ceil(hh/window) bytes, half_buf the size of the binary image.
int launch_front_generic(const void* frames_dev, int n, const FrameGeom& g, int channels, size_t pitch, size_t frame_stride,
                         int window, uint8_t* gray_out, size_t gray_fstride, uint8_t* half_buf, uint8_t* tile_buf,
                         uint8_t* bin_out, size_t bin_fstride, cudaStream_t stream, int* launches);

// K2/K3 (ccl.cu): block-based union-find labelling, stats, ordered legal-component list.
// legal[frame][k] = {root, area, x0, y0, x1, y1}; counters[frame] = {n_components, n_legal, overflow, 0}.
// `tile_any` (nullable, ccl_tile_hint_bytes(g) per frame): the front kernel's per-tile foreground flags (TileHint).
size_t ccl_seg_flag_bytes(const FrameGeom& g);
size_t ccl_tile_hint_bytes(const FrameGeom& g);
TileHint ccl_tile_hint(const FrameGeom& g, uint8_t* buf);
size_t ccl_roots_ints(const FrameGeom& g);                  // roots_tmp, per frame
size_t ccl_tile_list_ints(const FrameGeom& g, int frames);  // tile_list, per batch
int launch_ccl(const uint8_t* bin, size_t bin_fstride, int n, const FrameGeom& g, int* labels, int* st_area, int* st_x0,
               int* st_y0, int* st_x1, int* st_y1, int* roots_tmp, int* tile_list, uint8_t* seg_flags, const uint8_t* tile_any,
               int* legal, int legal_cap, int* counters, cudaStream_t stream, int* launches);

// K4 (quad.cu): edges (warp per component) -> Welsch fits (thread per restart, merge, exact fallback) -> corner
// selection -> ordered compaction.
size_t quad_scratch_bytes_per_warp(const FrameGeom& g);
size_t quad_fitrec_bytes();
size_t quad_tailrec_bytes();
int quad_tail_cap(int sms);
size_t quad_fitresult_bytes();
size_t quad_traj_bytes_per_cta();
int quad_edge_warps(int sms);
int quad_exact_ctas(int sms);
void quad_build_pick_table(uint16_t* host_table, int max_count);
int launch_quad(int n, const FrameGeom& g, const uint8_t* bin, size_t bin_fstride, const int* labels, const int* legal,
                int legal_cap, const int* counters, int* prefix, int* qctl, uint8_t* scratch, int edge_warps, void* fits,
                int fit_cap, int fit_per_frame, int* frame_fit, int* pool, int pool_cap, const uint16_t* pick_table, int table_max, void* results,
                int* exact_list, int* fit_order, void* tails, int tail_cap, void* traj, int exact_ctas, int sms, float* lines, int* quad_status,
                float* quad_corners, int quad_cap, float* quads, int* quad_comp, int* n_quads, cudaStream_t stream,
                int* launches);

// Compressed ingest (jpeg.cu): baseline JPEG with restart markers -> interleaved BGR frames.  jpeg_parse_frame returns
// 0, 1 (not a JPEG / corrupt) or 2 (outside the decoder's envelope: progressive, no restart markers, ...).
size_t jpeg_header_bytes();
int jpeg_parse_frame(const uint8_t* data, size_t len, void* header_out, size_t* scan_off, size_t* scan_len, int* w, int* h,
                     int* n_intervals, size_t* plane_bytes);
void jpeg_place_frame(void* header, uint32_t data_off, uint32_t data_len, int interval_first);
int jpeg_frame_blocks(const void* header);
int jpeg_frame_is_420(const void* header);
int launch_jpeg_decode(const void* d_hdr, int n, const uint8_t* d_bytes, uint32_t* d_ivl, int max_intervals, int16_t* d_coefs,
                       size_t coef_stride, uint8_t* d_last, size_t last_stride, int max_blocks, uint8_t* d_planes, size_t plane_stride,
                       uint8_t* d_bgr, size_t pitch, size_t frame_stride, int w, int h, int all_420, int* d_status, cudaStream_t stream,
                       int* launches);

// Synthetic frames on the GPU (render.cu).
size_t render_marker_dev_bytes();
void render_pack_marker(const float* spec, int cols, int row_off, const float* K, int w, int h, void* out);
int launch_render(float* d_img, int n, int w, int h, const float* d_fparams, const void* d_markers, const int* d_marker_start,
                  const int* d_states, uint8_t* out, size_t pitch, size_t frame_stride, int channels, cudaStream_t stream);

// K5/K6 (feature.cu): quad pairing, coordinate lift, edge refinement.  fstate[frame] = {status, n_features,
// n_features going on, overflow}.
size_t sizeof_quad_geom();
size_t sizeof_feature_rec();
int launch_features(int n, const FrameGeom& g, const float* quads, const int* n_quads, int quad_cap, void* geom, void* feats,
                    int feat_cap, int feature_size, int* fstate, const uint8_t* gray, size_t gray_pitch, size_t gray_fstride,
                    int corner_subpix, int subpix_dist, cudaStream_t stream, int* launches);

// K7 (feature.cu): grouping, cross-ratio IDs, dictionary decode; packs the markers of the batch and writes
// summary[frame][12] = {status, n_labels, n_legal, n_quads, n_features, n_groups, n_markers, flagged, stale, offset, stored}.
size_t decode_smem_bytes(int srows, int scols);
int launch_decode(int n, const void* feats, int feat_cap, const int* fstate, const int* state, int srows, int scols, int fsz,
                  ctag_marker* markers, int marker_cap, const int* counters, const int* n_quads, int quad_cap,
                  const int* batch_overflow, const int* frame_overflow, ctag_marker* packed, int* packed_count, int* summary, cudaStream_t stream,
                  int* launches);

}  // namespace ctag
