/*
 * ctag.h -- C ABI of the B200-native CylinderTag detection front end.
 *
 * This is the drop-in boundary for the reference's per-frame detect path
 * (wsakobe/CylinderTag).  Every entry point names the reference interface it
 * replaces (file:line in the reference tree).  Plain pointers and sizes only;
 * no C++/torch/OpenCV types cross this boundary.  The C++ class
 * `CylinderTag` in include/cylindertag/CylinderTag.h and the Python class
 * `cylindertag_b200.CylinderTag` are thin wrappers over these calls.
 *
 * All compute runs in hand-written sm_100a CUDA kernels; there is no CPU
 * fallback: if no CUDA device is usable every call fails with
 * CTAG_ERR_CUDA / CTAG_ERR_NO_DEVICE.
 *
 * Threading: like the reference object (scratch members,
 * header/corner_detector.h:61-156) a ctag_detector is NOT re-entrant.  Use
 * one detector per host thread / CUDA device; calls on one handle are
 * serialised by the caller.
 */
#ifndef CTAG_H
#define CTAG_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CTAG_API __attribute__((visibility("default")))
#else
#define CTAG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define CTAG_MAX_FEATURES 20 /* reference: code[20], header/corner_detector.h:152 */
#define CTAG_MAX_FRAME_FEATURES 100 /* reference: father[100], header/corner_detector.h:143 */
#define CTAG_MAX_FRAME_QUADS 1000 /* reference: isVisited[1000], header/corner_detector.h:124 */

/* Return codes. 0 = OK (zero markers is not an error, CylinderTag.cpp:87-96). */
enum {
  CTAG_OK = 0,
  CTAG_ERR_ARG = -1,         /* bad argument (null pointer, odd size, ...) */
  CTAG_ERR_FILE = -2,        /* replaces `throw "... could not open the file"` (CylinderTag.cpp:21,165) */
  CTAG_ERR_DICTIONARY = -3,  /* replaces `throw "... must between 0 to 63"` (CylinderTag.cpp:56-65) */
  CTAG_ERR_CUDA = -4,        /* CUDA runtime/driver failure; see ctag_last_error() */
  CTAG_ERR_NO_DEVICE = -5,   /* no usable sm_100 device: there is NO CPU fallback */
  CTAG_ERR_UNSUPPORTED = -6, /* configuration outside the implemented envelope (half-res side > 4095, dictionary > 200 KB) */
  CTAG_ERR_CAPACITY = -7,    /* an internal device list overflowed its capacity */
  CTAG_ERR_ALIGNMENT = -8    /* device input not 16-byte aligned / pitch not a multiple of 16 */
};

/* Per-frame status (reference early exits, CylinderTag.cpp:87-96). */
enum {
  CTAG_FRAME_OK = 0,
  CTAG_FRAME_NO_CORNER = 1,  /* "No corner detected!"  -> reference leaves the output untouched */
  CTAG_FRAME_NO_FEATURE = 2  /* "No feature detected!" -> reference leaves the output untouched */
};

typedef struct ctag_detector ctag_detector;

/* POD image of the reference's MarkerInfo (header/corner_detector.h:16-22), one decoded marker. */
typedef struct ctag_marker {
  int32_t marker_id;  /* MarkerInfo::markerID = dictionary row */
  int32_t n_features; /* cornerLists.size() */
  int32_t inverse;    /* pos_with_ID::inverse; corner halves already swapped (corner_detector.cpp:1239-1246) */
  int32_t frame;      /* index of the frame inside the batch */
  int32_t feature_pos[CTAG_MAX_FEATURES];
  int32_t feature_id[CTAG_MAX_FEATURES];
  int32_t id_left[CTAG_MAX_FEATURES];
  int32_t id_right[CTAG_MAX_FEATURES];
  float cr_left[CTAG_MAX_FEATURES];
  float cr_right[CTAG_MAX_FEATURES];
  float edge_length[CTAG_MAX_FEATURES];
  float center[CTAG_MAX_FEATURES][2];
  float corners[CTAG_MAX_FEATURES][8][2]; /* full-resolution pixel coordinates, 8-corner order of SURVEY D.4 */
} ctag_marker;

/* Per-frame counters (the reference has no metrics; SURVEY 5 asks for per-stage counters). */
typedef struct ctag_frame_info {
  int32_t status;      /* CTAG_FRAME_* */
  int32_t n_labels;    /* connected components incl. background (connectedComponentsWithStats return value) */
  int32_t n_legal;     /* components passing the area test (corner_detector.cpp:88) */
  int32_t n_quads;     /* corners_init.size() after edgeExtraction */
  int32_t n_features;  /* features.size() after featureRecovery */
  int32_t n_groups;    /* markers.size() after markerOrganization */
  int32_t n_markers;   /* decoded markers */
  int32_t flagged;     /* 1 if a reference fixed-array limit was exceeded (SURVEY C-4): frame excluded from parity */
  int32_t stale_ids;   /* cross-ratio band misses that reuse the previous ID (SURVEY C-2) */
  int32_t reserved[3];
} ctag_frame_info;

/* ---- construction -------------------------------------------------------------------------- */

/* Replaces CylinderTag::CylinderTag(const Mat1i&) (CylinderTag.cpp:11-14,43-54).  `state` is rows x cols,
 * row-major, every entry in [0,63].  `feature_size` must be given (the reference leaves it unset, SURVEY C-3).
 * `cuda_device` < 0 selects the current device. */
CTAG_API int ctag_create(ctag_detector** out, const int32_t* state, int rows, int cols, int feature_size, int cuda_device);

/* Replaces CylinderTag::CylinderTag(const string&) / load_from_file (CylinderTag.cpp:6-9,16-41). */
CTAG_API int ctag_create_from_file(ctag_detector** out, const char* marker_path, int cuda_device);

CTAG_API void ctag_destroy(ctag_detector* det);

/* Tuning knobs of one detector (the defaults are the measured best; environment variables of the same meaning are
 * read ONCE, when the detector is created).  Keys: "chunk_frames" -- frames per chunk of a host batch in
 * ctag_detect_batch(is_device=0), 0 = automatic (about 192 MiB of frames, at most a quarter of the batch).
 * "jpeg_decoder" -- 0 (default): ctag_detect_batch_jpeg uses the library's CUDA decoder and falls back to nvJPEG for
 * batches it refuses, 1: nvJPEG only.
 * "debug_fail_chunk" -- fault injection for the tests: the next host batch fails with CTAG_ERR_CUDA when it reaches
 * that chunk index (one shot; -1 = off).  Not allowed while batches are pending.  Unknown key: CTAG_ERR_ARG. */
CTAG_API int ctag_set_option(ctag_detector* det, const char* key, int value);

/* Dictionary accessors (state matrix the detector holds, header/CylinderTag.h:44-45). */
CTAG_API int ctag_get_dictionary(const ctag_detector* det, int* rows, int* cols, int* feature_size, int32_t* state_out, int cap);

/* ---- detection ----------------------------------------------------------------------------- */

/* Replaces CylinderTag::detect(img, markers, adaptiveThresh, cornerSubPix, cornerSubPixDist)
 * (CylinderTag.cpp:67-128) for one 8-bit single-channel HOST image (any pitch), synchronous.
 * Writes at most `cap` markers and the true count to *n_out.  *frame_status (optional) receives CTAG_FRAME_*:
 * on the two reference early exits the C++ wrapper leaves the caller's vector untouched. */
CTAG_API int ctag_detect(ctag_detector* det, const uint8_t* gray, int w, int h, size_t pitch, int adaptive_thresh,
                int corner_subpix, int subpix_dist, ctag_marker* out, int cap, int* n_out, int* frame_status);

/* Batched form of the same path: `n` independent frames (the reference's video loop main.cpp:52-60).
 * frames      : host (is_device=0) or device (is_device=1) pointer to frame 0
 * channels    : 1 = gray (detect's contract), 3 = BGR interleaved (then the caller's cvtColor, main.cpp:54, is fused in)
 * pitch       : bytes between rows; frame_stride: bytes between frames (0 -> pitch*h)
 * out         : host array of n*cap_per_frame markers, frame f's markers start at out[f*cap_per_frame]
 * n_out       : host array of n counts; info (optional): host array of n ctag_frame_info
 * Device inputs must be 16-byte aligned with pitch % 16 == 0 (TMA requirement).
 * w and h must be even: the reference's resize (CylinderTag.cpp:79) is an exact 2x bicubic decimation only then; for odd
 * sizes cv::resize's own output depends on the library's CPU dispatch (tests/test_ref_pinning.py), so there is no single
 * reference answer to be bit-exact with -- such frames are refused with CTAG_ERR_ARG.
 * A call that fails half way (out of memory, a failed copy) leaves the detector usable: pending work is drained and the
 * next call starts clean.  out, n_out and info may each be NULL. */
CTAG_API int ctag_detect_batch(ctag_detector* det, const void* frames, int n, int w, int h, size_t pitch, size_t frame_stride,
                      int channels, int is_device, int adaptive_thresh, int corner_subpix, int subpix_dist,
                      ctag_marker* out, int cap_per_frame, int* n_out, ctag_frame_info* info);

/* The same call sharded over several detectors -- one per CUDA device, created with ctag_create(..., cuda_device = g) --
 * for the video loop of main.cpp:52-60 on a multi-GPU box.  Frames are independent (detect() clears its
 * state per call, CylinderTag.cpp:73-76), so the batch is cut into contiguous blocks, frame f -> detector
 * floor(f * n_det / n); each block runs ctag_detect_batch(is_device = 0) on its own host thread and there is no exchange
 * between devices.  `frames` is a HOST pointer; out / n_out / info are indexed by the global frame number and
 * ctag_marker::frame is the global index, i.e. the result equals the single-detector call on the whole batch.
 * n_out is required when out is given.  The detectors must be distinct objects (CTAG_ERR_ARG otherwise: a detector is not
 * re-entrant).  Returns the first non-OK block status; ctag_last_error() then holds that block's message. */
CTAG_API int ctag_detect_batch_multi(ctag_detector* const* dets, int n_det, const void* frames, int n, int w, int h,
                                     size_t pitch, size_t frame_stride, int channels, int adaptive_thresh, int corner_subpix,
                                     int subpix_dist, ctag_marker* out, int cap_per_frame, int* n_out, ctag_frame_info* info);

/* Compressed ingest: the same batch call on JPEG byte strings (host memory), replacing the reference's
 * decode-on-the-CPU front (cv::VideoCapture::read / cv::imread, main.cpp:29,45-52, then cvtColor, :54).  The frames are
 * decoded on the GPU straight into the interleaved BGR layout the fused front kernel reads, chunk by chunk on the
 * pipeline's streams, so only the compressed bytes cross PCIe.
 *   Decoder 2, the library's own CUDA decoder (csrc/jpeg.cu): baseline / sequential JPEG, 8 bit, gray or YCbCr 4:4:4 /
 *   4:2:2 / 4:2:0, WITH restart markers (DRI): a restart interval is the unit that decodes independently, one GPU thread
 *   each.  Its pixels are those of cv::imdecode (libjpeg-turbo defaults: integer IDCT, fancy upsampling), byte for byte.
 *   Decoders 0 / 1, nvJPEG (loaded with dlopen on first use): for batches the CUDA decoder refuses (no restart markers,
 *   progressive): the NVJPG hardware engine when nvJPEG offers it on the device, else nvJPEG's hybrid decoder.
 * ctag_set_option("jpeg_decoder", 1) forces nvJPEG.  All frames of a batch must have the same (even) size; *width_out /
 * *height_out (optional) report it.  There is no CPU decoder behind this call: when neither GPU decoder can take the
 * batch it fails with CTAG_ERR_UNSUPPORTED.  Parity is defined on the decoded pixels: ctag_debug_get_input copies them
 * back. */
CTAG_API int ctag_detect_batch_jpeg(ctag_detector* det, const uint8_t* const* jpeg, const size_t* jpeg_bytes, int n,
                                    int adaptive_thresh, int corner_subpix, int subpix_dist, ctag_marker* out, int cap_per_frame,
                                    int* n_out, ctag_frame_info* info, int* width_out, int* height_out);
/* Decoder used by the most recent compressed batch: 2 = the library's CUDA decoder, 0 = nvJPEG on the NVJPG hardware
 * engine, 1 = nvJPEG hybrid backend, -1 = none yet. */
CTAG_API int ctag_jpeg_backend(const ctag_detector* det);

/* Asynchronous pair for device-resident throughput runs: enqueue the whole detect path for a batch, then collect.
 * Up to ctag_max_in_flight() batches may be enqueued before the first collect (each has its own workspace and CUDA
 * stream, so the latency-bound sparse kernels of one batch overlap the dense kernels of the next); collect returns
 * them in FIFO order.  Between the calls the host is free (e.g. to upload the next batch). */
CTAG_API int ctag_detect_batch_enqueue(ctag_detector* det, const void* frames_dev, int n, int w, int h, size_t pitch,
                              size_t frame_stride, int channels, int adaptive_thresh, int corner_subpix, int subpix_dist);
CTAG_API int ctag_detect_batch_collect(ctag_detector* det, ctag_marker* out, int cap_per_frame, int* n_out, ctag_frame_info* info);
CTAG_API int ctag_max_in_flight(void);

/* ---- pose stage (host code, no GPU work) --------------------------------------- */

/* Replaces PoseEstimator::PnPSolver + PoseBA (pose_estimation.cpp:50-143; called from CylinderTag::estimatePose,
 * CylinderTag.cpp:198-209) for one marker: corner selection (pose_estimation.cpp:72-95), 5-coefficient undistortion
 * (:97-101), EPnP initial pose (:103) and Levenberg-Marquardt on the pinhole reprojection residual (:14-41,105-128).
 *   model_corners  [n_model_corners][3] floats of the marker's reconstructed model, index = featurePos * 8 + k
 *                  (.model layout, CylinderTag.cpp:168-188)
 *   intrinsic      3x3 row-major camera matrix, dist: k1 k2 p1 p2 k3 (cameraParams.yml, both dt: f)
 *   rvec, tvec     3 doubles each: x_cam = R(rvec) x_model + tvec;  rms_px (optional): reprojection RMS in pixels
 * Returns CTAG_ERR_ARG when fewer than 4 corners qualify or a corner has no model point. */
CTAG_API int ctag_estimate_pose(const ctag_marker* marker, const float* model_corners, int n_model_corners,
                                const float* intrinsic, const float* dist, int n_dist, double* rvec, double* tvec,
                                double* rms_px);
/* The corner selection alone: fills (feature index, corner index 0..7) pairs, returns their number. */
CTAG_API int ctag_pose_select_points(const ctag_marker* marker, int* feature_of_point, int* corner_of_point, int cap);

/* ---- drawAxis overlay (host code, no GPU work) --------------------------------- */

/* cv::projectPoints as called at CylinderTag.cpp:234: points3 [n][3] model points -> out_xy [n][2] pixels through
 * R(rvec) X + tvec, the pinhole division, the k1 k2 p1 p2 k3 lens model and the 3x3 row-major camera matrix. */
CTAG_API int ctag_project_points(const float* points3, int n, const double* rvec, const double* tvec, const float* intrinsic,
                                 const float* dist, int n_dist, float* out_xy);
/* cvtColor(img, imgMark, COLOR_GRAY2RGB) (CylinderTag.cpp:214): gray w x h -> three equal channels, interleaved. */
CTAG_API int ctag_gray_to_3ch(const uint8_t* gray, int w, int h, size_t pitch, uint8_t* out3, size_t out_pitch);
/* The body of CylinderTag::drawAxis's loop over poses (CylinderTag.cpp:219-243) for one (marker, pose) pair, drawn
 * into a caller-owned 3-channel 8-bit image instead of an imshow window: filled circles (r = 5, colour 255,234,32) on
 * the projected model corners of the marker's features, arrows (thickness 10, tip 0.2; colours 255,0,0 / 0,255,0 /
 * 0,0,255 in channel order) from the projected base point along the model axis and the two directions the reference
 * hard-codes, scaled by axis_length, and a filled circle (r = 8, colour 247,235,235) on the base point.
 *   model_corners [n_model_corners][3], base[3], axis[3]: the ModelInfo of the pose (.model, CylinderTag.cpp:168-188) */
CTAG_API int ctag_draw_axis(uint8_t* img3, int w, int h, size_t pitch, const ctag_marker* marker, const float* model_corners,
                            int n_model_corners, const float* base, const float* axis, const float* intrinsic,
                            const float* dist, int n_dist, const double* rvec, const double* tvec, int axis_length);

/* ---- dictionary generation (host code, no GPU work) ---------------------------- */

/* The reference ships one dictionary; CylinderTag_generator.m (MATLAB) searches others.  ctag_generate_codebook is that
 * search (:34-216): rows are grown state by state, candidates ordered by the continuations they leave, with backtracking,
 * every cyclic window of feature_size states unique over the book read forwards and as its inverse (:247-286).
 * ctag_codebook_capacity = legal windows that differ from their inverse / (2 * cols) (:36-39): 41 for 2-state windows on
 * 12 columns, the size of the shipped CTag_2f12c.marker, which the search reaches.  `rows` is clamped to the capacity;
 * *rows_out rows were found (state_out: rows x cols, row-major, room for cap_rows rows).  feature_size 2..4. */
CTAG_API int ctag_codebook_capacity(int cols, int feature_size);
CTAG_API int ctag_generate_codebook(int cols, int feature_size, int rows, uint64_t seed, int32_t* state_out, int cap_rows, int* rows_out);
/* 1 if every state is legal and every window reading is unique (testConflict, :247-286), 0 if not, < 0 on bad arguments */
CTAG_API int ctag_check_codebook(const int32_t* state, int rows, int cols, int feature_size);

/* Synthetic camera frames rendered on the GPU, straight into device memory in the layout ctag_detect_batch_enqueue takes
 * (CylinderTag_generator.m:206-245 only draws flat marker bitmaps).  Scene model of SURVEY Appendix D.3 / D.6:
 * each marker is a row of the detector's dictionary wrapped once around a cylinder (column width W = 2 pi r / (1.5 cols),
 * height L = ratio * W), pinhole camera looking along +Z, x_cam = R(rvec) x_obj + tvec, 3 x 3 rays per pixel, smooth
 * background in [140, 220], Gaussian blur, Gaussian noise, 8-bit rounding; channels = 3 adds +-6 levels of chroma noise.
 *   marker_specs  [n_markers][16] floats: rvec(3) tvec(3) radius ratio black white dictionary_row, rest unused;
 *                 markers of frame f are marker_start[f] .. marker_start[f + 1] (far markers first: later ones occlude)
 *   frame_params  [n][8] floats: fx fy cx cy blur_sigma noise_sigma noise_seed(bit pattern) background_seed(bit pattern)
 * Synchronous.  Not bit-compatible with the host renderer of cylindertag_b200/synth.py. */
CTAG_API int ctag_render_frames(ctag_detector* det, void* frames_dev, int n, int w, int h, size_t pitch, size_t frame_stride,
                                int channels, const float* marker_specs, const int* marker_start, const float* frame_params);

/* ---- instrumentation ------------------------------------------------------------------------ */

/* Stage identifiers for ctag_stage_time_ms / ctag_debug_*. */
enum {
  CTAG_STAGE_FRONT = 0,   /* a0-a3: gray + 2x cubic decimation + tile min/max + threshold (fused) */
  CTAG_STAGE_CCL = 1,     /* a4: block-based union-find labelling + stats + ordered compaction */
  CTAG_STAGE_QUAD = 2,    /* a5: per-component boundary trace, RDP, line fits, quad selection */
  CTAG_STAGE_FEATURE = 3, /* a6-a8: pairing, coordinate lift, edge refinement */
  CTAG_STAGE_DECODE = 4,  /* a9-a10: grouping, cross-ratio IDs, dictionary match */
  CTAG_STAGE_COUNT = 5
};

/* GPU time (ms, CUDA events on the detector's stream) of each stage in the most recent collected batch. */
CTAG_API int ctag_stage_time_ms(const ctag_detector* det, float* ms_out /* CTAG_STAGE_COUNT */);
/* Start of every stage and end of the last one (ms since the detector was created, same events) for the most recent
 * collected batch: with several batches in flight these show how the stages of different batches overlap. */
CTAG_API int ctag_stage_timeline_ms(const ctag_detector* det, float* ms_out /* CTAG_STAGE_COUNT + 1 */);
/* Number of kernel launches issued by the most recent batch. */
CTAG_API int ctag_last_launch_count(const ctag_detector* det);
/* CUDA stream (cudaStream_t) the detector launches on, for callers that time with their own events. */
CTAG_API void* ctag_stream(const ctag_detector* det);

/* Stage dumps of the most recent batch for parity tests (copied to host buffers). */
CTAG_API int ctag_debug_get_gray(ctag_detector* det, int frame, uint8_t* out, size_t out_pitch);     /* w x h */
/* the staged input of a HOST or JPEG batch as the kernels saw it: w x h x channels, interleaved (last chunk only) */
CTAG_API int ctag_debug_get_input(ctag_detector* det, int frame, uint8_t* out, size_t out_pitch);
CTAG_API int ctag_debug_get_binary(ctag_detector* det, int frame, uint8_t* out, size_t out_pitch);   /* (w/2) x (h/2), {0,255} */
/* control words of the quad stage of the most recent batch (16 ints): [2] components that reached four edges, [1] edges
 * that took the exact sub-EPS bookkeeping path, [6] restarts fitted one per warp from the start (20 per edge of a large
 * cluster), [7] restarts the lane-per-restart kernel parked for the warp kernel at its tail, [10] reweighting passes that
 * fell back from the parallel to the sequential summation */
CTAG_API int ctag_debug_get_quad_counters(ctag_detector* det, int32_t* out16);
/* legal components in reference order: 6 ints each {root_block, area, x0, y0, x1, y1} (inclusive bbox, half-res) */
CTAG_API int ctag_debug_get_components(ctag_detector* det, int frame, int32_t* out, int cap, int* n_out);
/* quads in component order: comp index + 4 corners (half-res coords) */
CTAG_API int ctag_debug_get_quads(ctag_detector* det, int frame, int32_t* comp_index, float* corners /* [cap][4][2] */, int cap,
                         int* n_out);
/* features after refinement: 8 corners, centre, angle, source quad indices */
CTAG_API int ctag_debug_get_features(ctag_detector* det, int frame, float* corners /* [cap][8][2] */, float* center /* [cap][2] */,
                            float* angle, int32_t* quad_pair /* [cap][2] */, int cap, int* n_out);

CTAG_API const char* ctag_strerror(int code);
/* Text of the last CUDA error seen by this thread (empty string if none). */
CTAG_API const char* ctag_last_error(void);
/* Library/build identification, e.g. "cylindertag_b200 0.1 sm_100a". */
CTAG_API const char* ctag_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CTAG_H */
