// C ABI (include/ctag.h) over the sm_100a detection kernels: detector handle, workspaces, batch pipeline.
//
// A detector owns kSlots independent workspaces ("slots"), each with its own CUDA stream, so that several batches can
// be in flight: the latency-bound sparse kernels of batch i overlap the bandwidth-bound dense kernels of batch i+1
// (ctag_detect_batch_enqueue may be called up to ctag_max_in_flight() times before ctag_detect_batch_collect; results
// come back in FIFO order).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "nvjpeg_dl.h"

namespace ctag {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* what, cudaError_t e, const char* file, int line) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
}
void set_last_error_text(const char* text) { snprintf(g_last_error, sizeof(g_last_error), "%s", text); }

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

static FrameGeom make_geom(int w, int h) {
  FrameGeom g;
  g.w = w;
  g.h = h;
  g.hw = w / 2;
  g.hh = h / 2;
  g.gpitch = round_up(w, 16);
  g.bpitch = round_up(g.hw, 16);
  g.cn = (g.hw + kWin - 1) / kWin;
  g.rn = (g.hh + kWin - 1) / kWin;
  g.bw = (g.hw + 1) / 2;
  g.bh = (g.hh + 1) / 2;
  g.nblocks = g.bw * g.bh;
  g.area_max = (int)round(0.01 * g.hw * g.hh);  // corner_detector.cpp:88 (C round, half away from zero)
  return g;
}

constexpr int kMaxSlots = 8;
// batches in flight per detector: 6 unless CTAG_SLOTS (1..8) says otherwise (read once per process).  Measured on
// 64-frame 4K batches: 4 -> 1.41 ms per batch, 6 -> 1.39, 8 -> 1.38; a workspace is only allocated when its slot is used.
static int slot_count() {
  static int n = 0;
  if (!n) {
    const char* e = getenv("CTAG_SLOTS");
    int v = e ? atoi(e) : 6;
    n = v < 1 ? 1 : (v > kMaxSlots ? kMaxSlots : v);
  }
  return n;
}
#define kSlots (slot_count())
constexpr int kQuadCap = CTAG_MAX_FRAME_QUADS;
constexpr int kFeatCap = CTAG_MAX_FRAME_FEATURES;
constexpr int kMarkerCap = CTAG_MAX_FRAME_FEATURES / 2;
constexpr int kPickTableMax = 1024;

// Everything one batch needs on the device, plus its stream, events and pinned result buffers.
struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[CTAG_STAGE_COUNT + 1] = {};
  int cap_frames = 0, w = 0, h = 0;
  FrameGeom geo{};
  uint8_t* d_stage = nullptr;  // staging for host inputs
  size_t stage_bytes = 0;
  std::vector<void*> owned;  // device allocations of the workspace (freed together)
  uint8_t *d_gray = nullptr, *d_bin = nullptr, *d_quad_scratch = nullptr;
  uint8_t* d_tile_any = nullptr;   // front kernel -> CCL: which 64x64 tiles of the binary image hold foreground
  uint8_t* d_seg_flags = nullptr;  // CCL: which 32-block label segments exist (ccl.cu)
  uint8_t *d_half = nullptr, *d_tiles = nullptr;  // generic-window front end only (allocated on first use)
  size_t tiles_bytes = 0;
  size_t gray_fstride = 0, bin_fstride = 0;
  int *d_labels = nullptr, *d_st_area = nullptr, *d_st_x0 = nullptr, *d_st_y0 = nullptr, *d_st_x1 = nullptr,
      *d_st_y1 = nullptr, *d_roots_tmp = nullptr, *d_tile_list = nullptr, *d_legal = nullptr, *d_counters = nullptr;
  int legal_cap = 0;
  int *d_prefix = nullptr, *d_qctl = nullptr, *d_quad_status = nullptr, *d_quad_comp = nullptr, *d_n_quads = nullptr,
      *d_exact_list = nullptr, *d_fit_order = nullptr, *d_frame_fit = nullptr, *d_pool = nullptr;
  float *d_quad_corners = nullptr, *d_quads = nullptr, *d_lines = nullptr;
  void *d_tails = nullptr;
  int tail_cap = 0;
  void *d_fits = nullptr, *d_traj = nullptr, *d_fit_results = nullptr, *d_geom = nullptr, *d_feats = nullptr;
  int edge_warps = 0, exact_ctas = 0, fit_cap = 0, fit_per_frame = 0, pool_cap = 0;
  int *d_fstate = nullptr, *d_packed_count = nullptr, *d_summary = nullptr;
  ctag_marker *d_markers = nullptr, *d_packed = nullptr;
  int* h_summary = nullptr;         // pinned
  ctag_marker* h_packed = nullptr;  // pinned
  // the batch this slot holds
  int n = 0, channels = 0, subpix = 0, launches = 0;
  const uint8_t* gray = nullptr;  // full-res gray of the batch (the input itself when channels == 1)
  size_t gray_pitch = 0, gray_fs = 0;
  bool busy = false;
  // compressed ingest (ctag_detect_batch_jpeg): one batched-decode state per backend, sized for jpeg_batch[] frames
  nvjpegJpegState_t jpeg_state[2] = {nullptr, nullptr};
  int jpeg_batch[2] = {0, 0};
  // the CUDA JPEG decoder's buffers (jpeg.cu), grown on demand
  uint8_t *d_jbytes = nullptr, *d_jplanes = nullptr, *d_jhdr = nullptr, *h_jhdr = nullptr, *d_jlast = nullptr;
  uint32_t* d_jivl = nullptr;
  int16_t* d_jcoefs = nullptr;
  int *d_jstatus = nullptr, *h_jstatus = nullptr;
  size_t jbytes_cap = 0, jplanes_cap = 0, jivl_cap = 0, jcoefs_cap = 0, jlast_cap = 0;
  int w_expect = 0, h_expect = 0;  // frame size of the compressed batch being staged
  int jhdr_cap = 0, jpeg_frames = 0;  // jpeg_frames > 0: the batch in this slot came through the CUDA decoder
};

}  // namespace ctag

using namespace ctag;

struct ctag_detector {
  int device = 0, sms = 0;
  // dictionary (header/CylinderTag.h:44-45)
  std::vector<int32_t> state;
  int rows = 0, cols = 0, feature_size = 0;
  int32_t* d_state = nullptr;
  uint16_t* d_pick_table = nullptr;  // cv::fitLine restart subsets per point count (fit_core.cuh)
  Slot slot[kMaxSlots];
  cudaEvent_t epoch = nullptr;  // recorded once at creation: origin of ctag_stage_timeline_ms
  float stage_stamp_ms[CTAG_STAGE_COUNT + 1] = {};
  int next_enqueue = 0, next_collect = 0, in_flight = 0;
  int last = -1;  // slot of the most recently collected batch (debug getters, stage times)
  float stage_ms[CTAG_STAGE_COUNT] = {0, 0, 0, 0, 0};
  int last_launches = 0;
  nvjpegHandle_t jpeg_handle[2] = {nullptr, nullptr};  // [0] hardware engine (NVJPG), [1] default (CUDA) backend
  int jpeg_tried[2] = {0, 0};
  int jpeg_backend_used = -1;  // decoder of the most recent compressed batch: 2 = CUDA decoder (jpeg.cu), 0 / 1 = nvJPEG
  int jpeg_decoder = 0;  // ctag_set_option("jpeg_decoder"): 0 = CUDA decoder first, nvJPEG for what it refuses; 1 = nvJPEG only
  int debug_fail_chunk = -1;  // fault injection for the tests: fail the host pipeline when this chunk index is reached
  int chunk_frames = 0;  // frames per chunk of a host batch; 0 = automatic (CTAG_CHUNK at creation, ctag_set_option later)
};

// Brings the slot queue back to "nothing in flight" after a failure in the middle of a pipelined call: waits for
// whatever was queued on the slot streams (the staging buffers they read stay alive), then resets the ring.  Without
// this a transient error (out of memory while growing a staging buffer, a failed copy) would leave in_flight > 0 and
// every later call on the handle would be refused.
static void drain_slots(ctag_detector* d) {
  for (int i = 0; i < kMaxSlots; ++i) {
    Slot& s = d->slot[i];
    if (s.stream) cudaStreamSynchronize(s.stream);
    s.busy = false;
  }
  cudaGetLastError();  // clear a non-sticky error so that the next call starts clean
  d->next_enqueue = d->next_collect = 0;
  d->in_flight = 0;
}

static int select_device(int cuda_device, int* chosen, int* sms) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_last_error_text("no CUDA device available (the detection path has no CPU fallback)");
    return CTAG_ERR_NO_DEVICE;
  }
  int dev = cuda_device;
  if (dev < 0) CTAG_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev >= count) return CTAG_ERR_ARG;
  cudaDeviceProp prop;
  CTAG_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    snprintf(g_last_error, sizeof(g_last_error), "device %d is sm_%d%d; this library carries sm_100a code only", dev,
             prop.major, prop.minor);
    return CTAG_ERR_NO_DEVICE;
  }
  *chosen = dev;
  *sms = prop.multiProcessorCount;
  return CTAG_OK;
}

static void free_workspace(Slot* s) {
  // d_stage is managed separately (ensure_stage): it may hold the batch that is about to be processed
  for (void* p : s->owned) cudaFree(p);
  s->owned.clear();
  s->d_half = s->d_tiles = nullptr;
  s->tiles_bytes = 0;
  cudaFreeHost(s->h_summary);
  cudaFreeHost(s->h_packed);
  s->h_summary = nullptr;
  s->h_packed = nullptr;
  s->cap_frames = 0;
}

template <typename T>
static cudaError_t slot_alloc(Slot* s, T** p, size_t count) {
  void* raw = nullptr;
  cudaError_t e = cudaMalloc(&raw, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
  if (e == cudaSuccess) {
    s->owned.push_back(raw);
    *p = static_cast<T*>(raw);
  }
  return e;
}
static cudaError_t slot_alloc_bytes(Slot* s, void** p, size_t bytes) {
  uint8_t* q = nullptr;
  cudaError_t e = slot_alloc(s, &q, bytes);
  *p = q;
  return e;
}

static int ensure_workspace(ctag_detector* d, Slot* s, int n, int w, int h) {
  if (s->cap_frames >= n && s->w == w && s->h == h) return CTAG_OK;
  const int cap = n;
  free_workspace(s);
  s->geo = make_geom(w, h);
  s->w = w;
  s->h = h;
  const FrameGeom& g = s->geo;
  s->gray_fstride = (size_t)g.gpitch * g.h;
  s->bin_fstride = (size_t)g.bpitch * g.hh;
  const size_t nb = (size_t)g.nblocks * cap;
  s->legal_cap = g.hw * g.hh / kAreaMin + 1;  // every legal component has >= 30 pixels
  if (s->legal_cap > 65536) s->legal_cap = 65536;
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_gray, s->gray_fstride * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_bin, s->bin_fstride * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_labels, nb));
  // CCL writes labels only inside tiles that hold foreground; the quad stage loads the label next to every pixel word it
  // reads and uses it only where the pixel is set.  Background once, so that those loads never see unwritten memory.
  CTAG_CUDA_CHECK(cudaMemset(s->d_labels, 0xFF, nb * sizeof(int)));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_st_area, nb));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_st_x0, nb));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_st_y0, nb));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_st_x1, nb));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_st_y1, nb));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_roots_tmp, ccl_roots_ints(g) * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_tile_list, ccl_tile_list_ints(g, cap)));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_seg_flags, ccl_seg_flag_bytes(g) * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_tile_any, ccl_tile_hint_bytes(g) * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_legal, (size_t)s->legal_cap * 6 * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_counters, (size_t)4 * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_prefix, (size_t)cap + 1));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_qctl, 16));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_quad_status, (size_t)s->legal_cap * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_quad_corners, (size_t)s->legal_cap * 8 * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_quads, (size_t)kQuadCap * 8 * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_quad_comp, (size_t)kQuadCap * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_n_quads, (size_t)cap));
  CTAG_CUDA_CHECK(slot_alloc_bytes(s, &s->d_geom, sizeof_quad_geom() * kQuadCap * cap));
  CTAG_CUDA_CHECK(slot_alloc_bytes(s, &s->d_feats, sizeof_feature_rec() * kFeatCap * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_fstate, (size_t)4 * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_packed_count, 4));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_summary, (size_t)12 * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_markers, (size_t)kMarkerCap * cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_packed, (size_t)kMarkerCap * cap));
  CTAG_CUDA_CHECK(cudaMallocHost(&s->h_summary, sizeof(int) * 12 * cap));
  CTAG_CUDA_CHECK(cudaMallocHost(&s->h_packed, sizeof(ctag_marker) * kMarkerCap * cap));
  s->edge_warps = quad_edge_warps(d->sms);
  s->exact_ctas = quad_exact_ctas(d->sms);
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_quad_scratch, quad_scratch_bytes_per_warp(g) * s->edge_warps));
  CTAG_CUDA_CHECK(slot_alloc_bytes(s, &s->d_traj, quad_traj_bytes_per_cta() * s->exact_ctas));
  // components that reach four edges / their cluster points: the reference caps a frame at 1000 quads
  // (isVisited[1000]), so 1024 four-edge components per frame is already past its envelope; overflow drops the
  // component and flags the frame instead of writing out of bounds
  s->fit_per_frame = s->legal_cap < 1024 ? s->legal_cap : 1024;
  s->fit_cap = cap * s->fit_per_frame;
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_frame_fit, (size_t)2 * cap));
  CTAG_CUDA_CHECK(slot_alloc_bytes(s, &s->d_fit_results, quad_fitresult_bytes() * 80 * (size_t)s->fit_cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_exact_list, (size_t)4 * s->fit_cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_fit_order, (size_t)4 * s->fit_cap));
  // boundary points are distinct foreground pixels: one int per half-res pixel cannot overflow
  const long long per_frame_pts = (long long)g.hw * g.hh;
  s->pool_cap = (int)(per_frame_pts * cap < 0x7fffffff ? per_frame_pts * cap : 0x7fffffff);
  CTAG_CUDA_CHECK(slot_alloc_bytes(s, &s->d_fits, quad_fitrec_bytes() * s->fit_cap));
  s->tail_cap = quad_tail_cap(d->sms);
  CTAG_CUDA_CHECK(slot_alloc_bytes(s, &s->d_tails, quad_tailrec_bytes() * s->tail_cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_lines, (size_t)16 * s->fit_cap));
  CTAG_CUDA_CHECK(slot_alloc(s, &s->d_pool, (size_t)s->pool_cap));
  s->cap_frames = cap;
  return CTAG_OK;
}

static int ensure_stage(Slot* s, size_t bytes) {
  if (s->stage_bytes >= bytes) return CTAG_OK;
  cudaFree(s->d_stage);
  s->d_stage = nullptr;
  s->stage_bytes = 0;
  CTAG_CUDA_CHECK(cudaMalloc(&s->d_stage, bytes));
  s->stage_bytes = bytes;
  return CTAG_OK;
}

// Enqueues the whole detect path for one batch on slot `s` (frames already on the device).
static int enqueue_on_slot(ctag_detector* d, Slot* s, const void* frames_dev, int n, int w, int h, size_t pitch,
                           size_t frame_stride, int channels, int window, int corner_subpix, int subpix_dist) {
  int rc = ensure_workspace(d, s, n, w, h);
  if (rc != CTAG_OK) return rc;
  if (window != kWin) {
    if (!s->d_half) CTAG_CUDA_CHECK(slot_alloc(s, &s->d_half, s->bin_fstride * s->cap_frames));
    const size_t need = 2 * (size_t)s->cap_frames * ((s->geo.hw + window - 1) / window) * ((s->geo.hh + window - 1) / window);
    if (s->tiles_bytes < need) {
      CTAG_CUDA_CHECK(slot_alloc(s, &s->d_tiles, need));  // the previous (smaller) buffer stays owned until the workspace is freed
      s->tiles_bytes = need;
    }
  }
  s->n = n;
  s->channels = channels;
  s->subpix = corner_subpix;
  s->launches = 0;
  if (channels == 1) {
    s->gray = static_cast<const uint8_t*>(frames_dev);
    s->gray_pitch = pitch;
    s->gray_fs = frame_stride;
  } else {
    s->gray = s->d_gray;
    s->gray_pitch = s->geo.gpitch;
    s->gray_fs = s->gray_fstride;
  }
  cudaStream_t st = s->stream;
  const uint8_t* tile_any = nullptr;
  if (window == kWin) {
    CTAG_CUDA_CHECK(cudaMemsetAsync(s->d_tile_any, 0, ccl_tile_hint_bytes(s->geo) * n, st));
    tile_any = s->d_tile_any;
    rc = launch_front(frames_dev, n, s->geo, channels, pitch, frame_stride, s->d_gray, s->gray_fstride, s->d_bin,
                      s->bin_fstride, ccl_tile_hint(s->geo, s->d_tile_any), st, s->ev[0]);
    s->launches += 1;
  } else {
    CTAG_CUDA_CHECK(cudaEventRecord(s->ev[0], st));
    rc = launch_front_generic(frames_dev, n, s->geo, channels, pitch, frame_stride, window, s->d_gray, s->gray_fstride,
                              s->d_half, s->d_tiles, s->d_bin, s->bin_fstride, st, &s->launches);
  }
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaEventRecord(s->ev[1], st));
  rc = launch_ccl(s->d_bin, s->bin_fstride, n, s->geo, s->d_labels, s->d_st_area, s->d_st_x0, s->d_st_y0, s->d_st_x1,
                  s->d_st_y1, s->d_roots_tmp, s->d_tile_list, s->d_seg_flags, tile_any, s->d_legal, s->legal_cap, s->d_counters, st, &s->launches);
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaEventRecord(s->ev[2], st));
  rc = launch_quad(n, s->geo, s->d_bin, s->bin_fstride, s->d_labels, s->d_legal, s->legal_cap, s->d_counters, s->d_prefix,
                   s->d_qctl, s->d_quad_scratch, s->edge_warps, s->d_fits, s->fit_cap, s->fit_per_frame, s->d_frame_fit, s->d_pool, s->pool_cap,
                   d->d_pick_table, kPickTableMax, s->d_fit_results, s->d_exact_list, s->d_fit_order, s->d_tails, s->tail_cap, s->d_traj, s->exact_ctas, d->sms,
                   s->d_lines, s->d_quad_status, s->d_quad_corners, kQuadCap, s->d_quads, s->d_quad_comp, s->d_n_quads, st,
                   &s->launches);
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaEventRecord(s->ev[3], st));
  rc = launch_features(n, s->geo, s->d_quads, s->d_n_quads, kQuadCap, s->d_geom, s->d_feats, kFeatCap, d->feature_size,
                       s->d_fstate, s->gray, s->gray_pitch, s->gray_fs, corner_subpix, subpix_dist, st, &s->launches);
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaEventRecord(s->ev[4], st));
  rc = launch_decode(n, s->d_feats, kFeatCap, s->d_fstate, d->d_state, d->rows, d->cols, d->feature_size, s->d_markers,
                     kMarkerCap, s->d_counters, s->d_n_quads, kQuadCap, s->d_qctl + 4 /* QC_OVERFLOW */, s->d_frame_fit + n, s->d_packed,
                     s->d_packed_count, s->d_summary, st, &s->launches);
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaEventRecord(s->ev[5], st));
  CTAG_CUDA_CHECK(cudaMemcpyAsync(s->h_summary, s->d_summary, sizeof(int) * 12 * n, cudaMemcpyDeviceToHost, st));
  s->busy = true;
  return CTAG_OK;
}

static int collect_slot(ctag_detector* d, Slot* s, ctag_marker* out, int cap_per_frame, int* n_out, ctag_frame_info* info) {
  s->busy = false;
  CTAG_CUDA_CHECK(cudaStreamSynchronize(s->stream));
  for (int i = 0; i < CTAG_STAGE_COUNT; ++i) cudaEventElapsedTime(&d->stage_ms[i], s->ev[i], s->ev[i + 1]);
  for (int i = 0; i <= CTAG_STAGE_COUNT; ++i) cudaEventElapsedTime(&d->stage_stamp_ms[i], d->epoch, s->ev[i]);
  d->last_launches = s->launches;
  int total = 0;
  for (int f = 0; f < s->n; ++f) total += s->h_summary[12 * f + 10];
  if (total > 0) {
    CTAG_CUDA_CHECK(cudaMemcpyAsync(s->h_packed, s->d_packed, sizeof(ctag_marker) * total, cudaMemcpyDeviceToHost, s->stream));
    CTAG_CUDA_CHECK(cudaStreamSynchronize(s->stream));
  }
  for (int f = 0; f < s->n; ++f) {
    const int* sm = s->h_summary + 12 * f;
    if (n_out) n_out[f] = sm[6];
    if (info) {
      memset(&info[f], 0, sizeof(ctag_frame_info));
      info[f].status = sm[0];
      info[f].n_labels = sm[1];
      info[f].n_legal = sm[2];
      info[f].n_quads = sm[3];
      info[f].n_features = sm[4];
      info[f].n_groups = sm[5];
      info[f].n_markers = sm[6];
      info[f].flagged = sm[7];
      info[f].stale_ids = sm[8];
    }
    if (out && cap_per_frame > 0) {
      int ncopy = sm[10] < cap_per_frame ? sm[10] : cap_per_frame;
      if (ncopy > 0) memcpy(out + (size_t)f * cap_per_frame, s->h_packed + sm[9], sizeof(ctag_marker) * ncopy);
    }
  }
  return CTAG_OK;
}

static int check_args(ctag_detector* d, const void* frames, int n, int w, int h, int channels, int adaptive_thresh,
                      int corner_subpix, int subpix_dist) {
  if (!d || !frames || n <= 0 || w <= 0 || h <= 0) return CTAG_ERR_ARG;
  if ((w & 1) || (h & 1)) return CTAG_ERR_ARG;  // exact 2x decimation needs even sizes
  if (w / 2 > 4095 || h / 2 > 4095) return CTAG_ERR_UNSUPPORTED;  // packed 12-bit coordinates in the trace stack
  if (channels != 1 && channels != 3) return CTAG_ERR_ARG;
  if (adaptive_thresh < 1) return CTAG_ERR_ARG;
  if (corner_subpix && (subpix_dist < 0 || subpix_dist > 64)) return CTAG_ERR_ARG;
  if (decode_smem_bytes(d->rows, d->cols) > 200 * 1024) return CTAG_ERR_UNSUPPORTED;
  return CTAG_OK;
}

// Host frames: cut the batch into chunks and rotate the slots, so that the H2D copy of chunk c+1 (its own stream)
// overlaps the kernels of chunk c.  The 2-D copy normalises any host pitch to a 16-byte multiple (TMA).
static int detect_batch_host(ctag_detector* d, const void* frames, int n, int w, int h, size_t pitch, size_t frame_stride,
                             int channels, int adaptive_thresh, int corner_subpix, int subpix_dist, ctag_marker* out,
                             int cap_per_frame, int* n_out, ctag_frame_info* info) {
  int rc = CTAG_OK;
  const size_t dpitch = (size_t)round_up(w * channels, 16);
  const size_t dfs = dpitch * h;
  // Small batches stay whole (the debug getters then see all frames).  Larger ones go in chunks of about 192 MiB (8 4K
  // BGR frames), at most a quarter of the batch: what cannot be hidden behind the copies is the processing of the LAST
  // chunk, and the sparse stages' latency barely shrinks with the chunk, so small chunks end sooner (64 4K BGR frames:
  // 30.7 ms with 16-frame chunks, 30.0 ms with 8; the bare copy takes 28.6 ms).  ctag_set_option("chunk_frames") overrides.
  int chunk = n;
  if (n >= 8) {
    const size_t target = (size_t)192 << 20;
    chunk = (int)((target + dfs - 1) / dfs);
    if (chunk > (n + 3) / 4) chunk = (n + 3) / 4;
    if (chunk < 1) chunk = 1;
  }
  if (d->chunk_frames > 0) chunk = d->chunk_frames < n ? d->chunk_frames : n;
  int done = 0, queued = 0;
  int q_first[kMaxSlots], q_count[kMaxSlots];
  // at most three chunks in flight: copies queued on more streams share the copy engine and every chunk arrives later
  // (64 4K BGR frames: 29.8 ms with 3 or 4, 31.2 ms with 6).  The synchronous calls therefore rotate over the first three
  // workspaces only (nothing is pending at entry), which also keeps the staging memory to three buffers.
  const int ring = kSlots < 3 ? kSlots : 3;
  d->next_enqueue = d->next_collect = 0;
  while (done < n) {
    while (queued < n && d->in_flight < ring) {
      const int c = n - queued < chunk ? n - queued : chunk;
      Slot* s = &d->slot[d->next_enqueue];
      if (d->debug_fail_chunk >= 0 && queued / chunk == d->debug_fail_chunk) {
        d->debug_fail_chunk = -1;  // one shot
        set_last_error_text("injected failure (debug_fail_chunk)");
        return CTAG_ERR_CUDA;
      }
      rc = ensure_stage(s, dfs * c);
      if (rc != CTAG_OK) return rc;
      if (pitch == dpitch && frame_stride == dfs) {
        // densely packed frames with a TMA-compatible pitch: one linear copy for the whole chunk
        CTAG_CUDA_CHECK(cudaMemcpyAsync(s->d_stage, static_cast<const uint8_t*>(frames) + frame_stride * queued, dfs * c,
                                        cudaMemcpyHostToDevice, s->stream));
      } else {
        // one 2-D copy per frame normalises the pitch and keeps frame_stride != pitch*h inputs correct
        for (int f = 0; f < c; ++f)
          CTAG_CUDA_CHECK(cudaMemcpy2DAsync(s->d_stage + dfs * f, dpitch,
                                            static_cast<const uint8_t*>(frames) + frame_stride * (queued + f), pitch,
                                            (size_t)w * channels, h, cudaMemcpyHostToDevice, s->stream));
      }
      rc = enqueue_on_slot(d, s, s->d_stage, c, w, h, dpitch, dfs, channels, adaptive_thresh, corner_subpix, subpix_dist);
      if (rc != CTAG_OK) return rc;
      q_first[d->next_enqueue] = queued;
      q_count[d->next_enqueue] = c;
      d->next_enqueue = (d->next_enqueue + 1) % ring;
      d->in_flight += 1;
      queued += c;
    }
    const int si = d->next_collect;
    Slot* s = &d->slot[si];
    const int first = q_first[si], c = q_count[si];
    d->last = si;
    d->next_collect = (d->next_collect + 1) % ring;
    d->in_flight -= 1;
    rc = collect_slot(d, s, out ? out + (size_t)first * cap_per_frame : nullptr, cap_per_frame, n_out ? n_out + first : nullptr,
                      info ? info + first : nullptr);
    if (rc != CTAG_OK) return rc;
    if (out)
      for (int f = 0; f < c; ++f) {
        const int nm = s->h_summary[12 * f + 10];  // markers stored for the frame (n_out may be NULL)
        for (int k = 0; k < nm && k < cap_per_frame; ++k) out[(size_t)(first + f) * cap_per_frame + k].frame = first + f;
      }
    done += c;
  }
  d->next_enqueue = d->next_collect = 0;
  return CTAG_OK;
}

// ---- compressed ingest (the reference reads its frames through cv::VideoCapture / imread, main.cpp:29,45-52)
// nvJPEG decodes straight into the staging buffer in the interleaved BGR layout (16-byte pitch) the front kernel's tensor
// map expects, on the slot's stream, so the decode of chunk c+1 overlaps the detection of chunk c and only the compressed
// bytes cross PCIe.  Backend 0 = the NVJPG hardware engine, 1 = nvJPEG's CUDA decoder (used when the engine or the
// bitstream does not qualify).
static nvjpegHandle_t jpeg_handle(ctag_detector* d, int backend) {
  if (!d->jpeg_tried[backend]) {
    d->jpeg_tried[backend] = 1;
    nvjpegHandle_t h = nullptr;
    if (nvjpeg_api().CreateEx(backend == 0 ? NVJPEG_BACKEND_HARDWARE : NVJPEG_BACKEND_DEFAULT, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT,
                              &h) == NVJPEG_STATUS_SUCCESS)
      d->jpeg_handle[backend] = h;
  }
  return d->jpeg_handle[backend];
}

static int jpeg_decode_chunk(ctag_detector* d, Slot* s, const uint8_t* const* jpeg, const size_t* bytes, int c, size_t dpitch, size_t dfs,
                             int first_backend) {
  {  // one batch = one frame size (the tensor map is per batch)
    nvjpegHandle_t h = jpeg_handle(d, 1) ? jpeg_handle(d, 1) : jpeg_handle(d, 0);
    if (!h) {
      set_last_error_text("nvjpegCreateEx failed");
      return CTAG_ERR_CUDA;
    }
    for (int f = 0; f < c; ++f) {
      int comps = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
      nvjpegChromaSubsampling_t ss;
      if (nvjpeg_api().GetImageInfo(h, jpeg[f], bytes[f], &comps, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS || ws[0] != s->w_expect ||
          hs[0] != s->h_expect) {
        set_last_error_text("ctag_detect_batch_jpeg: a frame is not a readable JPEG of the batch's size");
        return CTAG_ERR_ARG;
      }
    }
  }
  std::vector<nvjpegImage_t> dst(c);
  for (int f = 0; f < c; ++f) {
    memset(&dst[f], 0, sizeof(nvjpegImage_t));
    dst[f].channel[0] = s->d_stage + dfs * f;
    dst[f].pitch[0] = dpitch;
  }
  for (int b = first_backend; b < 2; ++b) {
    nvjpegHandle_t h = jpeg_handle(d, b);
    if (!h) continue;
    if (!s->jpeg_state[b] && nvjpeg_api().JpegStateCreate(h, &s->jpeg_state[b]) != NVJPEG_STATUS_SUCCESS) continue;
    if (s->jpeg_batch[b] != c) {
      if (nvjpeg_api().DecodeBatchedInitialize(h, s->jpeg_state[b], c, 1, NVJPEG_OUTPUT_BGRI) != NVJPEG_STATUS_SUCCESS) continue;
      s->jpeg_batch[b] = c;
    }
    if (nvjpeg_api().DecodeBatched(h, s->jpeg_state[b], jpeg, bytes, dst.data(), s->stream) == NVJPEG_STATUS_SUCCESS) {
      d->jpeg_backend_used = b;
      return CTAG_OK;
    }
    s->jpeg_batch[b] = 0;  // the state may be half way through a batch: initialise it again next time
  }
  set_last_error_text("nvJPEG could not decode the batch (baseline JPEG, 3 components, equal sizes expected)");
  return CTAG_ERR_UNSUPPORTED;
}

// ---- the CUDA decoder (jpeg.cu): baseline JPEG with restart markers, one GPU thread per restart interval ------------
template <typename T>
static int grow(T** p, size_t* cap, size_t need) {
  if (*cap >= need) return CTAG_OK;
  cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  const size_t want = need + need / 4;
  CTAG_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(p), want * sizeof(T)));
  // zero once: the JPEG byte buffer is read in 16-byte pieces that reach into the alignment gaps between frames and a few
  // bytes past the last one (never consumed), and those bytes should not be unwritten memory
  CTAG_CUDA_CHECK(cudaMemset(*p, 0, want * sizeof(T)));
  *cap = want;
  return CTAG_OK;
}

// Parses, uploads and decodes frames [0, c) of a chunk into s->d_stage.  Returns CTAG_OK, CTAG_ERR_UNSUPPORTED when a
// frame is outside the decoder's envelope (*unsupported_frame says which), or an error.
static int jpeg_cuda_chunk(ctag_detector* d, Slot* s, const uint8_t* const* jpeg, const size_t* bytes, int c, int w, int h, size_t dpitch,
                           size_t dfs) {
  const size_t hb = jpeg_header_bytes();
  if (s->jhdr_cap < c) {
    cudaFree(s->d_jhdr);
    cudaFree(s->d_jstatus);
    cudaFreeHost(s->h_jhdr);
    cudaFreeHost(s->h_jstatus);
    s->d_jhdr = s->h_jhdr = nullptr;
    s->d_jstatus = s->h_jstatus = nullptr;
    s->jhdr_cap = 0;
    CTAG_CUDA_CHECK(cudaMalloc(&s->d_jhdr, hb * c));
    CTAG_CUDA_CHECK(cudaMalloc(&s->d_jstatus, sizeof(int) * c));
    CTAG_CUDA_CHECK(cudaMallocHost(&s->h_jhdr, hb * c));
    CTAG_CUDA_CHECK(cudaMallocHost(&s->h_jstatus, sizeof(int) * c));
    s->jhdr_cap = c;
  }
  size_t total_bytes = 0, plane_stride = 0;
  int total_ivl = 0, max_ivl = 0, max_blocks = 0, all_420 = 1;
  std::vector<size_t> scan_off(c), scan_len(c), place(c);
  for (int f = 0; f < c; ++f) {
    int fw = 0, fh = 0, ni = 0;
    size_t pb = 0;
    const int rc = jpeg_parse_frame(jpeg[f], bytes[f], s->h_jhdr + hb * f, &scan_off[f], &scan_len[f], &fw, &fh, &ni, &pb);
    if (rc == 2) return CTAG_ERR_UNSUPPORTED;
    if (rc != 0 || fw != w || fh != h) {
      set_last_error_text("ctag_detect_batch_jpeg: a frame is not a readable JPEG of the batch's size");
      return CTAG_ERR_ARG;
    }
    place[f] = total_bytes;  // frames start 16-byte aligned in the device buffer
    jpeg_place_frame(s->h_jhdr + hb * f, (uint32_t)(total_bytes + scan_off[f]), (uint32_t)scan_len[f], total_ivl);
    total_bytes += (bytes[f] + 15) & ~(size_t)15;
    total_ivl += ni + 1;
    if (ni > max_ivl) max_ivl = ni;
    if (pb > plane_stride) plane_stride = pb;
    const int nb = jpeg_frame_blocks(s->h_jhdr + hb * f);
    if (nb > max_blocks) max_blocks = nb;
    all_420 &= jpeg_frame_is_420(s->h_jhdr + hb * f);
  }
  const size_t coef_stride = (size_t)max_blocks * 64, last_stride = ((size_t)max_blocks + 15) & ~(size_t)15;
  if (total_bytes + 32 >= 0xFFFFFFFFull) return CTAG_ERR_UNSUPPORTED;  // 32-bit offsets inside a chunk
  int rc = grow(&s->d_jbytes, &s->jbytes_cap, total_bytes + 32);
  if (rc == CTAG_OK) rc = grow(&s->d_jivl, &s->jivl_cap, (size_t)total_ivl);
  if (rc == CTAG_OK) rc = grow(&s->d_jplanes, &s->jplanes_cap, plane_stride * c);
  if (rc == CTAG_OK) rc = grow(&s->d_jcoefs, &s->jcoefs_cap, coef_stride * c);
  if (rc == CTAG_OK) rc = grow(&s->d_jlast, &s->jlast_cap, last_stride * c);
  if (rc == CTAG_OK) rc = ensure_stage(s, dfs * c);
  if (rc != CTAG_OK) return rc;
  for (int f = 0; f < c; ++f)
    CTAG_CUDA_CHECK(cudaMemcpyAsync(s->d_jbytes + place[f], jpeg[f], bytes[f], cudaMemcpyHostToDevice, s->stream));
  CTAG_CUDA_CHECK(cudaMemcpyAsync(s->d_jhdr, s->h_jhdr, hb * c, cudaMemcpyHostToDevice, s->stream));
  int launches = 0;
  rc = launch_jpeg_decode(s->d_jhdr, c, s->d_jbytes, s->d_jivl, max_ivl, s->d_jcoefs, coef_stride, s->d_jlast, last_stride, max_blocks,
                          s->d_jplanes, plane_stride, s->d_stage, dpitch, dfs, w, h, all_420, s->d_jstatus, s->stream, &launches);
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaMemcpyAsync(s->h_jstatus, s->d_jstatus, sizeof(int) * c, cudaMemcpyDeviceToHost, s->stream));
  s->jpeg_frames = c;
  d->jpeg_backend_used = 2;
  return CTAG_OK;
}

static int detect_batch_jpeg(ctag_detector* d, const uint8_t* const* jpeg, const size_t* bytes, int n, int w, int h, int adaptive_thresh,
                             int corner_subpix, int subpix_dist, ctag_marker* out, int cap_per_frame, int* n_out, ctag_frame_info* info) {
  const size_t dpitch = (size_t)round_up(w * 3, 16), dfs = dpitch * h;
  bool use_cuda = d->jpeg_decoder != 1;
  // Chunks: the Huffman kernel runs one thread per restart interval, a long sequential chain each, so it wants MANY frames
  // per launch to fill the machine (a 4K frame with 16-MCU intervals is 2,025 threads): half the batch per chunk, up to
  // about 1 GiB of decoded frames (the second chunk's decode overlaps the first chunk's detection).  nvJPEG decodes on the
  // host side of the stream: small chunks as for raw host frames.
  int chunk = n;
  if (n >= 8) {
    const size_t target = use_cuda ? (size_t)1 << 30 : (size_t)192 << 20;
    chunk = (int)((target + dfs - 1) / dfs);
    const int share = use_cuda ? (n + 1) / 2 : (n + 3) / 4;
    if (chunk > share) chunk = share;
    if (chunk < 1) chunk = 1;
  }
  if (d->chunk_frames > 0) chunk = d->chunk_frames < n ? d->chunk_frames : n;
  int done = 0, queued = 0, backend = 0;
  int q_first[kMaxSlots], q_count[kMaxSlots];
  const int ring = kSlots < 3 ? kSlots : 3;  // as in detect_batch_host: three workspaces, nothing pending at entry
  d->next_enqueue = d->next_collect = 0;
  while (done < n) {
    while (queued < n && d->in_flight < ring) {
      const int c = n - queued < chunk ? n - queued : chunk;
      Slot* s = &d->slot[d->next_enqueue];
      int rc = CTAG_ERR_UNSUPPORTED;
      s->jpeg_frames = 0;
      s->w_expect = w;
      s->h_expect = h;
      if (use_cuda) {
        rc = jpeg_cuda_chunk(d, s, jpeg + queued, bytes + queued, c, w, h, dpitch, dfs);
        if (rc == CTAG_ERR_UNSUPPORTED && queued == 0) use_cuda = false;  // e.g. no restart markers: nvJPEG for this batch
        else if (rc != CTAG_OK) return rc;
      }
      if (!use_cuda) {
        if (!nvjpeg_api().ok) {
          set_last_error_text("the frames are outside the CUDA decoder's envelope (baseline JPEG with restart markers) and libnvjpeg.so.12 is not available");
          return CTAG_ERR_UNSUPPORTED;
        }
        rc = ensure_stage(s, dfs * c);
        if (rc != CTAG_OK) return rc;
        rc = jpeg_decode_chunk(d, s, jpeg + queued, bytes + queued, c, dpitch, dfs, backend);
        if (rc != CTAG_OK) return rc;
        backend = d->jpeg_backend_used;  // once the engine has refused a chunk, do not ask it again for this batch
      }
      rc = enqueue_on_slot(d, s, s->d_stage, c, w, h, dpitch, dfs, 3, adaptive_thresh, corner_subpix, subpix_dist);
      if (rc != CTAG_OK) return rc;
      q_first[d->next_enqueue] = queued;
      q_count[d->next_enqueue] = c;
      d->next_enqueue = (d->next_enqueue + 1) % ring;
      d->in_flight += 1;
      queued += c;
    }
    const int si = d->next_collect;
    Slot* s = &d->slot[si];
    const int first = q_first[si], c = q_count[si];
    d->last = si;
    d->next_collect = (d->next_collect + 1) % ring;
    d->in_flight -= 1;
    int rc = collect_slot(d, s, out ? out + (size_t)first * cap_per_frame : nullptr, cap_per_frame, n_out ? n_out + first : nullptr,
                          info ? info + first : nullptr);
    if (rc != CTAG_OK) return rc;
    for (int f = 0; f < s->jpeg_frames; ++f)
      if (s->h_jstatus[f] != 0) {
        set_last_error_text("ctag_detect_batch_jpeg: restart markers of a frame do not match its header (corrupt stream)");
        return CTAG_ERR_ARG;
      }
    if (out)
      for (int f = 0; f < c; ++f) {
        const int nm = s->h_summary[12 * f + 10];
        for (int k = 0; k < nm && k < cap_per_frame; ++k) out[(size_t)(first + f) * cap_per_frame + k].frame = first + f;
      }
    done += c;
  }
  return CTAG_OK;
}

extern "C" {

int ctag_create(ctag_detector** out, const int32_t* state, int rows, int cols, int feature_size, int cuda_device) {
  if (!out || !state || rows <= 0 || cols <= 0 || feature_size <= 0) return CTAG_ERR_ARG;
  *out = nullptr;
  for (int i = 0; i < rows * cols; ++i)
    if (!(state[i] >= 0 && state[i] <= 63)) return CTAG_ERR_DICTIONARY;  // check_dictionary, CylinderTag.cpp:56-65
  int dev = 0, sms = 0;
  int rc = select_device(cuda_device, &dev, &sms);
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaSetDevice(dev));
  ctag_detector* d = new ctag_detector();
  d->device = dev;
  d->sms = sms;
  d->state.assign(state, state + rows * cols);
  d->rows = rows;
  d->cols = cols;
  d->feature_size = feature_size;
  if (const char* e = getenv("CTAG_CHUNK")) d->chunk_frames = atoi(e) > 0 ? atoi(e) : 0;  // read once, at creation
  // initial subsets of cv::fitLine's 20 restarts for every point count up to kPickTableMax (fit_core.cuh)
  std::vector<uint16_t> table((size_t)kPickTableMax * 200);
  quad_build_pick_table(table.data(), kPickTableMax);
  bool ok = cudaMalloc(&d->d_state, sizeof(int32_t) * rows * cols) == cudaSuccess &&
            cudaMemcpy(d->d_state, state, sizeof(int32_t) * rows * cols, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMalloc(&d->d_pick_table, table.size() * sizeof(uint16_t)) == cudaSuccess &&
            cudaMemcpy(d->d_pick_table, table.data(), table.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && cudaEventCreate(&d->epoch) == cudaSuccess && cudaEventRecord(d->epoch, 0) == cudaSuccess &&
       cudaEventSynchronize(d->epoch) == cudaSuccess;
  for (Slot& s : d->slot) {
    ok = ok && cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& e : s.ev) ok = ok && cudaEventCreate(&e) == cudaSuccess;
  }
  if (!ok) {
    set_last_error("detector setup", cudaGetLastError(), __FILE__, __LINE__);
    ctag_destroy(d);
    return CTAG_ERR_CUDA;
  }
  *out = d;
  return CTAG_OK;
}

int ctag_create_from_file(ctag_detector** out, const char* marker_path, int cuda_device) {
  if (!out || !marker_path) return CTAG_ERR_ARG;
  std::ifstream in(marker_path);
  if (!in.is_open()) return CTAG_ERR_FILE;  // CylinderTag.cpp:19-22
  int n = 0, cols = 0, fsz = 0;
  in >> n >> cols >> fsz;  // CylinderTag.cpp:24-25
  if (!in || n <= 0 || cols <= 0) return CTAG_ERR_FILE;
  std::vector<int32_t> st((size_t)n * cols, 0);
  for (auto& v : st) in >> v;  // missing values stay 0 like a failed operator>> on a zero-initialised Mat1i
  return ctag_create(out, st.data(), n, cols, fsz, cuda_device);
}

void ctag_destroy(ctag_detector* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  for (Slot& s : d->slot) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    free_workspace(&s);
    cudaFree(s.d_stage);
    cudaFree(s.d_jbytes);
    cudaFree(s.d_jplanes);
    cudaFree(s.d_jcoefs);
    cudaFree(s.d_jlast);
    cudaFree(s.d_jhdr);
    cudaFree(s.d_jivl);
    cudaFree(s.d_jstatus);
    cudaFreeHost(s.h_jhdr);
    cudaFreeHost(s.h_jstatus);
    for (auto& e : s.ev)
      if (e) cudaEventDestroy(e);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  if (nvjpeg_api().ok) {
    for (Slot& s : d->slot)
      for (auto& st : s.jpeg_state)
        if (st) nvjpeg_api().JpegStateDestroy(st);
    for (auto& h : d->jpeg_handle)
      if (h) nvjpeg_api().Destroy(h);
  }
  if (d->epoch) cudaEventDestroy(d->epoch);
  cudaFree(d->d_state);
  cudaFree(d->d_pick_table);
  delete d;
}

int ctag_set_option(ctag_detector* d, const char* key, int value) {
  if (!d || !key) return CTAG_ERR_ARG;
  if (d->in_flight != 0) return CTAG_ERR_ARG;
  if (!strcmp(key, "chunk_frames")) {
    if (value < 0) return CTAG_ERR_ARG;
    d->chunk_frames = value;
    return CTAG_OK;
  }
  if (!strcmp(key, "jpeg_decoder")) {
    if (value < 0 || value > 1) return CTAG_ERR_ARG;
    d->jpeg_decoder = value;
    return CTAG_OK;
  }
  if (!strcmp(key, "debug_fail_chunk")) {
    d->debug_fail_chunk = value;
    return CTAG_OK;
  }
  return CTAG_ERR_ARG;
}

int ctag_get_dictionary(const ctag_detector* d, int* rows, int* cols, int* feature_size, int32_t* state_out, int cap) {
  if (!d) return CTAG_ERR_ARG;
  if (rows) *rows = d->rows;
  if (cols) *cols = d->cols;
  if (feature_size) *feature_size = d->feature_size;
  if (state_out) {
    if (cap < d->rows * d->cols) return CTAG_ERR_ARG;
    memcpy(state_out, d->state.data(), sizeof(int32_t) * d->state.size());
  }
  return CTAG_OK;
}

int ctag_detect_batch_enqueue(ctag_detector* d, const void* frames_dev, int n, int w, int h, size_t pitch,
                              size_t frame_stride, int channels, int adaptive_thresh, int corner_subpix,
                              int subpix_dist) {
  int rc = check_args(d, frames_dev, n, w, h, channels, adaptive_thresh, corner_subpix, subpix_dist);
  if (rc != CTAG_OK) return rc;
  if (d->in_flight >= kSlots) return CTAG_ERR_ARG;  // collect first
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  if (frame_stride == 0) frame_stride = pitch * (size_t)h;
  Slot* s = &d->slot[d->next_enqueue];
  rc = enqueue_on_slot(d, s, frames_dev, n, w, h, pitch, frame_stride, channels, adaptive_thresh, corner_subpix, subpix_dist);
  if (rc != CTAG_OK) return rc;
  d->next_enqueue = (d->next_enqueue + 1) % kSlots;
  d->in_flight += 1;
  return CTAG_OK;
}

int ctag_detect_batch_collect(ctag_detector* d, ctag_marker* out, int cap_per_frame, int* n_out, ctag_frame_info* info) {
  if (!d || d->in_flight <= 0) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  Slot* s = &d->slot[d->next_collect];
  d->last = d->next_collect;
  d->next_collect = (d->next_collect + 1) % kSlots;
  d->in_flight -= 1;
  return collect_slot(d, s, out, cap_per_frame, n_out, info);
}

int ctag_detect_batch(ctag_detector* d, const void* frames, int n, int w, int h, size_t pitch, size_t frame_stride,
                      int channels, int is_device, int adaptive_thresh, int corner_subpix, int subpix_dist,
                      ctag_marker* out, int cap_per_frame, int* n_out, ctag_frame_info* info) {
  int rc = check_args(d, frames, n, w, h, channels, adaptive_thresh, corner_subpix, subpix_dist);
  if (rc != CTAG_OK) return rc;
  if (d->in_flight != 0) return CTAG_ERR_ARG;  // the synchronous call does not mix with pending asynchronous batches
  if (frame_stride == 0) frame_stride = pitch * (size_t)h;
  if (is_device) {
    rc = ctag_detect_batch_enqueue(d, frames, n, w, h, pitch, frame_stride, channels, adaptive_thresh, corner_subpix,
                                   subpix_dist);
    if (rc != CTAG_OK) return rc;
    return ctag_detect_batch_collect(d, out, cap_per_frame, n_out, info);
  }
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  if (pitch < (size_t)w * channels) return CTAG_ERR_ARG;
  rc = detect_batch_host(d, frames, n, w, h, pitch, frame_stride, channels, adaptive_thresh, corner_subpix, subpix_dist, out,
                         cap_per_frame, n_out, info);
  if (rc != CTAG_OK) drain_slots(d);  // the handle stays usable after a failed call
  return rc;
}

int ctag_detect_batch_multi(ctag_detector* const* dets, int n_det, const void* frames, int n, int w, int h, size_t pitch,
                            size_t frame_stride, int channels, int adaptive_thresh, int corner_subpix, int subpix_dist,
                            ctag_marker* out, int cap_per_frame, int* n_out, ctag_frame_info* info) {
  if (!dets || n_det <= 0 || !frames || n <= 0) return CTAG_ERR_ARG;
  for (int g = 0; g < n_det; ++g) {
    if (!dets[g]) return CTAG_ERR_ARG;
    for (int k = 0; k < g; ++k)
      if (dets[k] == dets[g]) return CTAG_ERR_ARG;  // a detector is not re-entrant: every block needs its own
  }
  if (frame_stride == 0) frame_stride = pitch * (size_t)h;
  std::vector<int> rcs(n_det, CTAG_OK);
  std::vector<std::string> errs(n_det);  // ctag_last_error() is per thread: carry the workers' texts back to the caller
  std::vector<std::thread> threads;
  // contiguous blocks, frame f -> detector floor(f * n_det / n) (video locality); one host thread per detector
  for (int g = 0; g < n_det; ++g) {
    const int first = (int)((long long)n * g / n_det), last = (int)((long long)n * (g + 1) / n_det);
    if (last <= first) continue;
    threads.emplace_back([=, &rcs, &errs]() {
      rcs[g] = ctag_detect_batch(dets[g], static_cast<const uint8_t*>(frames) + frame_stride * first, last - first, w, h, pitch,
                                 frame_stride, channels, 0, adaptive_thresh, corner_subpix, subpix_dist,
                                 out ? out + (size_t)first * cap_per_frame : nullptr, cap_per_frame, n_out ? n_out + first : nullptr,
                                 info ? info + first : nullptr);
      if (rcs[g] == CTAG_OK && out && n_out)
        for (int f = first; f < last; ++f)
          for (int k = 0; k < n_out[f] && k < cap_per_frame; ++k) out[(size_t)f * cap_per_frame + k].frame = f;
      if (rcs[g] != CTAG_OK) errs[g] = g_last_error;
    });
  }
  for (auto& t : threads) t.join();
  for (int g = 0; g < n_det; ++g)
    if (rcs[g] != CTAG_OK) {
      set_last_error_text(errs[g].c_str());
      return rcs[g];
    }
  return CTAG_OK;
}

int ctag_detect_batch_jpeg(ctag_detector* d, const uint8_t* const* jpeg, const size_t* jpeg_bytes, int n, int adaptive_thresh,
                           int corner_subpix, int subpix_dist, ctag_marker* out, int cap_per_frame, int* n_out, ctag_frame_info* info,
                           int* width_out, int* height_out) {
  if (!d || !jpeg || !jpeg_bytes || n <= 0) return CTAG_ERR_ARG;
  if (d->in_flight != 0) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  // the batch's frame size: from the SOF segment of the first frame (every frame is checked against it when its chunk
  // is parsed / decoded)
  int w = 0, hgt = 0;
  for (int f = 0; f < n; ++f)
    if (!jpeg[f] || jpeg_bytes[f] < 4) return CTAG_ERR_ARG;
  {
    const uint8_t* p = jpeg[0];
    const size_t len = jpeg_bytes[0];
    if (p[0] != 0xFF || p[1] != 0xD8) return CTAG_ERR_ARG;
    size_t i = 2;
    while (i + 9 < len && p[i] == 0xFF) {
      const int m = p[i + 1];
      if (m == 0xFF) { ++i; continue; }
      const size_t seg = ((size_t)p[i + 2] << 8) | p[i + 3];
      if (m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
        hgt = (p[i + 5] << 8) | p[i + 6];
        w = (p[i + 7] << 8) | p[i + 8];
        break;
      }
      if (m == 0xDA) break;
      i += 2 + seg;
    }
    if (w <= 0 || hgt <= 0) return CTAG_ERR_ARG;
  }
  if (width_out) *width_out = w;
  if (height_out) *height_out = hgt;
  int rc = check_args(d, jpeg, n, w, hgt, 3, adaptive_thresh, corner_subpix, subpix_dist);
  if (rc != CTAG_OK) return rc;
  rc = detect_batch_jpeg(d, jpeg, jpeg_bytes, n, w, hgt, adaptive_thresh, corner_subpix, subpix_dist, out, cap_per_frame, n_out, info);
  if (rc != CTAG_OK) drain_slots(d);
  return rc;
}

int ctag_jpeg_backend(const ctag_detector* d) { return d ? d->jpeg_backend_used : -1; }

int ctag_render_frames(ctag_detector* d, void* frames_dev, int n, int w, int h, size_t pitch, size_t frame_stride, int channels,
                       const float* marker_specs, const int* marker_start, const float* frame_params) {
  if (!d || !frames_dev || n <= 0 || w <= 0 || h <= 0 || (channels != 1 && channels != 3) || !marker_start || !frame_params) return CTAG_ERR_ARG;
  if (pitch < (size_t)w * channels) return CTAG_ERR_ARG;
  if (frame_stride == 0) frame_stride = pitch * (size_t)h;
  const int nm = marker_start[n];
  if (nm < 0 || (nm > 0 && !marker_specs)) return CTAG_ERR_ARG;
  for (int f = 0; f < n; ++f)
    if (marker_start[f + 1] < marker_start[f]) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  const size_t mb = render_marker_dev_bytes();
  std::vector<uint8_t> packed(mb * (size_t)(nm > 0 ? nm : 1));
  for (int f = 0; f < n; ++f)
    for (int m = marker_start[f]; m < marker_start[f + 1]; ++m) {
      const float* sp = marker_specs + 16 * (size_t)m;
      const int row = (int)sp[10];
      if (row < 0 || row >= d->rows) return CTAG_ERR_ARG;
      render_pack_marker(sp, d->cols, row * d->cols, frame_params + 8 * (size_t)f, w, h, packed.data() + mb * m);
    }
  float* d_img = nullptr;
  float* d_fp = nullptr;
  void* d_mk = nullptr;
  int* d_ms = nullptr;
  // frames are rendered in groups that keep the float scene buffer below 1 GiB
  int group = (int)(((size_t)1 << 30) / ((size_t)w * h * sizeof(float)));
  if (group < 1) group = 1;
  if (group > n) group = n;
  cudaError_t e = cudaMalloc(&d_img, (size_t)group * w * h * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_fp, sizeof(float) * 8 * n);
  if (e == cudaSuccess) e = cudaMalloc(&d_mk, packed.size());
  if (e == cudaSuccess) e = cudaMalloc(&d_ms, sizeof(int) * (n + 1));
  if (e == cudaSuccess) e = cudaMemcpy(d_fp, frame_params, sizeof(float) * 8 * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_mk, packed.data(), packed.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_ms, marker_start, sizeof(int) * (n + 1), cudaMemcpyHostToDevice);
  int rc = CTAG_OK;
  if (e == cudaSuccess) {
    for (int f0 = 0; f0 < n && rc == CTAG_OK; f0 += group) {
      const int c = n - f0 < group ? n - f0 : group;
      rc = launch_render(d_img, c, w, h, d_fp + 8 * f0, d_mk, d_ms + f0, d->d_state, static_cast<uint8_t*>(frames_dev) + frame_stride * f0,
                         pitch, frame_stride, channels, 0);
    }
    if (rc == CTAG_OK) e = cudaDeviceSynchronize();
  }
  cudaFree(d_img);
  cudaFree(d_fp);
  cudaFree(d_mk);
  cudaFree(d_ms);
  if (e != cudaSuccess) {
    set_last_error("ctag_render_frames", e, __FILE__, __LINE__);
    return CTAG_ERR_CUDA;
  }
  return rc;
}

int ctag_detect(ctag_detector* d, const uint8_t* gray, int w, int h, size_t pitch, int adaptive_thresh, int corner_subpix,
                int subpix_dist, ctag_marker* out, int cap, int* n_out, int* frame_status) {
  ctag_frame_info info;
  int n = 0;
  int rc = ctag_detect_batch(d, gray, 1, w, h, pitch, 0, 1, 0, adaptive_thresh, corner_subpix, subpix_dist, out, cap, &n,
                             &info);
  if (rc != CTAG_OK) return rc;
  if (n_out) *n_out = n;
  if (frame_status) *frame_status = info.status;
  return CTAG_OK;
}

int ctag_stage_time_ms(const ctag_detector* d, float* ms_out) {
  if (!d || !ms_out) return CTAG_ERR_ARG;
  for (int i = 0; i < CTAG_STAGE_COUNT; ++i) ms_out[i] = d->stage_ms[i];
  return CTAG_OK;
}

int ctag_stage_timeline_ms(const ctag_detector* d, float* ms_out) {
  if (!d || !ms_out) return CTAG_ERR_ARG;
  for (int i = 0; i <= CTAG_STAGE_COUNT; ++i) ms_out[i] = d->stage_stamp_ms[i];
  return CTAG_OK;
}

int ctag_last_launch_count(const ctag_detector* d) { return d ? d->last_launches : 0; }
int ctag_max_in_flight(void) { return kSlots; }
void* ctag_stream(const ctag_detector* d) { return d ? (void*)d->slot[0].stream : nullptr; }

// The debug getters address the most recently collected batch (for a chunked host call: its last chunk, so tests that
// use them pass batches of at most two frames or device frames).
static Slot* last_slot(ctag_detector* d, int frame) {
  if (!d || d->last < 0) return nullptr;
  Slot* s = &d->slot[d->last];
  return (frame >= 0 && frame < s->n) ? s : nullptr;
}

int ctag_debug_get_input(ctag_detector* d, int frame, uint8_t* out, size_t out_pitch) {
  Slot* s = last_slot(d, frame);
  if (!s || !out || !s->d_stage) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  const size_t row = (size_t)s->w * s->channels, dpitch = (size_t)round_up((int)row, 16);
  CTAG_CUDA_CHECK(cudaMemcpy2D(out, out_pitch, s->d_stage + dpitch * s->h * frame, dpitch, row, s->h, cudaMemcpyDeviceToHost));
  return CTAG_OK;
}

int ctag_debug_get_gray(ctag_detector* d, int frame, uint8_t* out, size_t out_pitch) {
  Slot* s = last_slot(d, frame);
  if (!s || !out || !s->gray) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  CTAG_CUDA_CHECK(cudaMemcpy2D(out, out_pitch, s->gray + s->gray_fs * frame, s->gray_pitch, s->w, s->h, cudaMemcpyDeviceToHost));
  return CTAG_OK;
}

int ctag_debug_get_binary(ctag_detector* d, int frame, uint8_t* out, size_t out_pitch) {
  Slot* s = last_slot(d, frame);
  if (!s || !out) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  CTAG_CUDA_CHECK(cudaMemcpy2D(out, out_pitch, s->d_bin + s->bin_fstride * frame, s->geo.bpitch, s->geo.hw, s->geo.hh,
                               cudaMemcpyDeviceToHost));
  return CTAG_OK;
}

int ctag_debug_get_quad_counters(ctag_detector* d, int32_t* out16) {
  Slot* s = last_slot(d, 0);
  if (!s || !out16) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  CTAG_CUDA_CHECK(cudaMemcpy(out16, s->d_qctl, sizeof(int32_t) * 16, cudaMemcpyDeviceToHost));
  return CTAG_OK;
}

int ctag_debug_get_components(ctag_detector* d, int frame, int32_t* out, int cap, int* n_out) {
  Slot* s = last_slot(d, frame);
  if (!s || !out || !n_out) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  int c[4];
  CTAG_CUDA_CHECK(cudaMemcpy(c, s->d_counters + 4 * frame, sizeof(c), cudaMemcpyDeviceToHost));
  int n = c[1] < cap ? c[1] : cap;
  CTAG_CUDA_CHECK(cudaMemcpy(out, s->d_legal + (size_t)frame * s->legal_cap * 6, sizeof(int) * 6 * n, cudaMemcpyDeviceToHost));
  *n_out = n;
  return CTAG_OK;
}

int ctag_debug_get_quads(ctag_detector* d, int frame, int32_t* comp_index, float* corners, int cap, int* n_out) {
  Slot* s = last_slot(d, frame);
  if (!s || !comp_index || !corners || !n_out) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  int nq = 0;
  CTAG_CUDA_CHECK(cudaMemcpy(&nq, s->d_n_quads + frame, sizeof(int), cudaMemcpyDeviceToHost));
  if (nq > kQuadCap) nq = kQuadCap;
  if (nq > cap) nq = cap;
  CTAG_CUDA_CHECK(cudaMemcpy(comp_index, s->d_quad_comp + (size_t)frame * kQuadCap, sizeof(int) * nq, cudaMemcpyDeviceToHost));
  CTAG_CUDA_CHECK(cudaMemcpy(corners, s->d_quads + (size_t)frame * kQuadCap * 8, sizeof(float) * 8 * nq, cudaMemcpyDeviceToHost));
  *n_out = nq;
  return CTAG_OK;
}

int ctag_debug_get_features(ctag_detector* d, int frame, float* corners, float* center, float* angle, int32_t* quad_pair,
                            int cap, int* n_out) {
  Slot* s = last_slot(d, frame);
  if (!s || !n_out) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  int fs[4];
  CTAG_CUDA_CHECK(cudaMemcpy(fs, s->d_fstate + 4 * frame, sizeof(fs), cudaMemcpyDeviceToHost));
  int nf = fs[1] < kFeatCap ? fs[1] : kFeatCap;
  if (nf > cap) nf = cap;
  struct Rec {
    float c[16];
    float cx, cy, angle;
    int qi, qj;
  };
  if (sizeof(Rec) != sizeof_feature_rec()) return CTAG_ERR_UNSUPPORTED;
  std::vector<Rec> recs(nf);
  if (nf)
    CTAG_CUDA_CHECK(cudaMemcpy(recs.data(), static_cast<const uint8_t*>(s->d_feats) + sizeof(Rec) * kFeatCap * frame,
                               sizeof(Rec) * nf, cudaMemcpyDeviceToHost));
  for (int i = 0; i < nf; ++i) {
    if (corners) memcpy(corners + 16 * i, recs[i].c, sizeof(float) * 16);
    if (center) center[2 * i] = recs[i].cx, center[2 * i + 1] = recs[i].cy;
    if (angle) angle[i] = recs[i].angle;
    if (quad_pair) quad_pair[2 * i] = recs[i].qi, quad_pair[2 * i + 1] = recs[i].qj;
  }
  *n_out = nf;
  return CTAG_OK;
}

const char* ctag_strerror(int code) {
  switch (code) {
    case CTAG_OK: return "ok";
    case CTAG_ERR_ARG: return "invalid argument";
    case CTAG_ERR_FILE: return "could not open the file";
    case CTAG_ERR_DICTIONARY: return "the number in state matrix must between 0 to 63";
    case CTAG_ERR_CUDA: return "CUDA error";
    case CTAG_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (no CPU fallback)";
    case CTAG_ERR_UNSUPPORTED: return "unsupported configuration";
    case CTAG_ERR_CAPACITY: return "internal capacity exceeded";
    case CTAG_ERR_ALIGNMENT: return "device input must be 16-byte aligned with pitch % 16 == 0";
    default: return "unknown error";
  }
}

const char* ctag_last_error(void) { return g_last_error; }
const char* ctag_version(void) { return "cylindertag_b200 0.1 sm_100a"; }

}  // extern "C"
