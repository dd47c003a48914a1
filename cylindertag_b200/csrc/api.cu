// C ABI (include/ctag.h) over the sm_100a detection kernels: detector handle, workspace, batch pipeline.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <fstream>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

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

}  // namespace ctag

using namespace ctag;

struct ctag_detector {
  int device = 0;
  cudaStream_t stream = nullptr;
  // dictionary (header/CylinderTag.h:44-45)
  std::vector<int32_t> state;
  int rows = 0, cols = 0, feature_size = 0;
  int32_t* d_state = nullptr;

  // workspace, sized for (cap_frames, w, h, channels)
  int cap_frames = 0, w = 0, h = 0;
  FrameGeom geo{};
  uint8_t* d_stage = nullptr;  // staging for host inputs
  size_t stage_bytes = 0;
  uint8_t* d_gray = nullptr;   // [cap_frames][h][gpitch]   (BGR input only)
  uint8_t* d_bin = nullptr;    // [cap_frames][hh][bpitch]
  size_t gray_fstride = 0, bin_fstride = 0;
  // CCL (a4)
  int *d_labels = nullptr, *d_st_area = nullptr, *d_st_x0 = nullptr, *d_st_y0 = nullptr, *d_st_x1 = nullptr,
      *d_st_y1 = nullptr, *d_roots_tmp = nullptr, *d_span_count = nullptr, *d_legal = nullptr, *d_counters = nullptr;
  int legal_cap = 0, spans = 0;
  // quads (a5)
  int *d_prefix = nullptr, *d_work_counter = nullptr, *d_quad_status = nullptr, *d_quad_comp = nullptr,
      *d_n_quads = nullptr;
  float *d_quad_corners = nullptr, *d_quads = nullptr;
  uint8_t* d_quad_scratch = nullptr;
  int edge_warps = 0, exact_ctas = 0, sms = 0, fit_cap = 0, pool_cap = 0;
  void *d_fits = nullptr, *d_traj = nullptr, *d_fit_results = nullptr;
  int* d_exact_list = nullptr;
  uint16_t* d_pick_table = nullptr;  // per detector, independent of the frame geometry
  static constexpr int kPickTableMax = 1024;
  int* d_pool = nullptr;
  float* d_lines = nullptr;
  static constexpr int kQuadCap = CTAG_MAX_FRAME_QUADS;
  // features / markers (a6-a10)
  void *d_geom = nullptr, *d_feats = nullptr;
  int *d_fstate = nullptr, *d_packed_count = nullptr, *d_summary = nullptr;
  ctag_marker *d_markers = nullptr, *d_packed = nullptr;
  int* h_summary = nullptr;         // pinned
  ctag_marker* h_packed = nullptr;  // pinned
  static constexpr int kFeatCap = CTAG_MAX_FRAME_FEATURES;
  static constexpr int kMarkerCap = CTAG_MAX_FRAME_FEATURES / 2;
  int cur_subpix = 0;

  // state of the batch in flight / last batch
  int cur_n = 0, cur_channels = 0;
  const uint8_t* cur_gray = nullptr;  // full-res gray of the batch (input itself when channels == 1)
  size_t cur_gray_pitch = 0, cur_gray_fstride = 0;
  int launches = 0;
  float stage_ms[CTAG_STAGE_COUNT] = {0, 0, 0, 0, 0};
  cudaEvent_t ev[CTAG_STAGE_COUNT + 1] = {};
  bool in_flight = false;
};

static int select_device(int cuda_device, int* chosen) {
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
  return CTAG_OK;
}

static void free_workspace(ctag_detector* d) {
  // d_stage is managed separately (ensure_stage): it may hold the batch that is about to be processed
  void* ptrs[] = {d->d_gray, d->d_bin, d->d_labels, d->d_st_area, d->d_st_x0, d->d_st_y0, d->d_st_x1, d->d_st_y1,
                  d->d_roots_tmp, d->d_span_count, d->d_legal, d->d_counters, d->d_prefix, d->d_work_counter,
                  d->d_quad_status, d->d_quad_comp, d->d_n_quads, d->d_quad_corners, d->d_quads, d->d_quad_scratch,
                  d->d_geom, d->d_feats, d->d_fstate, d->d_packed_count, d->d_summary, d->d_markers, d->d_packed,
                  d->d_fits, d->d_traj, d->d_pool, d->d_lines, d->d_fit_results, d->d_exact_list};
  for (void* p : ptrs) cudaFree(p);
  cudaFreeHost(d->h_summary);
  cudaFreeHost(d->h_packed);
  d->h_summary = nullptr;
  d->h_packed = nullptr;
  d->d_geom = d->d_feats = d->d_fits = d->d_traj = d->d_fit_results = nullptr;
  d->d_exact_list = nullptr;
  d->d_pool = nullptr;
  d->d_lines = nullptr;
  d->d_fstate = d->d_packed_count = d->d_summary = nullptr;
  d->d_markers = d->d_packed = nullptr;
  d->d_gray = d->d_bin = d->d_quad_scratch = nullptr;
  d->d_labels = d->d_st_area = d->d_st_x0 = d->d_st_y0 = d->d_st_x1 = d->d_st_y1 = d->d_roots_tmp = d->d_span_count =
      d->d_legal = d->d_counters = d->d_prefix = d->d_work_counter = d->d_quad_status = d->d_quad_comp = d->d_n_quads =
          nullptr;
  d->d_quad_corners = d->d_quads = nullptr;
  d->cap_frames = 0;
}

template <typename T>
static cudaError_t dev_alloc(T** p, size_t count) {
  return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}

static int ensure_workspace(ctag_detector* d, int n, int w, int h) {
  if (d->cap_frames >= n && d->w == w && d->h == h) return CTAG_OK;
  int cap = n > d->cap_frames || d->w != w || d->h != h ? n : d->cap_frames;
  free_workspace(d);
  d->geo = make_geom(w, h);
  d->w = w;
  d->h = h;
  const FrameGeom& g = d->geo;
  d->gray_fstride = (size_t)g.gpitch * g.h;
  d->bin_fstride = (size_t)g.bpitch * g.hh;
  CTAG_CUDA_CHECK(cudaMalloc(&d->d_gray, d->gray_fstride * cap));
  CTAG_CUDA_CHECK(cudaMalloc(&d->d_bin, d->bin_fstride * cap));
  const size_t nb = (size_t)g.nblocks * cap;
  d->spans = (g.nblocks + 1023) / 1024;
  d->legal_cap = g.hw * g.hh / kAreaMin + 1;  // every legal component has >= 30 pixels
  if (d->legal_cap > 65536) d->legal_cap = 65536;
  CTAG_CUDA_CHECK(dev_alloc(&d->d_labels, nb));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_st_area, nb));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_st_x0, nb));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_st_y0, nb));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_st_x1, nb));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_st_y1, nb));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_roots_tmp, nb));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_span_count, (size_t)d->spans * cap));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_legal, (size_t)d->legal_cap * 6 * cap));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_counters, (size_t)4 * cap));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_prefix, (size_t)cap + 1));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_work_counter, 8));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_quad_status, (size_t)d->legal_cap * cap));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_quad_corners, (size_t)d->legal_cap * 8 * cap));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_quads, (size_t)ctag_detector::kQuadCap * 8 * cap));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_quad_comp, (size_t)ctag_detector::kQuadCap * cap));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_n_quads, (size_t)cap));
  CTAG_CUDA_CHECK(cudaMalloc(&d->d_geom, sizeof_quad_geom() * ctag_detector::kQuadCap * cap));
  CTAG_CUDA_CHECK(cudaMalloc(&d->d_feats, sizeof_feature_rec() * ctag_detector::kFeatCap * cap));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_fstate, (size_t)4 * cap));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_packed_count, 4));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_summary, (size_t)12 * cap));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_markers, (size_t)ctag_detector::kMarkerCap * cap));
  CTAG_CUDA_CHECK(dev_alloc(&d->d_packed, (size_t)ctag_detector::kMarkerCap * cap));
  CTAG_CUDA_CHECK(cudaMallocHost(&d->h_summary, sizeof(int) * 12 * cap));
  CTAG_CUDA_CHECK(cudaMallocHost(&d->h_packed, sizeof(ctag_marker) * ctag_detector::kMarkerCap * cap));
  {
    cudaDeviceProp prop;
    CTAG_CUDA_CHECK(cudaGetDeviceProperties(&prop, d->device));
    d->sms = prop.multiProcessorCount;
    d->edge_warps = quad_edge_warps(d->sms);
    d->exact_ctas = quad_exact_ctas(d->sms);
    CTAG_CUDA_CHECK(cudaMalloc(&d->d_quad_scratch, quad_scratch_bytes_per_warp(g) * d->edge_warps));
    CTAG_CUDA_CHECK(cudaMalloc(&d->d_traj, quad_traj_bytes_per_cta() * d->exact_ctas));
    // components that reach four edges / their cluster points: the reference caps a frame at 1000 quads
    // (isVisited[1000]), so 1024 four-edge components per frame is already past its envelope; overflow drops the
    // component and flags the frame instead of writing out of bounds
    d->fit_cap = cap * (d->legal_cap < 1024 ? d->legal_cap : 1024);
    CTAG_CUDA_CHECK(cudaMalloc(&d->d_fit_results, quad_fitresult_bytes() * 80 * (size_t)d->fit_cap));
    CTAG_CUDA_CHECK(dev_alloc(&d->d_exact_list, (size_t)4 * d->fit_cap));
    const long long per_frame_pts = (long long)g.hw * g.hh < 262144 ? (long long)g.hw * g.hh : 262144;
    d->pool_cap = (int)(per_frame_pts * cap < 0x7fffffff ? per_frame_pts * cap : 0x7fffffff);
    CTAG_CUDA_CHECK(cudaMalloc(&d->d_fits, quad_fitrec_bytes() * d->fit_cap));
    CTAG_CUDA_CHECK(dev_alloc(&d->d_lines, (size_t)16 * d->fit_cap));
    CTAG_CUDA_CHECK(dev_alloc(&d->d_pool, (size_t)d->pool_cap));
  }
  d->cap_frames = cap;
  return CTAG_OK;
}

static int ensure_stage(ctag_detector* d, size_t bytes) {
  if (d->stage_bytes >= bytes) return CTAG_OK;
  cudaFree(d->d_stage);
  d->d_stage = nullptr;
  d->stage_bytes = 0;
  CTAG_CUDA_CHECK(cudaMalloc(&d->d_stage, bytes));
  d->stage_bytes = bytes;
  return CTAG_OK;
}

extern "C" {

int ctag_create(ctag_detector** out, const int32_t* state, int rows, int cols, int feature_size, int cuda_device) {
  if (!out || !state || rows <= 0 || cols <= 0 || feature_size <= 0) return CTAG_ERR_ARG;
  *out = nullptr;
  for (int i = 0; i < rows * cols; ++i)
    if (!(state[i] >= 0 && state[i] <= 63)) return CTAG_ERR_DICTIONARY;  // check_dictionary, CylinderTag.cpp:56-65
  int dev = 0;
  int rc = select_device(cuda_device, &dev);
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaSetDevice(dev));
  ctag_detector* d = new ctag_detector();
  d->device = dev;
  d->state.assign(state, state + rows * cols);
  d->rows = rows;
  d->cols = cols;
  d->feature_size = feature_size;
  if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&d->d_state, sizeof(int32_t) * rows * cols) != cudaSuccess ||
      cudaMemcpy(d->d_state, state, sizeof(int32_t) * rows * cols, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_last_error("detector setup", cudaGetLastError(), __FILE__, __LINE__);
    ctag_destroy(d);
    return CTAG_ERR_CUDA;
  }
  for (auto& e : d->ev) cudaEventCreate(&e);
  {
    // initial subsets of cv::fitLine's 20 restarts for every point count up to kPickTableMax (fit_core.cuh)
    std::vector<uint16_t> table((size_t)ctag_detector::kPickTableMax * 200);
    quad_build_pick_table(table.data(), ctag_detector::kPickTableMax);
    if (cudaMalloc(&d->d_pick_table, table.size() * sizeof(uint16_t)) != cudaSuccess ||
        cudaMemcpy(d->d_pick_table, table.data(), table.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_last_error("pick table", cudaGetLastError(), __FILE__, __LINE__);
      ctag_destroy(d);
      return CTAG_ERR_CUDA;
    }
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
  if (d->stream) cudaStreamSynchronize(d->stream);
  free_workspace(d);
  cudaFree(d->d_stage);
  cudaFree(d->d_state);
  cudaFree(d->d_pick_table);
  for (auto& e : d->ev)
    if (e) cudaEventDestroy(e);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
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
  if (!d || !frames_dev || n <= 0 || w <= 0 || h <= 0) return CTAG_ERR_ARG;
  if ((w & 1) || (h & 1)) return CTAG_ERR_ARG;  // exact 2x decimation needs even sizes
  if (channels != 1 && channels != 3) return CTAG_ERR_ARG;
  if (adaptive_thresh != kWin) return CTAG_ERR_UNSUPPORTED;
  if (d->in_flight) return CTAG_ERR_ARG;
  if (corner_subpix && (subpix_dist < 0 || subpix_dist > 64)) return CTAG_ERR_ARG;
  if (decode_smem_bytes(d->rows, d->cols) > 200 * 1024) return CTAG_ERR_UNSUPPORTED;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  if (frame_stride == 0) frame_stride = pitch * (size_t)h;
  int rc = ensure_workspace(d, n, w, h);
  if (rc != CTAG_OK) return rc;
  d->cur_n = n;
  d->cur_channels = channels;
  d->launches = 0;
  if (channels == 1) {
    d->cur_gray = static_cast<const uint8_t*>(frames_dev);
    d->cur_gray_pitch = pitch;
    d->cur_gray_fstride = frame_stride;
  } else {
    d->cur_gray = d->d_gray;
    d->cur_gray_pitch = d->geo.gpitch;
    d->cur_gray_fstride = d->gray_fstride;
  }
  CTAG_CUDA_CHECK(cudaEventRecord(d->ev[0], d->stream));
  rc = launch_front(frames_dev, n, d->geo, channels, pitch, frame_stride, d->d_gray, d->gray_fstride, d->d_bin,
                    d->bin_fstride, d->stream);
  if (rc != CTAG_OK) return rc;
  d->launches += 1;
  CTAG_CUDA_CHECK(cudaEventRecord(d->ev[1], d->stream));
  rc = launch_ccl(d->d_bin, d->bin_fstride, n, d->geo, d->d_labels, d->d_st_area, d->d_st_x0, d->d_st_y0, d->d_st_x1,
                  d->d_st_y1, d->d_roots_tmp, d->d_span_count, d->d_legal, d->legal_cap, d->d_counters, d->stream,
                  &d->launches);
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaEventRecord(d->ev[2], d->stream));
  rc = launch_quad(n, d->geo, d->d_bin, d->bin_fstride, d->d_labels, d->d_legal, d->legal_cap, d->d_counters, d->d_prefix,
                   d->d_work_counter, d->d_quad_scratch, d->edge_warps, d->d_fits, d->fit_cap, d->d_pool, d->pool_cap,
                   d->d_pick_table, ctag_detector::kPickTableMax, d->d_fit_results, d->d_exact_list, d->d_traj,
                   d->exact_ctas, d->sms, d->d_lines, d->d_quad_status, d->d_quad_corners, ctag_detector::kQuadCap,
                   d->d_quads, d->d_quad_comp, d->d_n_quads, d->stream, &d->launches);
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaEventRecord(d->ev[3], d->stream));
  rc = launch_features(n, d->geo, d->d_quads, d->d_n_quads, ctag_detector::kQuadCap, d->d_geom, d->d_feats,
                       ctag_detector::kFeatCap, d->feature_size, d->d_fstate, d->cur_gray, d->cur_gray_pitch,
                       d->cur_gray_fstride, corner_subpix, subpix_dist, d->stream, &d->launches);
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaEventRecord(d->ev[4], d->stream));
  rc = launch_decode(n, d->d_feats, ctag_detector::kFeatCap, d->d_fstate, d->d_state, d->rows, d->cols, d->feature_size,
                     d->d_markers, ctag_detector::kMarkerCap, d->d_counters, d->d_n_quads, ctag_detector::kQuadCap,
                     d->d_work_counter + 4 /* QC_OVERFLOW */, d->d_packed, d->d_packed_count, d->d_summary, d->stream, &d->launches);
  if (rc != CTAG_OK) return rc;
  CTAG_CUDA_CHECK(cudaEventRecord(d->ev[5], d->stream));
  CTAG_CUDA_CHECK(cudaMemcpyAsync(d->h_summary, d->d_summary, sizeof(int) * 12 * n, cudaMemcpyDeviceToHost, d->stream));
  d->cur_subpix = corner_subpix;
  d->in_flight = true;
  return CTAG_OK;
}

int ctag_detect_batch_collect(ctag_detector* d, ctag_marker* out, int cap_per_frame, int* n_out, ctag_frame_info* info) {
  if (!d || !d->in_flight) return CTAG_ERR_ARG;
  (void)out;
  (void)cap_per_frame;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  d->in_flight = false;
  CTAG_CUDA_CHECK(cudaStreamSynchronize(d->stream));
  for (int sidx = 0; sidx < CTAG_STAGE_COUNT; ++sidx) cudaEventElapsedTime(&d->stage_ms[sidx], d->ev[sidx], d->ev[sidx + 1]);
  int total = 0;
  for (int f = 0; f < d->cur_n; ++f) total += d->h_summary[12 * f + 10];
  if (total > 0) {
    CTAG_CUDA_CHECK(cudaMemcpyAsync(d->h_packed, d->d_packed, sizeof(ctag_marker) * total, cudaMemcpyDeviceToHost, d->stream));
    CTAG_CUDA_CHECK(cudaStreamSynchronize(d->stream));
  }
  for (int f = 0; f < d->cur_n; ++f) {
    const int* sm = d->h_summary + 12 * f;
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
      if (ncopy > 0) memcpy(out + (size_t)f * cap_per_frame, d->h_packed + sm[9], sizeof(ctag_marker) * ncopy);
    }
  }
  return CTAG_OK;
}

int ctag_detect_batch(ctag_detector* d, const void* frames, int n, int w, int h, size_t pitch, size_t frame_stride,
                      int channels, int is_device, int adaptive_thresh, int corner_subpix, int subpix_dist,
                      ctag_marker* out, int cap_per_frame, int* n_out, ctag_frame_info* info) {
  if (!d || !frames || n <= 0 || w <= 0 || h <= 0) return CTAG_ERR_ARG;
  if (channels != 1 && channels != 3) return CTAG_ERR_ARG;
  const void* src = frames;
  if (frame_stride == 0) frame_stride = pitch * (size_t)h;
  if (!is_device) {
    CTAG_CUDA_CHECK(cudaSetDevice(d->device));
    // host frames: copy into a 16B-pitched device staging buffer (2-D copy normalises any host pitch)
    size_t dpitch = (size_t)round_up(w * channels, 16);
    size_t dfs = dpitch * h;
    int rc = ensure_stage(d, dfs * n);
    if (rc != CTAG_OK) return rc;
    if (pitch < (size_t)w * channels) return CTAG_ERR_ARG;
    for (int f = 0; f < n; ++f)
      CTAG_CUDA_CHECK(cudaMemcpy2DAsync(d->d_stage + dfs * f, dpitch, static_cast<const uint8_t*>(frames) + frame_stride * f,
                                        pitch, (size_t)w * channels, h, cudaMemcpyHostToDevice, d->stream));
    src = d->d_stage;
    pitch = dpitch;
    frame_stride = dfs;
  }
  int rc = ctag_detect_batch_enqueue(d, src, n, w, h, pitch, frame_stride, channels, adaptive_thresh, corner_subpix,
                                     subpix_dist);
  if (rc != CTAG_OK) return rc;
  return ctag_detect_batch_collect(d, out, cap_per_frame, n_out, info);
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

int ctag_last_launch_count(const ctag_detector* d) { return d ? d->launches : 0; }
void* ctag_stream(const ctag_detector* d) { return d ? (void*)d->stream : nullptr; }

int ctag_debug_get_gray(ctag_detector* d, int frame, uint8_t* out, size_t out_pitch) {
  if (!d || !out || frame < 0 || frame >= d->cur_n || !d->cur_gray) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  CTAG_CUDA_CHECK(cudaMemcpy2D(out, out_pitch, d->cur_gray + d->cur_gray_fstride * frame, d->cur_gray_pitch, d->w, d->h,
                               cudaMemcpyDeviceToHost));
  return CTAG_OK;
}

int ctag_debug_get_binary(ctag_detector* d, int frame, uint8_t* out, size_t out_pitch) {
  if (!d || !out || frame < 0 || frame >= d->cur_n) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  CTAG_CUDA_CHECK(cudaMemcpy2D(out, out_pitch, d->d_bin + d->bin_fstride * frame, d->geo.bpitch, d->geo.hw, d->geo.hh,
                               cudaMemcpyDeviceToHost));
  return CTAG_OK;
}

int ctag_debug_get_components(ctag_detector* d, int frame, int32_t* out, int cap, int* n_out) {
  if (!d || !out || !n_out || frame < 0 || frame >= d->cur_n) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  int c[4];
  CTAG_CUDA_CHECK(cudaMemcpy(c, d->d_counters + 4 * frame, sizeof(c), cudaMemcpyDeviceToHost));
  int n = c[1] < cap ? c[1] : cap;
  CTAG_CUDA_CHECK(cudaMemcpy(out, d->d_legal + (size_t)frame * d->legal_cap * 6, sizeof(int) * 6 * n, cudaMemcpyDeviceToHost));
  *n_out = n;
  return CTAG_OK;
}
int ctag_debug_get_quads(ctag_detector* d, int frame, int32_t* comp_index, float* corners, int cap, int* n_out) {
  if (!d || !comp_index || !corners || !n_out || frame < 0 || frame >= d->cur_n) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  int nq = 0;
  CTAG_CUDA_CHECK(cudaMemcpy(&nq, d->d_n_quads + frame, sizeof(int), cudaMemcpyDeviceToHost));
  if (nq > ctag_detector::kQuadCap) nq = ctag_detector::kQuadCap;
  if (nq > cap) nq = cap;
  CTAG_CUDA_CHECK(cudaMemcpy(comp_index, d->d_quad_comp + (size_t)frame * ctag_detector::kQuadCap, sizeof(int) * nq,
                             cudaMemcpyDeviceToHost));
  CTAG_CUDA_CHECK(cudaMemcpy(corners, d->d_quads + (size_t)frame * ctag_detector::kQuadCap * 8, sizeof(float) * 8 * nq,
                             cudaMemcpyDeviceToHost));
  *n_out = nq;
  return CTAG_OK;
}
int ctag_debug_get_features(ctag_detector* d, int frame, float* corners, float* center, float* angle, int32_t* quad_pair,
                            int cap, int* n_out) {
  if (!d || !n_out || frame < 0 || frame >= d->cur_n) return CTAG_ERR_ARG;
  CTAG_CUDA_CHECK(cudaSetDevice(d->device));
  int fs[4];
  CTAG_CUDA_CHECK(cudaMemcpy(fs, d->d_fstate + 4 * frame, sizeof(fs), cudaMemcpyDeviceToHost));
  int nf = fs[1] < ctag_detector::kFeatCap ? fs[1] : ctag_detector::kFeatCap;
  if (nf > cap) nf = cap;
  struct Rec {
    float c[16];
    float cx, cy, angle;
    int qi, qj;
  };
  if (sizeof(Rec) != sizeof_feature_rec()) return CTAG_ERR_UNSUPPORTED;
  std::vector<Rec> recs(nf);
  if (nf)
    CTAG_CUDA_CHECK(cudaMemcpy(recs.data(), static_cast<const uint8_t*>(d->d_feats) + sizeof(Rec) * ctag_detector::kFeatCap * frame,
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
