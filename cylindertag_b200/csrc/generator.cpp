// C ABI over the dictionary generator of include/cylindertag/generator.h (host code, no GPU work): the
// depth-first search of CylinderTag_generator.m:34-216 and the uniqueness check of :247-286.
#include "../../include/ctag.h"
#include "../../include/cylindertag/generator.h"

extern "C" {

int ctag_codebook_capacity(int cols, int feature_size) {
  if (cols < 2 || feature_size < 2 || feature_size > 4 || cols <= feature_size) return CTAG_ERR_ARG;
  return ctag_api::codebook_capacity(cols, feature_size);
}

int ctag_generate_codebook(int cols, int feature_size, int rows, uint64_t seed, int32_t* state_out, int cap_rows, int* rows_out) {
  if (!state_out || !rows_out || rows <= 0 || cap_rows < rows || cols < 2 || feature_size < 2 || feature_size > 4 || cols <= feature_size)
    return CTAG_ERR_ARG;
  ctag_api::Mat1i cb = ctag_api::generate_codebook_dfs(cols, feature_size, rows, seed);
  *rows_out = cb.rows;
  for (size_t i = 0; i < cb.data.size(); ++i) state_out[i] = cb.data[i];
  return CTAG_OK;
}

int ctag_check_codebook(const int32_t* state, int rows, int cols, int feature_size) {
  if (!state || rows <= 0 || cols <= 0 || feature_size <= 0) return CTAG_ERR_ARG;
  ctag_api::Mat1i m;
  m.rows = rows;
  m.cols = cols;
  m.data.assign(state, state + (size_t)rows * cols);
  return ctag_api::check_codebook(m, feature_size) ? 1 : 0;
}
}
