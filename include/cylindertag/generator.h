// generator.h -- dictionary ("codebook") generation for CylinderTag markers, header only, host code.
//
// The reference ships ONE dictionary (CTag_2f12c.marker); its MATLAB generator (CylinderTag_generator.m) builds others
// by search.  Rules:
//   * a state is 8 * left + right with digits 0..7, legal iff both digits lie in the same half (<= 3 or >= 4)
//     (CylinderTag_generator.m:18,96,114,164);
//   * inverse(s) = (7 - s % 8) * 8 + (7 - s / 8)  (:198 and corner_detector.cpp:1299);
//   * every cyclic window of `feature_size` states, read forward and read as flipped + inverted, is unique over the whole
//     dictionary, and no window equals an inverse reading of its own row (:247-286, :27,179).
// The .marker format is the one CylinderTag::load_from_file reads (CylinderTag.cpp:24-32):
//   `rows cols featureSize` then rows x cols integers.
#pragma once
#include <cstdint>
#include <fstream>
#include <set>
#include <string>
#include <vector>

#include "CylinderTag.h"

namespace ctag_api {

inline bool legal_state(int s) { return s >= 0 && s <= 63 && ((s / 8 <= 3) == (s % 8 <= 3)); }
inline int inverse_state(int s) { return (7 - s % 8) * 8 + (7 - s / 8); }

namespace detail {
// forward and inverse readings of every cyclic window of one row
inline void row_windows(const int32_t* row, int cols, int f, std::vector<std::vector<int>>& out) {
  for (int j = 0; j < cols; ++j) {
    std::vector<int> fw(f), inv(f);
    for (int k = 0; k < f; ++k) {
      fw[k] = row[(j + k) % cols];
      inv[k] = inverse_state(row[((j - k) % cols + cols) % cols]);
    }
    out.push_back(fw);
    out.push_back(inv);
  }
}
}  // namespace detail

// true iff every state is legal and every window reading is unique over the dictionary
inline bool check_codebook(const Mat1i& state, int feature_size) {
  if (state.rows <= 0 || state.cols <= 0 || feature_size <= 0 || (int)state.data.size() < state.rows * state.cols) return false;
  std::set<std::vector<int>> seen;
  for (int r = 0; r < state.rows; ++r) {
    const int32_t* row = &state.data[(size_t)r * state.cols];
    for (int c = 0; c < state.cols; ++c)
      if (!legal_state(row[c])) return false;
    std::vector<std::vector<int>> w;
    detail::row_windows(row, state.cols, feature_size, w);
    for (const auto& v : w)
      if (!seen.insert(v).second) return false;
  }
  return true;
}

// Random search (splitmix64 stream from `seed`): draws rows of legal states and keeps those whose windows are new.
// Returns an empty matrix when `rows` rows could not be found within `max_attempts` draws (a random search fills about
// half of the window space: 2-state windows over 32 legal states give 1024 readings, 24 per 12-column row, so the
// shipped 41-row 2f12c dictionary needs the reference's exhaustive DFS, not this).
inline Mat1i generate_codebook(int cols, int feature_size, int rows, uint64_t seed = 7, int max_attempts = 200000) {
  std::vector<int> states;
  for (int s = 0; s < 64; ++s)
    if (legal_state(s)) states.push_back(s);
  uint64_t x = seed;
  auto next = [&x]() {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  };
  Mat1i out;
  out.cols = cols;
  std::set<std::vector<int>> used;
  for (int attempts = 0; out.rows < rows && attempts < max_attempts; ++attempts) {
    std::vector<int32_t> row(cols);
    for (int c = 0; c < cols; ++c) row[c] = states[next() % states.size()];
    std::vector<std::vector<int>> w;
    detail::row_windows(row.data(), cols, feature_size, w);
    std::set<std::vector<int>> own(w.begin(), w.end());
    if (own.size() != w.size()) continue;  // a window repeats inside the row or equals an inverse reading of it
    bool clash = false;
    for (const auto& v : w) clash = clash || used.count(v) != 0;
    if (clash) continue;
    used.insert(w.begin(), w.end());
    out.data.insert(out.data.end(), row.begin(), row.end());
    out.rows += 1;
  }
  if (out.rows < rows) return Mat1i();
  return out;
}

// .marker writer (the format of CylinderTag.cpp:24-32)
inline bool write_marker_file(const std::string& path, const Mat1i& state, int feature_size) {
  std::ofstream f(path);
  if (!f.is_open()) return false;
  f << state.rows << " " << state.cols << " " << feature_size << "\n";
  for (int r = 0; r < state.rows; ++r) {
    for (int c = 0; c < state.cols; ++c) f << (c ? "\t" : "") << state.data[(size_t)r * state.cols + c];
    f << "\n";
  }
  return (bool)f;
}

}  // namespace ctag_api
