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
#include <algorithm>
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

// ---- depth-first search of the reference's generator (CylinderTag_generator.m:34-216) ---------------------------------
// A window of `f` consecutive states is a number in [0, 64^f); `used` marks every window taken by a row, read forwards or as
// its inverse (:247-286).  A row grows one state at a time (:62-159): the candidates for the next state are the legal states
// whose new window and its inverse are both free (:97-112); they are tried in DESCENDING order of how many continuations
// each would leave (:113-139, ties in random order), backtracking when a branch dies; the last state closes the cycle and
// is drawn at random, at most 100 times, checking all f wrap-around windows (:160-190).  Rows are added until `rows` are
// found or the search budget is spent; a book that stalls is thrown away and started again from the next random stream
// (the MATLAB script is simply re-run by hand in that case).  Capacity (:36-39): the windows that are legal and differ from
// their own inverse, divided by 2 * cols -- 41 rows for two-state windows on twelve columns, the size of the shipped book.
namespace detail {
struct DfsGen {
  int cols, f, rows_wanted;
  std::vector<int> legal;
  std::vector<uint8_t> used;
  std::vector<int64_t> pw;  // 64^k
  uint64_t x;
  long long nodes = 0, node_budget = 0;
  std::vector<int32_t> book;
  int rows = 0;

  uint64_t next() {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  // window code of states w[0..f-1] (first state in the lowest digit, as in the .m file) and of its inverse reading
  int64_t code(const int* w) const {
    int64_t c = 0;
    for (int k = 0; k < f; ++k) c += (int64_t)w[k] * pw[k];
    return c;
  }
  int64_t inv_code(const int* w) const {
    int64_t c = 0;
    for (int k = 0; k < f; ++k) c += (int64_t)inverse_state(w[k]) * pw[f - 1 - k];
    return c;
  }
  bool free_window(const int* w) const {
    const int64_t a = code(w), b = inv_code(w);
    return a != b && !used[a] && !used[b];
  }
  void mark(const int* w, uint8_t v) {
    used[code(w)] = v;
    used[inv_code(w)] = v;
  }
  // one row by depth-first search; `row` holds the states placed so far
  bool dfs(std::vector<int>& row) {
    if (++nodes > node_budget) return false;
    const int n = (int)row.size();
    if (n == cols) return true;
    std::vector<int> w(f);
    if (n < f) {  // the first window: random free start (:69-91)
      for (int tries = 0; tries < 200; ++tries) {
        std::vector<int> start(f);
        for (int k = 0; k < f; ++k) start[k] = legal[next() % legal.size()];
        if (!free_window(start.data())) continue;
        mark(start.data(), 1);
        row = start;
        if (dfs(row)) return true;
        mark(start.data(), 0);
        row.clear();
        if (nodes > node_budget) return false;
      }
      return false;
    }
    if (n < cols - 1) {
      struct Cand { int s, options; uint64_t tie; };
      std::vector<Cand> cands;
      for (int s : legal) {
        for (int k = 0; k < f - 1; ++k) w[k] = row[n - f + 1 + k];
        w[f - 1] = s;
        if (!free_window(w.data())) continue;
        int options = 0;  // continuations one step further (:113-129)
        std::vector<int> w2(f);
        for (int k = 0; k < f - 2; ++k) w2[k] = row[n - f + 2 + k];
        if (f >= 2) w2[f - 2] = s;
        for (int t : legal) {
          w2[f - 1] = t;
          if (free_window(w2.data()) && !(code(w2.data()) == code(w.data()) || code(w2.data()) == inv_code(w.data()))) ++options;
        }
        cands.push_back({s, options, next()});
      }
      std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.options != b.options ? a.options > b.options : a.tie < b.tie; });
      for (const Cand& c : cands) {
        if (c.options == 0) break;  // (:141-143)
        for (int k = 0; k < f - 1; ++k) w[k] = row[n - f + 1 + k];
        w[f - 1] = c.s;
        mark(w.data(), 1);
        row.push_back(c.s);
        if (dfs(row)) return true;
        row.pop_back();
        for (int k = 0; k < f - 1; ++k) w[k] = row[n - f + 1 + k];
        w[f - 1] = c.s;
        mark(w.data(), 0);
        if (nodes > node_budget) return false;
      }
      return false;
    }
    // the last state closes the cycle: f windows wrap around (:160-190); every legal state is tried, in random order
    std::vector<int> order = legal;
    for (size_t i = order.size(); i > 1; --i) std::swap(order[i - 1], order[next() % i]);
    for (int s : order) {
      row.push_back(s);
      std::vector<int64_t> codes;
      bool ok = true;
      for (int j = 0; j < f && ok; ++j) {  // windows starting at positions cols - f + j
        for (int k = 0; k < f; ++k) w[k] = row[(cols - f + j + k) % cols];
        ok = free_window(w.data());
        const int64_t a = code(w.data()), b = inv_code(w.data());
        for (int64_t c : codes) ok = ok && c != a && c != b;
        codes.push_back(a);
        codes.push_back(b);
      }
      if (ok) {
        for (int64_t c : codes) used[c] = 1;
        return true;
      }
      row.pop_back();
    }
    return false;
  }
};
}  // namespace detail

inline int codebook_capacity(int cols, int feature_size) {
  detail::DfsGen g;
  g.f = feature_size;
  g.pw.assign(feature_size + 1, 1);
  for (int k = 1; k <= feature_size; ++k) g.pw[k] = g.pw[k - 1] * 64;
  long long usable = 0;
  std::vector<int> w(feature_size, 0), legal;
  for (int s = 0; s < 64; ++s)
    if (legal_state(s)) legal.push_back(s);
  std::vector<int> idx(feature_size, 0);
  while (true) {
    for (int k = 0; k < feature_size; ++k) w[k] = legal[idx[k]];
    if (g.code(w.data()) != g.inv_code(w.data())) ++usable;
    int k = 0;
    while (k < feature_size && ++idx[k] == (int)legal.size()) idx[k++] = 0;
    if (k == feature_size) break;
  }
  return (int)(usable / (2 * cols));
}

// The reference's search.  Returns the book (possibly with fewer than `rows` rows if the budget ran out; rows is clamped
// to the capacity).  `restarts`: books thrown away and begun again before giving up.
inline Mat1i generate_codebook_dfs(int cols, int feature_size, int rows, uint64_t seed = 7, int restarts = 200,
                                   long long nodes_per_row = 200000) {
  Mat1i best;
  best.cols = cols;
  if (cols < feature_size + 1 || feature_size < 1 || feature_size > 4) return best;
  const int cap = codebook_capacity(cols, feature_size);
  if (rows > cap) rows = cap;
  for (int attempt = 0; attempt <= restarts; ++attempt) {
    detail::DfsGen g;
    g.cols = cols;
    g.f = feature_size;
    g.rows_wanted = rows;
    for (int s = 0; s < 64; ++s)
      if (legal_state(s)) g.legal.push_back(s);
    g.pw.assign(feature_size + 1, 1);
    for (int k = 1; k <= feature_size; ++k) g.pw[k] = g.pw[k - 1] * 64;
    g.used.assign((size_t)g.pw[feature_size], 0);
    g.x = seed + 0x632BE59BD9B4E019ull * (uint64_t)attempt;
    int failures = 0;
    while (g.rows < rows && failures < 40) {
      std::vector<int> row;
      g.nodes = 0;
      g.node_budget = nodes_per_row;
      if (g.dfs(row)) {
        g.book.insert(g.book.end(), row.begin(), row.end());
        g.rows += 1;
        failures = 0;
      } else {
        // undo whatever the failed search still holds (dfs unmarks on the way back; a budget stop may leave marks)
        std::fill(g.used.begin(), g.used.end(), 0);
        for (int r = 0; r < g.rows; ++r) {
          std::vector<int> w(feature_size);
          for (int j = 0; j < cols; ++j) {
            for (int k = 0; k < feature_size; ++k) w[k] = g.book[(size_t)r * cols + (j + k) % cols];
            g.mark(w.data(), 1);
          }
        }
        ++failures;
      }
    }
    if (g.rows > best.rows) {
      best.rows = g.rows;
      best.data = g.book;
    }
    if (best.rows >= rows) break;
  }
  return best;
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
