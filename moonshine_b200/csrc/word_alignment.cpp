#include "word_alignment.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace msb {

namespace {

// Monotonic alignment of rows (decode steps) to columns (encoder frames) by dynamic time warping over
// cost[n][m]; ties prefer the diagonal, then "row advances", then "column advances".
void dtw_path(const std::vector<float>& cost, int n, int m, std::vector<int>& rows, std::vector<int>& cols) {
  const float inf = std::numeric_limits<float>::infinity();
  std::vector<float> acc((size_t)(n + 1) * (m + 1), inf);
  std::vector<uint8_t> from((size_t)n * m, 0);
  acc[0] = 0.f;
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < m; j++) {
      const float diag = acc[(size_t)i * (m + 1) + j];
      const float up = acc[(size_t)i * (m + 1) + j + 1];
      const float left = acc[(size_t)(i + 1) * (m + 1) + j];
      uint8_t pick;
      float best;
      if (diag <= up && diag <= left) { pick = 0; best = diag; }
      else if (up <= diag && up <= left) { pick = 1; best = up; }
      else { pick = 2; best = left; }
      from[(size_t)i * m + j] = pick;
      acc[(size_t)(i + 1) * (m + 1) + j + 1] = cost[(size_t)i * m + j] + best;
    }
  }
  rows.clear();
  cols.clear();
  int i = n - 1, j = m - 1;
  while (i >= 0 || j >= 0) {
    rows.push_back(i);
    cols.push_back(j);
    if (i == 0 && j == 0) break;
    const uint8_t pick = from[(size_t)i * m + j];
    if (pick == 0) { i--; j--; }
    else if (pick == 1) i--;
    else j--;
  }
  std::reverse(rows.begin(), rows.end());
  std::reverse(cols.begin(), cols.end());
}

// width-7 median along a row with reflect padding (index clamped when the row is shorter than the pad)
void median7_row(const float* src, int width, float* dst, std::vector<float>& padded) {
  const int fw = 7, pad = fw / 2;
  padded.resize((size_t)width + 2 * pad);
  for (int q = 0; q < pad; q++) padded[q] = src[std::min(pad - q, width - 1)];
  std::copy(src, src + width, padded.begin() + pad);
  for (int q = 0; q < pad; q++) padded[(size_t)pad + width + q] = src[std::max(width - 2 - q, 0)];
  float win[7];
  for (int w = 0; w < width; w++) {
    std::copy(padded.begin() + w, padded.begin() + w + fw, win);
    std::nth_element(win, win + fw / 2, win + fw);
    dst[w] = win[fw / 2];
  }
}

}  // namespace

std::vector<WordTiming> align_words(const float* xattn, int heads_total, int steps, int frames,
                                    const std::vector<int32_t>& tokens, float time_per_frame,
                                    const Tokenizer& tokenizer) {
  std::vector<WordTiming> out;
  if (xattn == nullptr || steps <= 0 || frames <= 0 || heads_total <= 0) return out;
  // per (head, step) row: z-score over frames, width-7 median filter; then the mean over heads
  std::vector<float> mean_map((size_t)steps * frames, 0.f), row(frames), filt(frames), padded;
  for (int h = 0; h < heads_total; h++) {
    for (int t = 0; t < steps; t++) {
      const float* src = xattn + ((size_t)h * steps + t) * frames;
      float sum = 0.f;
      for (int f = 0; f < frames; f++) sum += src[f];
      const float mean = sum / frames;
      float sq = 0.f;
      for (int f = 0; f < frames; f++) {
        const float dlt = src[f] - mean;
        sq += dlt * dlt;
      }
      float sd = std::sqrt(sq / frames);
      if (sd == 0.0f) sd = 1e-10f;
      for (int f = 0; f < frames; f++) row[f] = (src[f] - mean) / sd;
      median7_row(row.data(), frames, filt.data(), padded);
      float* acc = mean_map.data() + (size_t)t * frames;
      for (int f = 0; f < frames; f++) acc[f] += filt[f];
    }
  }
  const float inv_heads = 1.0f / heads_total;
  for (float& v : mean_map) v = -(v * inv_heads);  // DTW minimises: negate the attention
  std::vector<int> path_rows, path_cols;
  dtw_path(mean_map, steps, frames, path_rows, path_cols);

  // text tokens: everything between the start token and the last id
  if (tokens.size() < 2) return out;
  const std::vector<int32_t> text(tokens.begin() + 1, tokens.end() - 1);
  if (text.empty()) return out;
  struct Group { std::vector<int32_t> ids; int first_step, last_step; };
  std::vector<Group> groups;
  for (int i = 0; i < (int)text.size(); i++) {
    const bool starts_word = tokenizer.starts_word(text[i]);
    if (groups.empty() || (starts_word && !groups.back().ids.empty())) groups.push_back(Group{{}, i, i});
    groups.back().ids.push_back(text[i]);
    groups.back().last_step = i;
  }
  auto is_ws = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; };
  for (const Group& g : groups) {
    std::string w = tokenizer.tokens_to_text(g.ids, true);
    size_t b = 0, e = w.size();
    while (b < e && is_ws(w[b])) b++;
    while (e > b && is_ws(w[e - 1])) e--;
    w = w.substr(b, e - b);
    if (w.empty()) continue;
    int lo = frames, hi = -1;
    for (size_t k = 0; k < path_rows.size(); k++) {
      if (path_rows[k] >= g.first_step && path_rows[k] <= g.last_step) {
        lo = std::min(lo, path_cols[k]);
        hi = std::max(hi, path_cols[k]);
      }
    }
    WordTiming wt;
    wt.text = w;
    if (hi >= 0) {
      wt.start = lo * time_per_frame;
      wt.end = (hi + 1) * time_per_frame;
    }
    out.push_back(std::move(wt));
  }
  for (size_t i = 1; i < out.size(); i++) {  // overlapping neighbours meet in the middle
    if (out[i - 1].end > out[i].start) {
      const float mid = (out[i - 1].end + out[i].start) * 0.5f;
      out[i - 1].end = mid;
      out[i].start = mid;
    }
  }
  return out;
}

}  // namespace msb
