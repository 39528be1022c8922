// Word timestamps from decoder cross-attention (reference: align_words, core/word-alignment.cpp:181-394;
// called from MoonshineModel::compute_word_timestamps, core/moonshine-model.cpp:600-645, and
// Transcriber::update_transcript_from_segments, core/transcriber.cpp:1029-1070).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "tokenizer.h"

namespace msb {

struct WordTiming {
  std::string text;
  float start = 0.f, end = 0.f;  // seconds from the start of the segment
  float confidence = 1.0f;
};

// xattn: [heads_total][steps][frames] softmax probabilities (heads_total = layers * heads, layer-major).
// tokens: generated ids including the start token (and EOS when produced); row i of the attention is the
// decoder run that produced tokens[i + 1].
std::vector<WordTiming> align_words(const float* xattn, int heads_total, int steps, int frames,
                                    const std::vector<int32_t>& tokens, float time_per_frame,
                                    const Tokenizer& tokenizer);

}  // namespace msb
