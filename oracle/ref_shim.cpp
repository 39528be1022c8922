// extern "C" shim over the REFERENCE's own host-side helpers on the transcription path, compiled from
// the sources where they lie under /root/reference (see oracle/build_ref.py): bin-tokenizer
// (core/bin-tokenizer/bin-tokenizer.cpp), resampler (core/resampler.cpp), word alignment
// (core/word-alignment.cpp) and the key-term biaser (core/context-biaser.cpp).
// TEST INFRASTRUCTURE ONLY: the resulting oracle/_ref/libmoonshine_ref_helpers.so is loaded by tests/
// to pin this repo's C++ / numpy restatements against the real reference code.  No reference source is copied.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bin-tokenizer.h"
#include "context-biaser.h"
#include "context-extractor.h"
#include "resampler.h"
#include "voice-activity-detector.h"
#include "word-alignment.h"

// The reference's VoiceActivityDetector (core/voice-activity-detector.cpp) is compiled as is; its Silero
// network (ONNX Runtime, not buildable here) is replaced by this stand-in that reports a constant speech
// probability of 1.0 -- the same substitution moonshine-b200's Segmenter makes.
// The ORT members of the class are left untouched (never dereferenced).
SileroVad::SileroVad(int, int, float, int, int, int, float) {}
SileroVad::~SileroVad() {}
void SileroVad::predict(const std::vector<float>&, float* out_probability, int* out_flag) {
  if (out_probability) *out_probability = 1.0f;
  if (out_flag) *out_flag = 1;
}

extern "C" {

void* ref_vad_new(float threshold, int32_t window_size, int32_t hop_size, uint64_t look_behind, uint64_t max_segment) {
  return new VoiceActivityDetector(threshold, window_size, hop_size, (size_t)look_behind, (size_t)max_segment);
}
void ref_vad_free(void* v) { delete static_cast<VoiceActivityDetector*>(v); }
void ref_vad_start(void* v) { static_cast<VoiceActivityDetector*>(v)->start(); }
void ref_vad_stop(void* v) { static_cast<VoiceActivityDetector*>(v)->stop(); }
void ref_vad_process(void* v, const float* audio, uint64_t n, int32_t rate) {
  static_cast<VoiceActivityDetector*>(v)->process_audio(audio, (size_t)n, rate);
}
int32_t ref_vad_segment_count(void* v) { return (int32_t)static_cast<VoiceActivityDetector*>(v)->get_segments()->size(); }
// info = {start_time, end_time, is_complete, just_updated}; returns the audio sample count
int64_t ref_vad_segment(void* v, int32_t i, float* info, float* audio_out, int64_t cap) {
  const VoiceActivitySegment& s = static_cast<VoiceActivityDetector*>(v)->get_segments()->at((size_t)i);
  info[0] = s.start_time;
  info[1] = s.end_time;
  info[2] = s.is_complete ? 1.f : 0.f;
  info[3] = s.just_updated ? 1.f : 0.f;
  const int64_t m = (int64_t)s.audio_data.size() < cap ? (int64_t)s.audio_data.size() : cap;
  if (audio_out && m > 0) std::memcpy(audio_out, s.audio_data.data(), (size_t)m * sizeof(float));
  return (int64_t)s.audio_data.size();
}

void* ref_tokenizer_new(const uint8_t* data, uint64_t size) {
  try {
    return new BinTokenizer(data, (size_t)size);
  } catch (...) {
    return nullptr;
  }
}
void* ref_tokenizer_new_bpe(const uint8_t* data, uint64_t size) {
  try {
    return new BinTokenizer(data, (size_t)size, "\xE2\x96\x81", BinTokenizerEncoding::kBpe);
  } catch (...) {
    return nullptr;
  }
}
void ref_tokenizer_free(void* t) { delete static_cast<BinTokenizer*>(t); }

// returns the byte length (or -1 on error); writes at most cap bytes
int64_t ref_tokens_to_text(void* t, const int32_t* ids, int32_t n, char* out, int64_t cap) {
  try {
    std::vector<int> v(ids, ids + n);
    const std::string s = static_cast<BinTokenizer*>(t)->tokens_to_text<int>(v, true);
    const int64_t m = (int64_t)s.size() < cap ? (int64_t)s.size() : cap;
    std::memcpy(out, s.data(), (size_t)m);
    return (int64_t)s.size();
  } catch (...) {
    return -1;
  }
}

int32_t ref_text_to_tokens(void* t, const char* text, int32_t* out, int32_t cap) {
  try {
    const std::vector<int> v = static_cast<BinTokenizer*>(t)->text_to_tokens<int>(std::string(text));
    for (int32_t i = 0; i < (int32_t)v.size() && i < cap; i++) out[i] = v[i];
    return (int32_t)v.size();
  } catch (...) {
    return -1;
  }
}

int64_t ref_resample(const float* in, int64_t n, float in_rate, float out_rate, float* out, int64_t cap) {
  const std::vector<float> a(in, in + n);
  const std::vector<float> r = resample_audio(a, in_rate, out_rate);
  const int64_t m = (int64_t)r.size() < cap ? (int64_t)r.size() : cap;
  std::memcpy(out, r.data(), (size_t)m * sizeof(float));
  return (int64_t)r.size();
}

// words: text bytes are written NUL-separated into text_out; starts/ends per word.  Returns the word count.
int32_t ref_align_words(void* tok, const float* xattn, int32_t layers, int32_t heads, int32_t steps,
                        int32_t frames, const int32_t* tokens, int32_t n_tokens, float time_per_frame,
                        float* starts, float* ends, char* text_out, int64_t text_cap, int32_t max_words) {
  const std::vector<int> tk(tokens, tokens + n_tokens);
  const std::vector<TranscriberWord> w =
      align_words(xattn, layers, heads, steps, frames, tk, time_per_frame, static_cast<BinTokenizer*>(tok));
  int64_t o = 0;
  for (int32_t i = 0; i < (int32_t)w.size() && i < max_words; i++) {
    starts[i] = w[i].start;
    ends[i] = w[i].end;
    if (o + (int64_t)w[i].text.size() + 1 <= text_cap) {
      std::memcpy(text_out + o, w[i].text.data(), w[i].text.size());
      o += (int64_t)w[i].text.size();
      text_out[o++] = 0;
    }
  }
  return (int32_t)w.size();
}

// key terms of a passage, judged by the given (BPE) tokenizer exactly as Transcriber::keyterms_from_context does
int32_t ref_extract_terms(void* tok, const char* context, int32_t max_terms, char* out, int64_t cap) {
  BinTokenizer* t = static_cast<BinTokenizer*>(tok);
  const std::vector<std::string> terms =
      ContextExtractor::extract(std::string(context), max_terms, [t](const std::string& word) -> size_t {
        try {
          return t->text_to_tokens<int32_t>(word).size();
        } catch (const std::exception&) {
          return 0;
        }
      });
  int64_t o = 0;
  for (const std::string& s : terms) {
    if (o + (int64_t)s.size() + 1 > cap) break;
    std::memcpy(out + o, s.data(), s.size());
    o += (int64_t)s.size();
    out[o++] = 0;
  }
  return (int32_t)terms.size();
}

void* ref_biaser_new() { return new ContextBiaser(); }
void ref_biaser_free(void* b) { delete static_cast<ContextBiaser*>(b); }
void ref_biaser_add(void* b, const int32_t* toks, int32_t n) {
  static_cast<ContextBiaser*>(b)->add_token_sequence(std::vector<int32_t>(toks, toks + n));
}
void ref_biaser_reset(void* b) { static_cast<ContextBiaser*>(b)->reset(); }
void ref_biaser_advance(void* b, int32_t token) { static_cast<ContextBiaser*>(b)->advance(token); }
void ref_biaser_apply(void* b, float* logits, int32_t vocab) { static_cast<ContextBiaser*>(b)->apply(logits, vocab); }
// variants of one key term, NUL-separated; returns the count
int32_t ref_biaser_variants(const char* term, char* out, int64_t cap) {
  const std::vector<std::string> v = ContextBiaser::variants_for_term(std::string(term));
  int64_t o = 0;
  for (const std::string& s : v) {
    if (o + (int64_t)s.size() + 1 > cap) break;
    std::memcpy(out + o, s.data(), s.size());
    o += (int64_t)s.size();
    out[o++] = 0;
  }
  return (int32_t)v.size();
}

}  // extern "C"
