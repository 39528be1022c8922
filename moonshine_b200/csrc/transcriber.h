// Host-side transcriber: segmentation, line bookkeeping, streams and the
// transcript_t storage the C ABI hands out.  Mirrors the behaviour contract of
// the reference's Transcriber (core/transcriber.{h,cpp}) for the TINY/BASE
// architectures; the model behind it is the CUDA `Model`.
#pragma once
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/moonshine_b200.h"
#include "model.h"
#include "tokenizer.h"
#include "word_alignment.h"

namespace msb {

// Option names and defaults: core/transcriber.h:181-229,
// parsing: core/moonshine-c-api.cpp:129-198.
struct TranscriberOptions {
  bool skip_transcription = false;
  float transcription_interval = 0.5f;
  float vad_threshold = 0.5f;
  float vad_window_duration = 0.5f;
  int32_t vad_hop_size = 512;
  size_t vad_look_behind_sample_count = 8192;
  float vad_max_segment_duration = 15.0f;
  float max_tokens_per_second = 6.5f;
  bool decode_incomplete_lines = true;
  bool return_audio_data = true;
  bool log_output_text = false;
  bool word_timestamps = false;
  bool identify_speakers = false;
  bool use_speculative_decoding = true;
  int32_t context_max_terms = 0;         // 0 = ContextExtractor::kDefaultMaxTerms (200)
  float keyterm_boost = 2.0f;            // ContextBiaser::kDefaultBoost (core/context-biaser.h:44)  // streaming archs: decides the token budget rule, see update_outputs
  std::vector<std::string> keyterms;
  std::string context;
  int device = -1;  // additive option "device": CUDA ordinal (-1 = current / LOCAL_RANK)
  std::vector<int> devices;  // additive option "devices" ("0,1,2,3" or "all"): one handle fans a batch out over these GPUs
};

// Segment audio is immutable once published: the segmenter replaces the buffer instead of
// growing it in place, so transcript lines can share it without copying.
using AudioRef = std::shared_ptr<const std::vector<float>>;

struct Segment {
  AudioRef audio;
  const float* data() const { return audio ? audio->data() : nullptr; }
  size_t size() const { return audio ? audio->size() : 0; }
  float start_time = 0.f;
  float end_time = 0.f;
  bool is_complete = false;
  bool just_updated = false;
  // streaming architectures: per-segment state of Transcriber::transcribe_segment_with_streaming_model
  // (core/transcriber.cpp:1321-1372) -- samples analysed so far (whole 1280-sample chunks), encoder
  // features released to the decoder, and whether this segment was decoded before.
  size_t stream_processed = 0;
  int stream_emitted = 0;
  bool stream_decoded = false;
  uint64_t stream_keyterm_epoch = 0;  // key-term list the last decode ran under
  std::vector<int32_t> stream_tokens; // content ids of the last decode (no BOS / EOS): the next update's speculative draft
                                      // (last_streaming_tokens, core/transcriber.cpp:1403-1412, 1475)
};

// The reference's VoiceActivityDetector (core/voice-activity-detector.cpp)
// with the Silero model replaced by a constant speech probability of 1.0:
// identical hop / look-behind / smoothing-window / max-segment-fade logic,
// so `vad_threshold=0` (the documented bypass, :139,152-157) behaves exactly
// like the reference.  The Silero network itself is out of scope.
class Segmenter {
 public:
  Segmenter(float threshold, int window_size, int hop_size, size_t look_behind, size_t max_segment);
  void start();
  void stop();
  bool is_active() const { return active_; }
  // `last_call`: stop() follows immediately, so the still-open segment need not be materialised
  void process_audio(const float* audio, size_t n, int32_t sample_rate, bool last_call = false);
  std::vector<Segment>& segments() { return segments_; }

 private:
  void process_hop(const float* hop);
  void sync_open_segment_audio();
  void on_voice_start();
  void on_voice_continuing();
  void on_voice_end();
  float threshold_;
  int window_size_, hop_size_;
  size_t look_behind_count_, max_segment_;
  bool active_ = false, previous_is_voice_ = false;
  size_t samples_processed_ = 0;
  std::vector<float> probability_window_;
  size_t probability_index_ = 0;
  size_t look_behind_pos_ = 0;  // ring write position of look_behind_
  std::vector<float> look_behind_, current_, remainder_;
  std::vector<Segment> segments_;
};

struct Line {
  bool has_text = false;
  std::string text;
  AudioRef audio;
  std::vector<WordTiming> words;  // absolute times (segment start added), filled when word_timestamps is on
  float start_time = 0.f, duration = 0.f;
  uint64_t id = 0;
  int8_t is_complete = 0, just_updated = 0, is_new = 0, has_text_changed = 0;
  uint32_t latency_ms = 0;
};

class TranscriptOutput {
 public:
  void clear();
  void clear_update_flags();
  void add_or_update(Line& line);
  void mark_all_complete();
  void rebuild();
  transcript_t transcript{nullptr, 0};
  std::map<uint64_t, Line> lines;
  std::vector<uint64_t> order;

 private:
  std::vector<transcript_line_t> c_lines_;
  std::vector<std::vector<transcript_word_t>> c_words_;
};

struct Stream {
  std::unique_ptr<Segmenter> vad;
  TranscriptOutput output;
  std::vector<float> new_audio;  // 16 kHz, not yet analysed
  std::mutex mutex;
};

std::vector<float> resample_audio(const float* audio, size_t n, float in_rate, float out_rate);

class Transcriber {
 public:
  Transcriber(const TranscriberOptions& options, uint32_t model_arch);
  ~Transcriber();
  void load_from_directory(const std::string& path);
  void load_from_memory(const uint8_t* weights, size_t weights_size, const uint8_t* tokenizer,
                        size_t tokenizer_size);

  void transcribe_without_streaming(const float* audio, uint64_t n, int32_t sample_rate,
                                    uint32_t flags, transcript_t** out);
  void transcribe_batch(const float* const* audio, const uint64_t* lengths, uint64_t count,
                        int32_t sample_rate, uint32_t flags, transcript_t** out);
  int32_t create_stream();
  void free_stream(int32_t id);
  void start_stream(int32_t id);
  void stop_stream(int32_t id);
  void add_audio_to_stream(int32_t id, const float* audio, uint64_t n, int32_t sample_rate);
  void transcribe_stream(int32_t id, uint32_t flags, transcript_t** out);

  // Key-term biasing (streaming architectures only; reference: Transcriber::set_keyterms,
  // core/transcriber.cpp:249-296).  Throws on TINY/BASE like the reference.
  void set_keyterms(const std::vector<std::string>& keyterms);
  // reference: Transcriber::set_context (core/transcriber.cpp:245-247): key terms picked out of a passage
  void set_context(const std::string& context, int32_t max_terms);

  Model* model() { return model_.get(); }
  std::mutex& model_mutex() { return model_mutex_; }
  const TranscriberOptions& options() const { return options_; }
  uint32_t arch() const { return arch_; }

 private:
  std::unique_ptr<Segmenter> make_segmenter() const;
  Stream* find_stream(int32_t id);
  // Transcribes every just_updated segment of every (stream, segments) pair in
  // ONE batched model call and writes the lines.
  struct Job {
    TranscriptOutput* output;
    std::vector<Segment>* segments;
    bool stopped;
  };
  void update_outputs(std::vector<Job>& jobs);

  TranscriberOptions options_;
  uint32_t arch_;
  std::unique_ptr<Model> model_;
  std::vector<std::unique_ptr<Model>> replicas_;  // devices[1..]: weights copied device-to-device from model_
  void make_replicas();
  std::unique_ptr<Tokenizer> tokenizer_;
  std::mutex model_mutex_;      // serialises model use (reference: stt_model_mutex)
  std::mutex biaser_mutex_;     // reference: context_biaser_mutex, taken before the model mutex
  KeytermBiaser biaser_;
  uint64_t keyterm_epoch_ = 0;  // bumped by set_keyterms: earlier decodes stop counting as speculative drafts
  std::mutex batch_mutex_;      // reference: batch_stream_mutex
  std::mutex streams_mutex_;
  std::map<int32_t, std::unique_ptr<Stream>> streams_;
  int32_t next_stream_id_ = 1;  // reference: stream ids start at 1 (transcriber.cpp:104)
  std::atomic<uint64_t> next_line_id_;
  // storage behind the non-streaming results
  std::vector<std::unique_ptr<TranscriptOutput>> batch_outputs_;
  std::vector<transcript_t> batch_transcripts_;
};

}  // namespace msb
