#include "transcriber.h"

#include <sys/stat.h>

#include <chrono>
#include <cmath>
#include <numeric>
#include <random>
#include <thread>

namespace msb {

namespace {
constexpr int kSampleRate = 16000;
float seconds_from_samples(size_t n) { return static_cast<float>(n) / kSampleRate; }
bool file_exists(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
bool dir_exists(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
}  // namespace

// ---------------------------------------------------------------------------
// Resampler: behaviour of core/resampler.cpp:5-86 (box average down, linear
// interpolation up, float arithmetic).
// ---------------------------------------------------------------------------
std::vector<float> resample_audio(const float* audio, size_t n, float in_rate, float out_rate) {
  if (in_rate == out_rate) return std::vector<float>(audio, audio + n);
  const size_t out_n = (size_t)(n * out_rate / in_rate);
  std::vector<float> out(out_n);
  const float ratio = in_rate / out_rate;
  if (in_rate > out_rate) {
    for (size_t i = 0; i < out_n; i++) {
      size_t a = (size_t)(i * ratio);
      size_t b = (size_t)((i + 1) * ratio);
      if (b >= n) b = n - 1;
      float sum = 0.f;
      size_t cnt = 0;
      for (size_t j = a; j <= b; j++) { sum += audio[j]; cnt++; }
      out[i] = cnt > 0 ? sum / cnt : 0.f;
    }
  } else {
    for (size_t i = 0; i < out_n; i++) {
      const float pos = i * ratio;
      const size_t idx = (size_t)pos;
      const float frac = pos - idx;
      if (idx >= n - 1) out[i] = audio[n - 1];
      else out[i] = audio[idx] + frac * (audio[idx + 1] - audio[idx]);
    }
  }
  return out;
}

// ---------------------------------------------------------------------------
// Segmenter
// ---------------------------------------------------------------------------
Segmenter::Segmenter(float threshold, int window_size, int hop_size, size_t look_behind,
                     size_t max_segment)
    : threshold_(threshold), window_size_(window_size < 1 ? 1 : window_size),
      hop_size_(hop_size < 1 ? 512 : hop_size), look_behind_count_(look_behind),
      max_segment_(max_segment) {
  probability_window_.assign(window_size_, 0.f);
  look_behind_.assign(look_behind_count_, 0.f);
}

void Segmenter::start() {
  active_ = true;
  samples_processed_ = 0;
  segments_.clear();
  current_.clear();
  look_behind_.assign(look_behind_count_, 0.f);
  look_behind_pos_ = 0;
  remainder_.clear();
  // The smoothing window is NOT cleared: the reference's start() calls
  // probability_window.resize(window_size, 0.0f) on an already-sized vector, a no-op, so the averages of
  // the previous session carry over a stop()/start() pair (core/voice-activity-detector.cpp start()).
  probability_index_ = 0;
  previous_is_voice_ = false;
}

void Segmenter::stop() {
  active_ = false;
  if (previous_is_voice_) on_voice_end();
}

void Segmenter::process_audio(const float* audio, size_t n, int32_t sample_rate, bool last_call) {
  if (!active_) return;
  for (Segment& s : segments_) s.just_updated = false;
  std::vector<float> resampled;
  if (sample_rate != kSampleRate) {
    resampled = resample_audio(audio, n, (float)sample_rate, (float)kSampleRate);
    audio = resampled.data();
    n = resampled.size();
  }
  // the open segment grows by at most this call's audio (+ the look-behind when it opens): one allocation
  // instead of a doubling chain of reallocate-and-copy steps through the hop loop
  {
    const size_t cap = max_segment_ ? std::min(max_segment_ + (size_t)hop_size_ + look_behind_.size(),
                                               current_.size() + n + look_behind_.size())
                                    : current_.size() + n + look_behind_.size();
    if (current_.capacity() < cap) current_.reserve(cap);
  }
  // hops are analysed where they lie; only a hop that straddles two calls is assembled
  size_t pos = 0;
  if (!remainder_.empty()) {
    const size_t need = (size_t)hop_size_ - remainder_.size();
    if (n < need) {
      remainder_.insert(remainder_.end(), audio, audio + n);
      pos = n;
    } else {
      remainder_.insert(remainder_.end(), audio, audio + need);
      pos = need;
      process_hop(remainder_.data());
      remainder_.clear();
    }
  }
  while (n - pos >= (size_t)hop_size_) {
    process_hop(audio + pos);
    pos += hop_size_;
  }
  if (pos < n) remainder_.insert(remainder_.end(), audio + pos, audio + n);
  if (!last_call) sync_open_segment_audio();
}

// The reference copies the growing segment buffer into the segment on every
// hop (O(n^2) per clip); the observable state only matters when a call
// returns, so the still-open segment is materialised once here.
void Segmenter::sync_open_segment_audio() {
  if (previous_is_voice_ && !segments_.empty() && !segments_.back().is_complete) {
    segments_.back().audio = std::make_shared<const std::vector<float>>(current_);
  }
}

void Segmenter::process_hop(const float* hop) {
  samples_processed_ += hop_size_;
  // look-behind as a ring (the reference shifts the whole 8192-sample buffer on every hop;
  // the content a voice start sees is identical).  While a segment is open the ring is not needed --
  // a start can only follow an end -- so it is left alone and rebuilt from the segment's tail at the end.
  const bool ring_live = !previous_is_voice_ || look_behind_.size() < (size_t)hop_size_;
  if (!look_behind_.empty() && ring_live) {
    const size_t n = look_behind_.size();
    const size_t take = std::min(n, (size_t)hop_size_);
    const float* src = hop + hop_size_ - take;
    const size_t first = std::min(take, n - look_behind_pos_);
    std::copy(src, src + first, look_behind_.begin() + look_behind_pos_);
    std::copy(src + first, src + take, look_behind_.begin());
    look_behind_pos_ = (look_behind_pos_ + take) % n;
  }
  float smoothed;
  if (threshold_ > 0.0f) {
    probability_window_[probability_index_] = 1.0f;  // constant speech probability
    probability_index_ = (probability_index_ + 1) % probability_window_.size();
    smoothed = std::accumulate(probability_window_.begin(), probability_window_.end(), 0.0f) /
               probability_window_.size();
  } else {
    smoothed = 1.0f;
  }
  const size_t fade = (max_segment_ * 2) / 3;
  if (max_segment_ && current_.size() > fade) {
    const float fade_factor = static_cast<float>(current_.size() - fade) / fade;
    smoothed *= fade_factor;
  }
  const bool is_voice = smoothed > threshold_;
  if (is_voice && !previous_is_voice_) {
    const size_t n = look_behind_.size();
    const size_t lb = std::min(n, samples_processed_);
    current_.resize(lb);
    const size_t start = (look_behind_pos_ + n - lb) % n;
    const size_t first = std::min(lb, n - start);
    std::copy(look_behind_.begin() + start, look_behind_.begin() + start + first, current_.begin());
    std::copy(look_behind_.begin(), look_behind_.begin() + (lb - first), current_.begin() + first);
    if (n < (size_t)hop_size_ && lb < (size_t)hop_size_) {
      // look-behind shorter than a hop: the hop itself still starts the segment
      current_.assign(hop, hop + hop_size_);
    }
    on_voice_start();
  } else if (!is_voice && previous_is_voice_) {
    current_.insert(current_.end(), hop, hop + hop_size_);
    if (!ring_live) {
      // the last look_behind samples analysed = the tail of the segment that ends here (a segment begins with
      // the look-behind of its start, so a short one still reaches back to what preceded it; zeros before that)
      const size_t n = look_behind_.size();
      const size_t have = std::min(n, current_.size());
      std::fill(look_behind_.begin(), look_behind_.begin() + (n - have), 0.f);
      std::copy(current_.end() - have, current_.end(), look_behind_.begin() + (n - have));
      look_behind_pos_ = 0;
    }
    on_voice_end();
    current_.clear();
    // (the reference's `look_behind.resize(count, 0.0f)` here is a no-op on an already
    // full-size vector, so the look-behind keeps its audio across the cut)
  } else if (is_voice && previous_is_voice_) {
    current_.insert(current_.end(), hop, hop + hop_size_);
    on_voice_continuing();
  }
  previous_is_voice_ = is_voice;
}

void Segmenter::on_voice_start() {
  segments_.emplace_back();
  Segment& s = segments_.back();
  const float now = seconds_from_samples(samples_processed_);
  s.start_time = now - seconds_from_samples(current_.size());
  s.end_time = now;
  s.is_complete = false;
  s.just_updated = true;
}
void Segmenter::on_voice_continuing() {
  Segment& s = segments_.back();
  s.end_time = seconds_from_samples(samples_processed_);
  s.is_complete = false;
  s.just_updated = true;
}
void Segmenter::on_voice_end() {
  Segment& s = segments_.back();
  s.audio = std::make_shared<const std::vector<float>>(std::move(current_));  // every caller clears current_ next
  current_.clear();
  s.end_time = seconds_from_samples(samples_processed_);
  s.is_complete = true;
  s.just_updated = true;
}

// ---------------------------------------------------------------------------
// TranscriptOutput (reference: TranscriptStreamOutput, transcriber.cpp:1627-1760)
// ---------------------------------------------------------------------------
void TranscriptOutput::clear() {
  lines.clear();
  order.clear();
  c_lines_.clear();
  transcript.lines = nullptr;
  transcript.line_count = 0;
}
void TranscriptOutput::clear_update_flags() {
  for (uint64_t id : order) {
    Line& l = lines[id];
    l.just_updated = 0;
    l.is_new = 0;
    l.has_text_changed = 0;
  }
  rebuild();
}
void TranscriptOutput::add_or_update(Line& line) {
  auto it = lines.find(line.id);
  if (it != lines.end()) {
    line.is_new = 0;
    const Line& old = it->second;
    line.has_text_changed = (old.has_text != line.has_text) ||
                            (old.has_text && line.has_text && old.text != line.text);
  } else {
    line.is_new = 1;
    line.has_text_changed = line.has_text ? 1 : 0;
  }
  lines[line.id] = std::move(line);
}
void TranscriptOutput::mark_all_complete() {
  for (uint64_t id : order) {
    Line& l = lines[id];
    if (!l.is_complete) {
      l.is_complete = 1;
      l.just_updated = 1;
    }
  }
  rebuild();
}
void TranscriptOutput::rebuild() {
  c_lines_.clear();
  c_lines_.reserve(order.size());
  c_words_.assign(order.size(), {});
  size_t line_index = 0;
  for (uint64_t id : order) {
    const Line& l = lines[id];
    std::vector<transcript_word_t>& cw = c_words_[line_index++];
    for (const WordTiming& w : l.words) cw.push_back(transcript_word_t{w.text.c_str(), w.start, w.end, w.confidence});
    transcript_line_t c{};
    c.text = l.has_text ? l.text.c_str() : nullptr;
    c.audio_data = (l.audio && !l.audio->empty()) ? l.audio->data() : nullptr;
    c.audio_data_count = l.audio ? l.audio->size() : 0;
    c.start_time = l.start_time;
    c.duration = l.duration;
    c.id = l.id;
    c.is_complete = l.is_complete;
    c.is_updated = l.just_updated;
    c.is_new = l.is_new;
    c.has_text_changed = l.has_text_changed;
    c.have_speakers_changed = 0;
    c.speaker_spans = nullptr;
    c.speaker_span_count = 0;
    c.last_transcription_latency_ms = l.latency_ms;
    c.words = cw.empty() ? nullptr : cw.data();
    c.word_count = cw.size();
    c_lines_.push_back(c);
  }
  transcript.lines = c_lines_.empty() ? nullptr : c_lines_.data();
  transcript.line_count = c_lines_.size();
}

// ---------------------------------------------------------------------------
// Transcriber
// ---------------------------------------------------------------------------
Transcriber::Transcriber(const TranscriberOptions& options, uint32_t model_arch)
    : options_(options), arch_(model_arch) {
  std::random_device rd;
  next_line_id_ = ((uint64_t)rd() << 32) | (uint64_t)rd();
  if (options_.vad_threshold > 0.0f) {
    // say it at run time, not only in the docs: there is no Silero network here
    static bool warned = false;
    if (!warned) {
      warned = true;
      MSB_LOGF("vad_threshold=%g: segmentation uses a constant speech probability of 1.0 (no Silero VAD network "
               "in this build): audio is cut by the max-segment fade only and silence is transcribed. "
               "Set vad_threshold=0 for the reference's documented bypass.", (double)options_.vad_threshold);
    }
  }
}

Transcriber::~Transcriber() {}

std::unique_ptr<Segmenter> Transcriber::make_segmenter() const {
  const int window =
      (int)std::ceil((options_.vad_window_duration * kSampleRate) / options_.vad_hop_size);
  const size_t max_seg = (size_t)std::round(options_.vad_max_segment_duration * kSampleRate);
  return std::make_unique<Segmenter>(options_.vad_threshold, window, options_.vad_hop_size,
                                     options_.vad_look_behind_sample_count, max_seg);
}

static int pick_device(int requested) {
  if (requested >= 0) return requested;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    throw std::runtime_error("No CUDA device available: moonshine-b200 has no CPU fallback");
  }
  return dev;
}

// Classic architectures have fixed dimensions; the streaming family reads them from the container.
static Dims dims_for(uint32_t arch, const WeightFile& wf) {
  if (!is_streaming_arch(arch)) return dims_for_arch(arch);
  const HostTensor& c = wf.get("streaming.config");
  return dims_from_streaming_config(arch, c.data, c.count);
}

// "devices": the primary model was built from the host weights on devices[0]; every further device gets a replica
// whose weights arrive by ONE device-to-device copy of the packed blob (NVLink between peers).  A batch is then cut
// into contiguous shards, one per device, each decoded by its own host thread; there is no data-path collective.
void Transcriber::make_replicas() {
  replicas_.clear();
  for (size_t i = 1; i < options_.devices.size(); i++) replicas_.push_back(std::make_unique<Model>(*model_, options_.devices[i]));
  if (!replicas_.empty()) MSB_LOGF("one handle, %zu device contexts (weights replicated device-to-device)", replicas_.size() + 1);
}

void Transcriber::load_from_directory(const std::string& path) {
  if (options_.skip_transcription) return;
  if (!dir_exists(path)) throw std::runtime_error("Model directory '" + path + "' does not exist");
  const std::string wpath = path + "/model.msw";
  const std::string tpath = path + "/tokenizer.bin";
  if (!file_exists(wpath)) {
    throw std::runtime_error(
        "'" + wpath + "' not found. moonshine-b200 loads float weights from model.msw "
        "(see moonshine_b200/weights.py); ONNX Runtime .ort graphs cannot be executed here.");
  }
  if (!file_exists(tpath)) throw std::runtime_error("Failed to open tokenizer file at " + tpath);
  WeightFile wf;
  load_msw_file(wpath, wf);
  if (wf.arch != arch_) {
    throw std::runtime_error(format("model.msw holds architecture %u but %u was requested. This "
                                    "often indicates you're specifying the wrong model architecture",
                                    wf.arch, arch_));
  }
  tokenizer_.reset(Tokenizer::from_file(tpath));
  model_ = std::make_unique<Model>(dims_for(arch_, wf), wf, options_.devices.empty() ? pick_device(options_.device) : options_.devices[0]);
  make_replicas();
  if (!options_.keyterms.empty()) set_keyterms(options_.keyterms);
  else if (!options_.context.empty()) set_context(options_.context, options_.context_max_terms);
}

void Transcriber::load_from_memory(const uint8_t* weights, size_t weights_size,
                                   const uint8_t* tokenizer, size_t tokenizer_size) {
  if (options_.skip_transcription) return;
  WeightFile wf;
  parse_msw(weights, weights_size, wf);
  if (wf.arch != arch_) {
    throw std::runtime_error(format("model.msw holds architecture %u but %u was requested", wf.arch, arch_));
  }
  tokenizer_ = std::make_unique<Tokenizer>(tokenizer, tokenizer_size);
  model_ = std::make_unique<Model>(dims_for(arch_, wf), wf, options_.devices.empty() ? pick_device(options_.device) : options_.devices[0]);
  make_replicas();
  if (!options_.keyterms.empty()) set_keyterms(options_.keyterms);
  else if (!options_.context.empty()) set_context(options_.context, options_.context_max_terms);
}

void Transcriber::set_keyterms(const std::vector<std::string>& keyterms) {
  std::lock_guard<std::mutex> lock(biaser_mutex_);
  options_.keyterms = keyterms;
  biaser_.clear();
  biaser_.set_boost(options_.keyterm_boost);
  keyterm_epoch_++;  // drops every speculative draft: the next decode of an open segment restarts as greedy
  if (keyterms.empty()) return;
  if (!model_) return;  // skip_transcription: nothing to tokenize against
  if (!model_->dims().streaming) {
    throw std::runtime_error(
        "Key-term biasing requires one of the streaming model architectures; the loaded model does not decode "
        "through a path that can apply it.");
  }
  for (const std::string& term : keyterms) {
    for (const std::string& variant : KeytermBiaser::variants_for_term(term)) {
      const std::vector<int32_t> ids = tokenizer_->text_to_tokens(variant, /*bpe=*/true);
      if (!ids.empty()) biaser_.add_token_sequence(ids);
    }
  }
}

void Transcriber::set_context(const std::string& context, int32_t max_terms) {
  if (!model_) {  // skip_transcription: nothing to judge words against
    set_keyterms({});
    return;
  }
  if (!model_->dims().streaming) {
    throw std::runtime_error(
        "Key-term biasing requires one of the streaming model architectures; the loaded model does not decode "
        "through a path that can apply it.");
  }
  set_keyterms(extract_key_terms(context, max_terms, *tokenizer_));
}

namespace {
// one trie walk per utterance of the batch over the transcriber's shared trie
struct BatchBiasHook : LogitHook {
  const KeytermBiaser& biaser;
  std::vector<KeytermBiaser::Walk> walks;
  BatchBiasHook(const KeytermBiaser& b, size_t n) : biaser(b), walks(n) {}
  void apply(int u, float* logits, int vocab) override { biaser.apply(walks[(size_t)u], logits, vocab); }
  void advance(int u, int token) override { biaser.advance(walks[(size_t)u], token); }
  std::vector<float> shared;
  int vocab_ = 0;
  const std::vector<float>* shared_bonus(int vocab) override {
    if (shared.size() != (size_t)vocab) shared = biaser.root_bonus(vocab);
    vocab_ = vocab;
    return &shared;
  }
  void step_bonus(int u, std::vector<std::pair<int32_t, float>>& out) override {
    biaser.step_bonus(walks[(size_t)u], vocab_, out);
  }
};
}  // namespace

// One batched model call for every just_updated segment of every job.
// Reference: Transcriber::update_transcript_from_segments
// (core/transcriber.cpp:989-1148), which runs the segments serially.
void Transcriber::update_outputs(std::vector<Job>& jobs) {
  std::lock_guard<std::mutex> biaser_lock(biaser_mutex_);  // held across the decode, like the reference
  struct Pending { size_t job; size_t seg; Line line; bool run; bool strip_eos; };
  std::vector<Pending> pend;
  std::vector<const float*> ptrs;
  std::vector<uint64_t> lens;
  std::vector<int> plan_emitted, plan_max_tokens;   // streaming architectures only
  std::vector<const int*> plan_draft;                // speculative drafts (previous ids of the same open segment)
  std::vector<int> plan_draft_len;
  const bool streaming = model_ && model_->dims().streaming;
  for (size_t j = 0; j < jobs.size(); j++) {
    jobs[j].output->clear_update_flags();
    std::vector<Segment>& segs = *jobs[j].segments;
    for (size_t si = 0; si < segs.size(); si++) {
      Segment& s = segs[si];
      if (!s.just_updated) continue;
      Pending p;
      p.job = j; p.seg = si; p.run = false; p.strip_eos = false;
      p.line.start_time = s.start_time;
      p.line.duration = s.end_time - s.start_time;
      p.line.is_complete = s.is_complete;
      p.line.just_updated = s.just_updated;
      TranscriptOutput& out = *jobs[j].output;
      if (si >= out.order.size()) out.order.push_back(next_line_id_.fetch_add(1));
      p.line.id = out.order.at(si);
      if (streaming) {
        // Host bookkeeping of transcribe_segment_with_streaming_model (core/transcriber.cpp:1331-1395) +
        // MoonshineStreamingModel::encode (moonshine-streaming-model.cpp:611-640): only whole 1280-sample
        // chunks are analysed; a non-final update holds back `total_lookahead` features; when no sample is
        // new the encoder does not run at all.  The GPU then encodes the analysed audio in one pass and
        // uses the first `emitted` features as memory, which equals the reference's chunk-by-chunk state.
        const size_t L = s.size();
        if (s.stream_processed < L) {
          s.stream_processed += (L - s.stream_processed) / 1280 * 1280;
          const int n = (int)(s.stream_processed / 320);
          const int stable = s.is_complete ? n : std::max(0, n - model_->dims().lookahead());
          if (n > 0 && stable > s.stream_emitted) s.stream_emitted = stable;
        }
        p.line.has_text = true;  // the streaming path always returns a string (possibly empty)
        if (s.stream_emitted > 0 && (s.is_complete || options_.decode_incomplete_lines)) {
          p.run = true;
          ptrs.push_back(s.data());
          lens.push_back(s.stream_processed);
          plan_emitted.push_back(s.stream_emitted);
          int budget;
          if (options_.use_speculative_decoding && s.stream_decoded && s.stream_keyterm_epoch == keyterm_epoch_) {
            // decode_full (moonshine-streaming-model.cpp:1217-1219): verify-then-continue returns what greedy
            // returns, but budgets by memory length and leaves EOS out
            const float dur = (float)s.stream_emitted * 0.020f;
            budget = std::min((int)std::ceil((double)dur * 6.5), model_->dims().max_seq_len);
            p.strip_eos = true;
            // (MOONSHINE_B200_SPEC_VERIFY=0: same budgets, plain greedy launches -- the A/B switch of the tests)
            const char* off = std::getenv("MOONSHINE_B200_SPEC_VERIFY");
            const bool use_draft = !(off && off[0] == '0') && !s.stream_tokens.empty();
            plan_draft.push_back(use_draft ? s.stream_tokens.data() : nullptr);
            plan_draft_len.push_back(use_draft ? (int)s.stream_tokens.size() : 0);
          } else {
            plan_draft.push_back(nullptr);
            plan_draft_len.push_back(0);
            const float dur = (float)L / (float)kSampleRate;
            budget = std::min((int)std::ceil(dur * options_.max_tokens_per_second), 256);
          }
          plan_max_tokens.push_back(budget);
          s.stream_decoded = true;
          s.stream_keyterm_epoch = keyterm_epoch_;
        }
      } else if (model_) {
        if (!s.is_complete && !options_.decode_incomplete_lines) {
          p.line.has_text = true;  // empty string, like the reference
        } else {
          p.run = true;
          ptrs.push_back(s.data());
          lens.push_back(s.size());
        }
      }
      pend.push_back(std::move(p));
    }
  }
  std::vector<std::vector<int32_t>> tokens;
  std::vector<CrossAttention> xattn;  // filled when word_timestamps is on
  uint32_t latency_ms = 0;
  if (!ptrs.empty()) {
    std::lock_guard<std::mutex> lock(model_mutex_);
    const auto t0 = std::chrono::steady_clock::now();
    StreamPlan plan;
    plan.emitted = plan_emitted.data();
    plan.max_tokens = plan_max_tokens.data();
    plan.draft = plan_draft.data();
    plan.draft_len = plan_draft_len.data();
    const size_t n_dev = std::min(replicas_.size() + 1, ptrs.size());
    if (n_dev <= 1) {
      BatchBiasHook bias_hook(biaser_, ptrs.size());
      model_->transcribe(ptrs.data(), lens.data(), (int)ptrs.size(), options_.max_tokens_per_second, tokens,
                         nullptr, streaming ? &plan : nullptr, options_.word_timestamps ? &xattn : nullptr,
                         (streaming && !biaser_.empty()) ? &bias_hook : nullptr);
    } else {
      // contiguous shards (sizes differ by at most one), one host thread per device context
      const size_t n = ptrs.size();
      std::vector<std::vector<std::vector<int32_t>>> tok_s(n_dev);
      std::vector<std::vector<CrossAttention>> xa_s(n_dev);
      std::vector<std::exception_ptr> errs(n_dev);
      std::vector<std::thread> pool;
      std::vector<size_t> lo(n_dev + 1, 0);
      for (size_t d = 0; d < n_dev; d++) lo[d + 1] = lo[d] + n / n_dev + (d < n % n_dev ? 1 : 0);
      for (size_t d = 0; d < n_dev; d++) {
        pool.emplace_back([&, d]() {
          try {
            Model* m = d == 0 ? model_.get() : replicas_[d - 1].get();
            const size_t a = lo[d], cnt = lo[d + 1] - lo[d];
            StreamPlan sp;
            sp.emitted = streaming ? plan_emitted.data() + a : nullptr;
            sp.max_tokens = streaming ? plan_max_tokens.data() + a : nullptr;
            sp.draft = streaming ? plan_draft.data() + a : nullptr;
            sp.draft_len = streaming ? plan_draft_len.data() + a : nullptr;
            BatchBiasHook hook(biaser_, cnt);
            m->transcribe(ptrs.data() + a, lens.data() + a, (int)cnt, options_.max_tokens_per_second, tok_s[d], nullptr,
                          streaming ? &sp : nullptr, options_.word_timestamps ? &xa_s[d] : nullptr,
                          (streaming && !biaser_.empty()) ? &hook : nullptr);
          } catch (...) {
            errs[d] = std::current_exception();
          }
        });
      }
      for (auto& th : pool) th.join();
      for (auto& e : errs)
        if (e) std::rethrow_exception(e);  // a failed device fails the call
      for (size_t d = 0; d < n_dev; d++) {
        for (auto& t2 : tok_s[d]) tokens.push_back(std::move(t2));
        if (options_.word_timestamps) {
          xa_s[d].resize(lo[d + 1] - lo[d]);
          for (auto& x : xa_s[d]) xattn.push_back(std::move(x));
        }
      }
    }
    latency_ms = (uint32_t)std::chrono::duration_cast<std::chrono::milliseconds>(
                     std::chrono::steady_clock::now() - t0).count();
    if (std::getenv("MOONSHINE_B200_HOST_PROF"))
      MSB_LOGF("host profile: model.transcribe %.2f ms for %zu segments",
               std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), ptrs.size());
  }
  size_t ti = 0;
  for (Pending& p : pend) {
    Segment& s = (*jobs[p.job].segments)[p.seg];
    if (p.run) {
      const size_t utt = ti;
      std::vector<int32_t>& ids = tokens[ti++];
      if (p.strip_eos && !ids.empty() && ids.back() == model_->dims().eos) ids.pop_back();
      if (streaming) {  // the next update of this (still open) segment verifies these ids instead of re-deriving them
        s.stream_tokens.clear();
        for (int32_t id : ids)
          if (id != model_->dims().bos && id != model_->dims().eos) s.stream_tokens.push_back(id);
      }
      // Word timestamps from the cross-attention the decode just produced.  Classic architectures align
      // complete lines only (core/transcriber.cpp:1103-1117), the streaming path every decode (:1029-1070);
      // time per frame = segment duration / memory frames, times offset by the segment start.
      if (options_.word_timestamps && utt < xattn.size() && (streaming || s.is_complete) && ids.size() >= 2) {
        const CrossAttention& xa = xattn[utt];
        if (xa.steps > 0 && xa.frames > 0) {
          const float seg_duration = (float)s.size() / (float)kSampleRate;
          const float time_per_frame = seg_duration / (float)xa.frames;
          std::vector<WordTiming> words =
              align_words(xa.prob.data(), xa.heads_total, xa.steps, xa.frames, ids, time_per_frame, *tokenizer_);
          for (WordTiming& w : words) {
            w.start += s.start_time;
            w.end += s.start_time;
          }
          p.line.words = std::move(words);
        }
      }
      const std::string text = tokenizer_->tokens_to_text(ids);
      if (options_.log_output_text) MSB_LOGF("Transcribed text: '%s'", text.c_str());
      p.line.text = sanitize_utf8(text);
      p.line.has_text = true;
      p.line.latency_ms = latency_ms;
    }
    if (options_.return_audio_data) p.line.audio = s.audio;
    jobs[p.job].output->add_or_update(p.line);
  }
  for (Job& j : jobs) {
    if (j.stopped) j.output->mark_all_complete();
    j.output->rebuild();
  }
}

void Transcriber::transcribe_batch(const float* const* audio, const uint64_t* lengths,
                                   uint64_t count, int32_t sample_rate, uint32_t flags,
                                   transcript_t** out) {
  (void)flags;
  std::lock_guard<std::mutex> lock(batch_mutex_);
  static const bool host_prof = std::getenv("MOONSHINE_B200_HOST_PROF") != nullptr;
  const auto tp0 = std::chrono::steady_clock::now();
  // every call starts from fresh line state, like stream->start() does for
  // the reference's batch stream (transcriber.cpp:679-683)
  batch_outputs_.clear();
  batch_transcripts_.assign(count, transcript_t{nullptr, 0});
  std::vector<std::unique_ptr<Segmenter>> vads;
  std::vector<Job> jobs;
  for (uint64_t i = 0; i < count; i++) {
    if (audio[i] == nullptr && lengths[i] > 0) throw std::runtime_error("Audio data is nullptr");
    batch_outputs_.push_back(std::make_unique<TranscriptOutput>());
    vads.push_back(make_segmenter());
  }
  // utterances are independent: resample + segment them on the persistent worker pool
  WorkerPool::instance().parallel_for((int)count, [&](int i) {
    vads[i]->start();
    vads[i]->process_audio(audio[i], (size_t)lengths[i], sample_rate, /*last_call=*/true);
    vads[i]->stop();
  });
  const auto tp1 = std::chrono::steady_clock::now();
  for (uint64_t i = 0; i < count; i++)
    jobs.push_back(Job{batch_outputs_[i].get(), &vads[i]->segments(), true});
  update_outputs(jobs);
  for (uint64_t i = 0; i < count; i++) batch_transcripts_[i] = batch_outputs_[i]->transcript;
  if (out) *out = batch_transcripts_.data();
  if (host_prof) {
    const auto tp2 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    MSB_LOGF("host profile: segment %.2f ms, model+text %.2f ms", ms(tp0, tp1), ms(tp1, tp2));
  }
}

void Transcriber::transcribe_without_streaming(const float* audio, uint64_t n, int32_t sample_rate,
                                               uint32_t flags, transcript_t** out) {
  const float* ptrs[1] = {audio};
  uint64_t lens[1] = {n};
  transcribe_batch(ptrs, lens, 1, sample_rate, flags, out);
}

int32_t Transcriber::create_stream() {
  std::lock_guard<std::mutex> lock(streams_mutex_);
  const int32_t id = next_stream_id_++;
  auto s = std::make_unique<Stream>();
  s->vad = make_segmenter();
  streams_[id] = std::move(s);
  return id;
}

Stream* Transcriber::find_stream(int32_t id) {
  std::lock_guard<std::mutex> lock(streams_mutex_);
  auto it = streams_.find(id);
  if (it == streams_.end()) {
    throw std::runtime_error("Stream with ID " + std::to_string(id) + " not found in " +
                             std::to_string(streams_.size()) + " streams");
  }
  return it->second.get();
}

void Transcriber::free_stream(int32_t id) {
  std::lock_guard<std::mutex> lock(streams_mutex_);
  if (streams_.erase(id) == 0) throw std::runtime_error("Stream with ID " + std::to_string(id) + " not found");
}

void Transcriber::start_stream(int32_t id) {
  Stream* s = find_stream(id);
  std::lock_guard<std::mutex> lock(s->mutex);
  s->output.clear();
  s->new_audio.clear();
  s->vad->start();
}

void Transcriber::stop_stream(int32_t id) {
  Stream* s = find_stream(id);
  std::lock_guard<std::mutex> lock(s->mutex);
  s->vad->stop();
}

void Transcriber::add_audio_to_stream(int32_t id, const float* audio, uint64_t n, int32_t sample_rate) {
  Stream* s = find_stream(id);
  std::lock_guard<std::mutex> lock(s->mutex);
  if (!s->vad->is_active()) {
    throw std::runtime_error("Adding new audio for stream with ID " + std::to_string(id) +
                             " but VAD is not active. Did you call start_stream()?");
  }
  if (sample_rate == kSampleRate) {
    s->new_audio.insert(s->new_audio.end(), audio, audio + n);
  } else {
    std::vector<float> r = resample_audio(audio, (size_t)n, (float)sample_rate, (float)kSampleRate);
    s->new_audio.insert(s->new_audio.end(), r.begin(), r.end());
  }
}

// Reference: Transcriber::transcribe_stream, core/transcriber.cpp:775-913.
void Transcriber::transcribe_stream(int32_t id, uint32_t flags, transcript_t** out) {
  Stream* s = find_stream(id);
  std::lock_guard<std::mutex> lock(s->mutex);
  const bool has_new = !s->new_audio.empty();
  const float new_dur = s->new_audio.size() / (float)kSampleRate;
  const bool should_update =
      ((new_dur >= options_.transcription_interval) || (flags & MOONSHINE_FLAG_FORCE_UPDATE)) && has_new;
  const bool stopped = !s->vad->is_active();
  if (!should_update) {
    s->output.clear_update_flags();
    if (stopped) s->output.mark_all_complete();
    if (out) *out = &s->output.transcript;
    return;
  }
  s->vad->process_audio(s->new_audio.data(), s->new_audio.size(), kSampleRate);
  s->new_audio.clear();
  std::vector<Job> jobs{Job{&s->output, &s->vad->segments(), stopped}};
  update_outputs(jobs);
  if (!options_.return_audio_data) {
    for (Segment& seg : s->vad->segments())
      if (seg.is_complete) seg.audio.reset();
  }
  if (out) *out = &s->output.transcript;
}

}  // namespace msb
