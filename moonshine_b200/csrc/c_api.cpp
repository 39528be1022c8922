// extern "C" surface: the reference's transcription ABI (core/moonshine-c-api.h)
// served by the B200 runtime, plus additive batched / device entry points.
// Error model as in the reference: nothing throws across the boundary, every
// failure is logged and mapped to a negative code; loaders return the code as
// the handle (core/moonshine-c-api.cpp:249-300).
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <set>

#include <cuda_runtime.h>

#include <atomic>
#include <thread>
#include "../../include/moonshine_b200.h"
#include "transcriber.h"
#include "word_alignment.h"

using namespace msb;

namespace {

std::mutex g_map_mutex;
std::map<int32_t, Transcriber*> g_transcribers;
int32_t g_next_handle = 0;
bool g_log_api_calls = false;

std::string to_lower(const std::string& s) {
  std::string r = s;
  std::transform(r.begin(), r.end(), r.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  return r;
}
bool bool_from_string(const std::string& v) {
  const std::string s = to_lower(v);
  if (s == "true" || s == "1") return true;   // same accepted spellings as the reference
  if (s == "false" || s == "0") return false;
  throw std::runtime_error("Invalid boolean string: '" + v + "'");
}
float float_from_string(const std::string& v) {
  size_t pos = 0;
  float f = std::stof(v, &pos);
  return f;
}
int32_t int_from_string(const std::string& v) { return (int32_t)std::stol(v); }

// Same key set as parse_transcriber_options (core/moonshine-c-api.cpp:129-198);
// unknown names throw, which fails the load.
// comma-separated key terms, trimmed, empties dropped (core/moonshine-c-api.cpp:118-127)
std::vector<std::string> parse_keyterms(const std::string& value) {
  std::vector<std::string> out;
  size_t start = 0;
  while (true) {
    const size_t end = value.find(',', start);
    const std::string piece = value.substr(start, end == std::string::npos ? std::string::npos : end - start);
    const size_t b = piece.find_first_not_of(" \t");
    if (b != std::string::npos) out.push_back(piece.substr(b, piece.find_last_not_of(" \t") - b + 1));
    if (end == std::string::npos) break;
    start = end + 1;
  }
  return out;
}

void parse_options(const moonshine_option_t* options, uint64_t count, TranscriberOptions& out) {
  for (uint64_t i = 0; i < count; i++) {
    if (options[i].name == nullptr) throw std::runtime_error("Option name is null");
    const std::string name = to_lower(options[i].name);
    const std::string value = options[i].value ? options[i].value : "";
    if (name == "log_api_calls") g_log_api_calls = bool_from_string(value);
    else if (name == "skip_transcription") out.skip_transcription = true;
    else if (name == "transcription_interval") out.transcription_interval = float_from_string(value);
    else if (name == "vad_threshold") out.vad_threshold = float_from_string(value);
    else if (name == "vad_window_duration") out.vad_window_duration = float_from_string(value);
    else if (name == "vad_hop_size") out.vad_hop_size = int_from_string(value);
    else if (name == "vad_look_behind_sample_count") out.vad_look_behind_sample_count = (size_t)std::stoull(value);
    else if (name == "vad_max_segment_duration") out.vad_max_segment_duration = float_from_string(value);
    else if (name == "max_tokens_per_second") out.max_tokens_per_second = float_from_string(value);
    else if (name == "decode_incomplete_lines") out.decode_incomplete_lines = bool_from_string(value);
    else if (name == "return_audio_data") out.return_audio_data = bool_from_string(value);
    else if (name == "use_speculative_decoding") out.use_speculative_decoding = bool_from_string(value);
    else if (name == "log_output_text") out.log_output_text = bool_from_string(value);
    else if (name == "word_timestamps") out.word_timestamps = bool_from_string(value);
    else if (name == "identify_speakers") out.identify_speakers = bool_from_string(value);
    else if (name == "keyterms") out.keyterms = parse_keyterms(value);
    else if (name == "keyterm_boost") out.keyterm_boost = float_from_string(value);
    else if (name == "context") out.context = value;
    else if (name == "context_max_terms") out.context_max_terms = int_from_string(value);
    else if (name == "device") out.device = int_from_string(value);  // additive
    else if (name == "devices") {                                      // additive: "0,1,2,3" or "all"
      out.devices.clear();
      if (value == "all") {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
        for (int i = 0; i < n; i++) out.devices.push_back(i);
      } else {
        size_t start = 0;
        while (start <= value.size()) {
          const size_t end = value.find(',', start);
          const std::string tok = value.substr(start, end == std::string::npos ? std::string::npos : end - start);
          if (!tok.empty()) out.devices.push_back(int_from_string(tok));
          if (end == std::string::npos) break;
          start = end + 1;
        }
      }
      if (out.devices.empty()) throw std::runtime_error("option 'devices' names no device");
    }
    // accepted for compatibility, no effect on this runtime (ORT / CPU-side features)
    else if (name == "save_input_wav_path" || name == "log_ort_run" ||
             name == "diarization_cluster_cadence" ||
             name == "diarization_analyze_cadence" || name == "diarization_cluster_window_sec" ||
             name == "diarization_model_dir" || name == "spelling_model_path" || name == "ort_providers" ||
             name == "ort_provider" || name == "coreml_cache_dir") {}
    else throw std::runtime_error("Unknown transcriber option: '" + name + "', value=" + value);
  }
  if (out.identify_speakers) throw std::runtime_error("identify_speakers is not supported by moonshine-b200");
}

int32_t register_transcriber(Transcriber* t) {
  std::lock_guard<std::mutex> lock(g_map_mutex);
  const int32_t h = g_next_handle++;
  g_transcribers[h] = t;
  return h;
}

Transcriber* lookup(int32_t handle) {
  std::lock_guard<std::mutex> lock(g_map_mutex);
  if (handle < 0) return nullptr;
  auto it = g_transcribers.find(handle);
  return it == g_transcribers.end() ? nullptr : it->second;
}

#define CHECK_HANDLE(t, handle)                                                \
  Transcriber* t = lookup(handle);                                             \
  if (t == nullptr) {                                                          \
    MSB_LOGF("Moonshine transcriber handle is invalid: handle %d", handle);    \
    return MOONSHINE_ERROR_INVALID_HANDLE;                                     \
  }

// Canonical asset names the reference accepts as in-memory keys
// (c-api.h:497-516) plus this runtime's weight container.
const std::set<std::string>& known_memory_keys() {
  static const std::set<std::string> k = {
      "encoder_model.ort", "decoder_model_merged.ort", "tokenizer.bin", "decoder_with_attention.ort",
      "alignment_model.ort", "frontend.ort", "encoder.ort", "adapter.ort", "cross_kv.ort",
      "decoder_kv.ort", "streaming_config.json", "decoder_kv_with_attention.ort", "spelling_cnn.ort",
      "segmentation.ort", "embedding.ort", "model.msw"};
  return k;
}

std::string basename_of(const std::string& p) {
  const size_t pos = p.find_last_of('/');
  return pos == std::string::npos ? p : p.substr(pos + 1);
}

std::vector<uint8_t> read_file(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("Failed to open '" + path + "'");
  std::fseek(f, 0, SEEK_END);
  long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)std::max<long>(sz, 0));
  if (sz > 0 && std::fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) {
    std::fclose(f);
    throw std::runtime_error("Failed to read '" + path + "'");
  }
  std::fclose(f);
  return buf;
}

}  // namespace

extern "C" {

int32_t moonshine_get_version(void) { return MOONSHINE_HEADER_VERSION; }

const char* moonshine_error_to_string(int32_t error) {
  if (error == MOONSHINE_ERROR_NONE) return "Success";
  if (error == MOONSHINE_ERROR_INVALID_HANDLE) return "Invalid handle";
  if (error == MOONSHINE_ERROR_INVALID_ARGUMENT) return "Invalid argument";
  return "Unknown error";
}

void moonshine_free_buffer(void* ptr) { std::free(ptr); }

int32_t moonshine_load_transcriber_from_files(const char* path, uint32_t model_arch,
                                              const moonshine_option_t* options,
                                              uint64_t options_count, int32_t moonshine_version) {
  (void)moonshine_version;
  Transcriber* t = nullptr;
  try {
    TranscriberOptions opts;
    parse_options(options, options_count, opts);
    if (g_log_api_calls) MSB_LOGF("moonshine_load_transcriber_from_files(path=%s, model_arch=%u)", path ? path : "(null)", model_arch);
    if (!opts.skip_transcription) {
      if (!is_streaming_arch(model_arch)) dims_for_arch(model_arch);  // validates
      if (path == nullptr) throw std::runtime_error("Model path is null");
    }
    t = new Transcriber(opts, model_arch);
    if (!opts.skip_transcription) t->load_from_directory(path);
  } catch (const std::exception& e) {
    MSB_LOGF("Failed to load transcriber: %s", e.what());
    delete t;
    return MOONSHINE_ERROR_UNKNOWN;
  }
  return register_transcriber(t);
}

int32_t moonshine_load_transcriber_from_memory(const uint8_t*, size_t, const uint8_t*, size_t,
                                               const uint8_t*, size_t, const uint8_t*, size_t,
                                               uint32_t, const moonshine_option_t*, uint64_t,
                                               int32_t moonshine_version) {
  if (moonshine_version >= MOONSHINE_FROM_MEMORY_REMOVED_VERSION) {
    MSB_LOGF("moonshine_load_transcriber_from_memory is no longer supported for header version %d; "
             "use moonshine_load_transcriber_from_memory_files", moonshine_version);
    return MOONSHINE_ERROR_INVALID_ARGUMENT;
  }
  MSB_LOGF("moonshine_load_transcriber_from_memory takes ONNX Runtime graph bytes, which moonshine-b200 "
           "cannot execute; pass model.msw through moonshine_load_transcriber_from_memory_files");
  return MOONSHINE_ERROR_UNKNOWN;
}

int32_t moonshine_load_transcriber_from_memory_files(const char** filenames, const uint8_t** memory,
                                                     const uint64_t* memory_sizes, uint64_t file_count,
                                                     uint32_t model_arch, const moonshine_option_t* options,
                                                     uint64_t options_count, int32_t moonshine_version) {
  (void)moonshine_version;
  Transcriber* t = nullptr;
  try {
    TranscriberOptions opts;
    parse_options(options, options_count, opts);
    if (file_count > 0 && filenames == nullptr) return MOONSHINE_ERROR_INVALID_ARGUMENT;
    const uint8_t *wbytes = nullptr, *tbytes = nullptr;
    size_t wsize = 0, tsize = 0;
    std::vector<uint8_t> wfile, tfile;
    for (uint64_t i = 0; i < file_count; i++) {
      if (filenames[i] == nullptr) return MOONSHINE_ERROR_INVALID_ARGUMENT;
      const std::string key = basename_of(filenames[i]);
      if (!known_memory_keys().count(key)) {
        MSB_LOGF("Unrecognized in-memory model file key '%s'", filenames[i]);
        return MOONSHINE_ERROR_INVALID_ARGUMENT;
      }
      const bool in_memory = memory && memory_sizes && memory[i] != nullptr && memory_sizes[i] > 0;
      if (key == "model.msw") {
        if (in_memory) { wbytes = memory[i]; wsize = (size_t)memory_sizes[i]; }
        else { wfile = read_file(filenames[i]); wbytes = wfile.data(); wsize = wfile.size(); }
      } else if (key == "tokenizer.bin") {
        if (in_memory) { tbytes = memory[i]; tsize = (size_t)memory_sizes[i]; }
        else { tfile = read_file(filenames[i]); tbytes = tfile.data(); tsize = tfile.size(); }
      }
    }
    if (!opts.skip_transcription) {
      if (!is_streaming_arch(model_arch)) dims_for_arch(model_arch);
      if (wbytes == nullptr) throw std::runtime_error("Missing required asset 'model.msw'");
      if (tbytes == nullptr) throw std::runtime_error("Missing required asset 'tokenizer.bin'");
    }
    t = new Transcriber(opts, model_arch);
    if (!opts.skip_transcription) t->load_from_memory(wbytes, wsize, tbytes, tsize);
  } catch (const std::exception& e) {
    MSB_LOGF("Failed to load transcriber from memory files: %s", e.what());
    delete t;
    return MOONSHINE_ERROR_UNKNOWN;
  }
  return register_transcriber(t);
}

void moonshine_free_transcriber(int32_t transcriber_handle) {
  Transcriber* t = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_transcribers.find(transcriber_handle);
    if (it == g_transcribers.end()) return;
    t = it->second;
    g_transcribers.erase(it);
  }
  delete t;
}

int32_t moonshine_transcribe_without_streaming(int32_t transcriber_handle, float* audio_data,
                                               uint64_t audio_length, int32_t sample_rate,
                                               uint32_t flags, transcript_t** out_transcript) {
  CHECK_HANDLE(t, transcriber_handle);
  try {
    t->transcribe_without_streaming(audio_data, audio_length, sample_rate, flags, out_transcript);
  } catch (const std::exception& e) {
    MSB_LOGF("Failed to transcribe without streaming: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
  return MOONSHINE_ERROR_NONE;
}

int32_t moonshine_transcribe_batch_without_streaming(int32_t transcriber_handle, const float* const* audio,
                                                     const uint64_t* lengths, uint64_t count,
                                                     int32_t sample_rate, uint32_t flags,
                                                     transcript_t** out_transcripts) {
  CHECK_HANDLE(t, transcriber_handle);
  if (count > 0 && (audio == nullptr || lengths == nullptr)) return MOONSHINE_ERROR_INVALID_ARGUMENT;
  try {
    t->transcribe_batch(audio, lengths, count, sample_rate, flags, out_transcripts);
  } catch (const std::exception& e) {
    MSB_LOGF("Failed to transcribe batch: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
  return MOONSHINE_ERROR_NONE;
}

int32_t moonshine_create_stream(int32_t transcriber_handle, uint32_t flags) {
  (void)flags;
  CHECK_HANDLE(t, transcriber_handle);
  try {
    return t->create_stream();
  } catch (const std::exception& e) {
    MSB_LOGF("Failed to create stream: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
}

#define STREAM_CALL(expr, desc)                          \
  CHECK_HANDLE(t, transcriber_handle);                   \
  try {                                                  \
    expr;                                                \
  } catch (const std::exception& e) {                    \
    MSB_LOGF("Failed to " desc ": %s", e.what());        \
    return MOONSHINE_ERROR_UNKNOWN;                      \
  }                                                      \
  return MOONSHINE_ERROR_NONE;

int32_t moonshine_free_stream(int32_t transcriber_handle, int32_t stream_handle) {
  STREAM_CALL(t->free_stream(stream_handle), "free stream")
}
int32_t moonshine_start_stream(int32_t transcriber_handle, int32_t stream_handle) {
  STREAM_CALL(t->start_stream(stream_handle), "start stream")
}
int32_t moonshine_stop_stream(int32_t transcriber_handle, int32_t stream_handle) {
  STREAM_CALL(t->stop_stream(stream_handle), "stop stream")
}
int32_t moonshine_transcribe_add_audio_to_stream(int32_t transcriber_handle, int32_t stream_handle,
                                                 const float* new_audio_data, uint64_t audio_length,
                                                 int32_t sample_rate, uint32_t flags) {
  (void)flags;
  STREAM_CALL(t->add_audio_to_stream(stream_handle, new_audio_data, audio_length, sample_rate),
              "add audio to stream")
}
int32_t moonshine_transcribe_stream(int32_t transcriber_handle, int32_t stream_handle, uint32_t flags,
                                    transcript_t** out_transcript) {
  STREAM_CALL(t->transcribe_stream(stream_handle, flags, out_transcript), "transcribe stream")
}

int32_t moonshine_transcriber_set_keyterms(int32_t transcriber_handle, const char* keyterms) {
  CHECK_HANDLE(t, transcriber_handle);
  try {
    t->set_keyterms(keyterms == nullptr ? std::vector<std::string>() : parse_keyterms(keyterms));
  } catch (const std::exception& e) {
    MSB_LOGF("Failed to set key terms: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
  return MOONSHINE_ERROR_NONE;
}
int32_t moonshine_transcriber_set_context(int32_t transcriber_handle, const char* context, int32_t max_terms) {
  CHECK_HANDLE(t, transcriber_handle);
  try {
    t->set_context(context == nullptr ? std::string() : std::string(context), max_terms);
  } catch (const std::exception& e) {
    MSB_LOGF("Failed to set context: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
  return MOONSHINE_ERROR_NONE;
}

const char* moonshine_transcript_to_string(const transcript_t* transcript) {
  static std::string description;  // process-static, like the reference
  std::string r;
  if (transcript == nullptr) {
    description = "0 lines\n";
    return description.c_str();
  }
  r += std::to_string(transcript->line_count) + " lines\n";
  for (uint64_t i = 0; i < transcript->line_count; i++) {
    const transcript_line_t& line = transcript->lines[i];
    char ts[32];
    snprintf(ts, sizeof(ts), "%.1fs: ", line.start_time);
    r += ts;
    r += line.text == nullptr ? std::string("<null>") : std::string(line.text);
    r += "\n";
  }
  description = r;
  return description.c_str();
}

// ---- additive: device path, timing, parity hooks ----
int32_t moonshine_b200_transcribe_device(int32_t transcriber_handle, const float* d_pcm, int64_t stride,
                                         const uint64_t* lengths, uint64_t count, int32_t* out_tokens,
                                         int32_t out_stride, int32_t* out_counts) {
  CHECK_HANDLE(t, transcriber_handle);
  if (t->model() == nullptr || d_pcm == nullptr || lengths == nullptr) return MOONSHINE_ERROR_INVALID_ARGUMENT;
  if (stride < 0 || count > (uint64_t)INT32_MAX) return MOONSHINE_ERROR_INVALID_ARGUMENT;
  for (uint64_t i = 0; i < count; i++)  // a row longer than the stride would read the neighbour's (or no) memory
    if (lengths[i] > (uint64_t)stride) return MOONSHINE_ERROR_INVALID_ARGUMENT;
  try {
    std::lock_guard<std::mutex> lock(t->model_mutex());
    std::vector<std::vector<int32_t>> tokens;
    t->model()->transcribe_device(d_pcm, stride, lengths, (int)count, t->options().max_tokens_per_second, tokens);
    for (uint64_t i = 0; i < count; i++) {
      const int n = (int)std::min<size_t>(tokens[i].size(), (size_t)std::max(out_stride, 0));
      if (out_counts) out_counts[i] = (int32_t)tokens[i].size();
      if (out_tokens) std::memcpy(out_tokens + (size_t)i * out_stride, tokens[i].data(), n * sizeof(int32_t));
    }
  } catch (const std::exception& e) {
    MSB_LOGF("Failed to transcribe device batch: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
  return MOONSHINE_ERROR_NONE;
}

void* moonshine_b200_get_stream(int32_t transcriber_handle) {
  Transcriber* t = lookup(transcriber_handle);
  if (t == nullptr || t->model() == nullptr) return nullptr;
  return (void*)t->model()->stream();
}

int32_t moonshine_b200_set_timing(int32_t transcriber_handle, int32_t enabled) {
  CHECK_HANDLE(t, transcriber_handle);
  if (t->model() == nullptr) return MOONSHINE_ERROR_INVALID_ARGUMENT;
  t->model()->set_timing(enabled != 0);
  return MOONSHINE_ERROR_NONE;
}

int32_t moonshine_b200_debug_stream_partial(int32_t transcriber_handle, int32_t enabled) {
  CHECK_HANDLE(t, transcriber_handle);
  if (t->model() == nullptr) return MOONSHINE_ERROR_INVALID_ARGUMENT;
  t->model()->set_debug_stream_partial(enabled != 0);
  return MOONSHINE_ERROR_NONE;
}

int32_t moonshine_b200_last_timings(int32_t transcriber_handle, double* out8) {
  CHECK_HANDLE(t, transcriber_handle);
  if (t->model() == nullptr || out8 == nullptr) return MOONSHINE_ERROR_INVALID_ARGUMENT;
  const StageTimes& s = t->model()->last_times();
  out8[0] = s.frontend_ms; out8[1] = s.encoder_ms; out8[2] = s.cross_kv_ms; out8[3] = s.decode_ms;
  out8[4] = s.decode_steps; out8[5] = s.kernel_launches; out8[6] = (double)t->model()->weight_bytes();
  out8[7] = s.decoder_version;
  return MOONSHINE_ERROR_NONE;
}

int32_t moonshine_b200_debug_run(int32_t transcriber_handle, const float* const* audio,
                                 const uint64_t* lengths, uint64_t count, float* enc_out,
                                 uint64_t enc_out_capacity, int32_t* enc_frames, const int32_t* forced,
                                 int32_t forced_stride, float* logits, int32_t logits_steps,
                                 int32_t* out_tokens, int32_t out_stride, int32_t* out_counts) {
  CHECK_HANDLE(t, transcriber_handle);
  if (t->model() == nullptr) return MOONSHINE_ERROR_INVALID_ARGUMENT;
  try {
    std::lock_guard<std::mutex> lock(t->model_mutex());
    DebugCapture dbg;
    std::vector<float> enc, lg;
    std::vector<int> frames;
    if (enc_out || enc_frames) { dbg.encoder_out = &enc; dbg.encoder_frames = &frames; }
    if (logits && logits_steps > 0) { dbg.logits = &lg; dbg.logits_steps = logits_steps; }
    dbg.forced = forced;
    dbg.forced_stride = forced_stride;
    std::vector<std::vector<int32_t>> tokens;
    t->model()->transcribe(audio, lengths, (int)count, t->options().max_tokens_per_second, tokens, &dbg);
    if (enc_out) {
      if (enc.size() > enc_out_capacity) return MOONSHINE_ERROR_INVALID_ARGUMENT;
      std::memcpy(enc_out, enc.data(), enc.size() * sizeof(float));
    }
    if (enc_frames) for (uint64_t i = 0; i < count; i++) enc_frames[i] = frames[i];
    if (logits) {
      std::memset(logits, 0, (size_t)logits_steps * count * t->model()->dims().vocab * sizeof(float));
      std::memcpy(logits, lg.data(), lg.size() * sizeof(float));
    }
    for (uint64_t i = 0; i < count; i++) {
      const int n = (int)std::min<size_t>(tokens[i].size(), (size_t)std::max(out_stride, 0));
      if (out_counts) out_counts[i] = (int32_t)tokens[i].size();
      if (out_tokens) std::memcpy(out_tokens + (size_t)i * out_stride, tokens[i].data(), n * sizeof(int32_t));
    }
  } catch (const std::exception& e) {
    MSB_LOGF("debug run failed: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
  return MOONSHINE_ERROR_NONE;
}

int32_t moonshine_b200_decode_with_drafts(int32_t transcriber_handle, const float* const* audio, const uint64_t* lengths,
                                          uint64_t count, const int32_t* drafts, int32_t draft_stride,
                                          const int32_t* draft_lens, int32_t* out_tokens, int32_t out_stride,
                                          int32_t* out_counts, int32_t* out_launches) {
  CHECK_HANDLE(t, transcriber_handle);
  if (t->model() == nullptr || audio == nullptr || lengths == nullptr) return MOONSHINE_ERROR_INVALID_ARGUMENT;
  if (count > 0 && (drafts == nullptr || draft_lens == nullptr || draft_stride < 0)) return MOONSHINE_ERROR_INVALID_ARGUMENT;
  try {
    std::lock_guard<std::mutex> lock(t->model_mutex());
    std::vector<const int*> dptr(count, nullptr);
    std::vector<int> dlen(count, 0);
    for (uint64_t i = 0; i < count; i++) {
      if (draft_lens[i] < 0 || draft_lens[i] > draft_stride) return MOONSHINE_ERROR_INVALID_ARGUMENT;
      dptr[i] = drafts + (size_t)i * draft_stride;
      dlen[i] = draft_lens[i];
    }
    std::vector<std::vector<int32_t>> tokens;
    t->model()->set_debug_drafts(dptr.data(), dlen.data());
    t->model()->transcribe(audio, lengths, (int)count, t->options().max_tokens_per_second, tokens, nullptr);
    if (out_launches) *out_launches = (int32_t)t->model()->last_times().decode_steps;
    for (uint64_t i = 0; i < count; i++) {
      const int n = (int)std::min<size_t>(tokens[i].size(), (size_t)std::max(out_stride, 0));
      if (out_counts) out_counts[i] = (int32_t)tokens[i].size();
      if (out_tokens) std::memcpy(out_tokens + (size_t)i * out_stride, tokens[i].data(), n * sizeof(int32_t));
    }
  } catch (const std::exception& e) {
    MSB_LOGF("decode_with_drafts failed: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
  return MOONSHINE_ERROR_NONE;
}

int32_t moonshine_b200_decode_tokens(int32_t transcriber_handle, const float* const* audio, const uint64_t* lengths,
                                     uint64_t count, const int32_t* tokens, int32_t tokens_stride, int32_t n_steps,
                                     int32_t rows_per_launch, float* logits_out) {
  CHECK_HANDLE(t, transcriber_handle);
  if (t->model() == nullptr || audio == nullptr || lengths == nullptr || tokens == nullptr || logits_out == nullptr ||
      n_steps <= 0 || tokens_stride < n_steps || rows_per_launch < 2 || rows_per_launch > 16)
    return MOONSHINE_ERROR_INVALID_ARGUMENT;
  try {
    std::lock_guard<std::mutex> lock(t->model_mutex());
    DebugCapture dbg;
    std::vector<float> lg;
    dbg.logits = &lg;
    dbg.logits_steps = n_steps;
    dbg.forced = tokens;
    dbg.forced_stride = tokens_stride;
    dbg.rows_per_launch = rows_per_launch;
    std::vector<std::vector<int32_t>> ids;
    t->model()->transcribe(audio, lengths, (int)count, t->options().max_tokens_per_second, ids, &dbg);
    std::memset(logits_out, 0, (size_t)n_steps * count * t->model()->dims().vocab * sizeof(float));
    std::memcpy(logits_out, lg.data(), std::min(lg.size(), (size_t)n_steps * count * t->model()->dims().vocab) * sizeof(float));
  } catch (const std::exception& e) {
    MSB_LOGF("decode_tokens failed: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
  return MOONSHINE_ERROR_NONE;
}

// Self-test of the persistent host worker pool (no GPU): `callers` threads each run parallel_for over `n` items `rounds`
// times and check every item ran exactly once; one extra round throws from an item and must surface in the caller.
// Returns 0 when everything held.
int32_t moonshine_b200_debug_pool_selftest(int32_t callers, int32_t n, int32_t rounds) {
  try {
    std::atomic<int> bad{0};
    std::vector<std::thread> ts;
    for (int c = 0; c < callers; c++) {
      ts.emplace_back([&]() {
        std::vector<std::atomic<int>> hits(n);
        for (int r = 0; r < rounds; r++) {
          for (auto& h : hits) h.store(0);
          WorkerPool::instance().parallel_for(n, [&](int i) { hits[i].fetch_add(1); });
          for (int i = 0; i < n; i++)
            if (hits[i].load() != 1) bad.fetch_add(1);
        }
        bool thrown = false;
        try {
          WorkerPool::instance().parallel_for(n, [&](int i) {
            if (i == n / 2) throw std::runtime_error("item failed");
          });
        } catch (const std::runtime_error&) {
          thrown = true;
        }
        if (!thrown && n > 0) bad.fetch_add(1);
      });
    }
    for (auto& t : ts) t.join();
    return bad.load();
  } catch (const std::exception& e) {
    MSB_LOGF("pool selftest failed: %s", e.what());
    return -1;
  }
}

// ---- host-only parity hooks (no GPU needed): the product's own helpers, callable from the CPU tests that
// compare them with a build of the reference's own sources ----
int64_t moonshine_b200_debug_tokens_to_text(const uint8_t* tokenizer, uint64_t tokenizer_size, const int32_t* ids,
                                            int32_t n, char* out, int64_t cap) {
  try {
    Tokenizer tk(tokenizer, (size_t)tokenizer_size);
    const std::string s = tk.tokens_to_text(std::vector<int32_t>(ids, ids + n));
    const int64_t m = std::min<int64_t>((int64_t)s.size(), cap);
    if (out && m > 0) std::memcpy(out, s.data(), (size_t)m);
    return (int64_t)s.size();
  } catch (const std::exception& e) {
    MSB_LOGF("debug_tokens_to_text failed: %s", e.what());
    return -1;
  }
}

int32_t moonshine_b200_debug_align_words(const uint8_t* tokenizer, uint64_t tokenizer_size, const float* xattn,
                                         int32_t heads_total, int32_t steps, int32_t frames, const int32_t* tokens,
                                         int32_t n_tokens, float time_per_frame, float* starts, float* ends,
                                         char* text_out, int64_t text_cap, int32_t max_words) {
  try {
    Tokenizer tk(tokenizer, (size_t)tokenizer_size);
    const std::vector<WordTiming> w = align_words(xattn, heads_total, steps, frames,
                                                  std::vector<int32_t>(tokens, tokens + n_tokens), time_per_frame, tk);
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
  } catch (const std::exception& e) {
    MSB_LOGF("debug_align_words failed: %s", e.what());
    return -1;
  }
}

int32_t moonshine_b200_debug_text_to_tokens(const uint8_t* tokenizer, uint64_t tokenizer_size, const char* text,
                                            int32_t bpe, int32_t* out, int32_t cap) {
  try {
    Tokenizer tk(tokenizer, (size_t)tokenizer_size);
    const std::vector<int32_t> v = tk.text_to_tokens(std::string(text ? text : ""), bpe != 0);
    for (int32_t i = 0; i < (int32_t)v.size() && i < cap; i++) out[i] = v[i];
    return (int32_t)v.size();
  } catch (const std::exception& e) {
    return -1;
  }
}

// Builds a biaser from `n_seqs` token sequences, walks it along `path`, then adds the bonuses to logits.
int32_t moonshine_b200_debug_biaser_apply(const int32_t* seqs, const int32_t* seq_lens, int32_t n_seqs, float boost,
                                          const int32_t* path, int32_t n_path, float* logits, int32_t vocab) {
  try {
    KeytermBiaser b;
    b.set_boost(boost);
    size_t o = 0;
    for (int32_t i = 0; i < n_seqs; i++) {
      b.add_token_sequence(std::vector<int32_t>(seqs + o, seqs + o + seq_lens[i]));
      o += (size_t)seq_lens[i];
    }
    KeytermBiaser::Walk w;
    for (int32_t i = 0; i < n_path; i++) b.advance(w, path[i]);
    b.apply(w, logits, vocab);
    return MOONSHINE_ERROR_NONE;
  } catch (const std::exception& e) {
    MSB_LOGF("debug_biaser_apply failed: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
}

// The same through the sparse form the on-device path uploads: dense root bonuses + per-step (token, delta) pairs.
int32_t moonshine_b200_debug_biaser_apply_sparse(const int32_t* seqs, const int32_t* seq_lens, int32_t n_seqs, float boost,
                                                 const int32_t* path, int32_t n_path, float* logits, int32_t vocab) {
  try {
    KeytermBiaser b;
    b.set_boost(boost);
    size_t o = 0;
    for (int32_t i = 0; i < n_seqs; i++) {
      b.add_token_sequence(std::vector<int32_t>(seqs + o, seqs + o + seq_lens[i]));
      o += (size_t)seq_lens[i];
    }
    KeytermBiaser::Walk w;
    for (int32_t i = 0; i < n_path; i++) b.advance(w, path[i]);
    const std::vector<float> shared = b.root_bonus(vocab);
    std::vector<std::pair<int32_t, float>> extra;
    b.step_bonus(w, vocab, extra);
    std::vector<float> bonus(shared);
    for (const auto& e : extra) bonus[(size_t)e.first] += e.second;
    for (int32_t v = 0; v < vocab; v++)
      if (bonus[(size_t)v] != 0.0f) logits[v] += bonus[(size_t)v];
    return MOONSHINE_ERROR_NONE;
  } catch (const std::exception& e) {
    MSB_LOGF("debug_biaser_apply_sparse failed: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
}

// key terms of a passage, NUL-separated; returns the term count
int32_t moonshine_b200_debug_extract_terms(const uint8_t* tokenizer, uint64_t tokenizer_size, const char* context,
                                           int32_t max_terms, char* out, int64_t cap) {
  try {
    Tokenizer tk(tokenizer, (size_t)tokenizer_size);
    const std::vector<std::string> terms = extract_key_terms(std::string(context ? context : ""), max_terms, tk);
    int64_t o = 0;
    for (const std::string& t2 : terms) {
      if (o + (int64_t)t2.size() + 1 > cap) break;
      std::memcpy(out + o, t2.data(), t2.size());
      o += (int64_t)t2.size();
      out[o++] = 0;
    }
    return (int32_t)terms.size();
  } catch (const std::exception& e) {
    MSB_LOGF("debug_extract_terms failed: %s", e.what());
    return -1;
  }
}

int64_t moonshine_b200_debug_resample(const float* in, int64_t n, float in_rate, float out_rate, float* out,
                                      int64_t cap) {
  try {
    const std::vector<float> r = resample_audio(in, (size_t)n, in_rate, out_rate);
    const int64_t m = std::min<int64_t>((int64_t)r.size(), cap);
    if (out && m > 0) std::memcpy(out, r.data(), (size_t)m * sizeof(float));
    return (int64_t)r.size();
  } catch (const std::exception& e) {
    MSB_LOGF("debug_resample failed: %s", e.what());
    return -1;
  }
}

float moonshine_b200_test_ring_bandwidth(int64_t bytes_per_cta, int32_t stage_bytes, int32_t stages, int32_t nsub,
                                         int32_t shared_src, int32_t grid) {
  try {
    return ring_bandwidth_test(bytes_per_cta, stage_bytes, stages, nsub, shared_src, grid);
  } catch (const std::exception& e) {
    MSB_LOGF("ring bandwidth test failed: %s", e.what());
    return -1.f;
  }
}

int32_t moonshine_b200_test_gemm(const float* dA, const float* dW, float* dC, int32_t M, int32_t N,
                                 int32_t K, int32_t lda, int32_t ldw, int32_t ldc, const float* d_bias,
                                 int32_t act, int32_t accumulate, int32_t impl) {
  try {
    GemmParams g;
    g.A = dA; g.W = dW; g.C = dC; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldw = ldw; g.rs = ldc;
    g.bias = d_bias; g.act = act; g.accumulate = accumulate;
    if (impl >= 3 && impl <= 6) {  // 3 / 4: one tile per CTA, 5 / 6: persistent macro tiles; 4 and 6 also round-trip plane output
      // plane-fed kernel: both operands are converted to hi/lo plane tiles on the device first.  impl 4 also makes the
      // epilogue write the result as plane tiles and reads those back through a second plane-fed product with I_N.
      DeviceBuffer<float> pa, pw, pc;
      pa.reserve(plane_tiles_bytes(M, K) / 4);
      pw.reserve(plane_tiles_bytes(N, K) / 4);
      launch_rows_to_planes(dA, lda, M, K, reinterpret_cast<unsigned char*>(pa.ptr), nullptr);
      CUDA_CHECK(cudaMemsetAsync(pw.ptr, 0, pw.bytes(), nullptr));
      launch_rows_to_planes(dW, ldw, N, K, reinterpret_cast<unsigned char*>(pw.ptr), nullptr);
      GemmPlanesParams q;
      q.A = reinterpret_cast<unsigned char*>(pa.ptr); q.W = reinterpret_cast<unsigned char*>(pw.ptr);
      q.M = M; q.N = N; q.K = K; q.C = dC; q.ldc = ldc; q.bias = d_bias; q.act = act; q.accumulate = accumulate;
      q.variant = impl >= 5 ? 2 : 1;
      if (impl == 4 || impl == 6) {
        pc.reserve(plane_tiles_bytes(M, N) / 4);
        q.P = reinterpret_cast<unsigned char*>(pc.ptr);
      }
      launch_gemm_planes(q, nullptr);
      if (impl == 4 || impl == 6) {  // dC <- (planes of C) x I_N
        DeviceBuffer<float> eye, pe;
        eye.reserve((size_t)N * N);
        CUDA_CHECK(cudaMemsetAsync(eye.ptr, 0, eye.bytes(), nullptr));
        std::vector<float> h((size_t)N * N, 0.f);
        for (int i = 0; i < N; i++) h[(size_t)i * N + i] = 1.0f;
        CUDA_CHECK(cudaMemcpy(eye.ptr, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
        pe.reserve(plane_tiles_bytes(N, N) / 4);
        CUDA_CHECK(cudaMemsetAsync(pe.ptr, 0, pe.bytes(), nullptr));
        launch_rows_to_planes(eye.ptr, N, N, N, reinterpret_cast<unsigned char*>(pe.ptr), nullptr);
        GemmPlanesParams r;
        r.A = q.P; r.W = reinterpret_cast<unsigned char*>(pe.ptr); r.M = M; r.N = N; r.K = N; r.C = dC; r.ldc = ldc;
        r.variant = q.variant;
        launch_gemm_planes(r, nullptr);
      }
      CUDA_CHECK(cudaDeviceSynchronize());
    } else if (impl == 1) launch_gemm_simt(g, nullptr);
    else if (impl == 2) launch_gemm_tc(g, nullptr);
    else launch_gemm(g, nullptr);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaDeviceSynchronize());
  } catch (const std::exception& e) {
    MSB_LOGF("test gemm failed: %s", e.what());
    return MOONSHINE_ERROR_UNKNOWN;
  }
  return MOONSHINE_ERROR_NONE;
}

// ---- Part 2 stubs: symbols exist so bindings resolve at dlopen; every call
// reports failure (TTS / G2P / embeddings / catalogs are out of scope). ----
#define STUB_LOG(name) MSB_LOGF(name " is not implemented by moonshine-b200 (transcription path only)")
int32_t moonshine_create_embedding_model(const char*, uint32_t, const char*) { STUB_LOG("moonshine_create_embedding_model"); return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_create_embedding_model_from_memory(uint32_t, const char*, const char**, uint64_t, const uint8_t**, const uint64_t*, const moonshine_option_t*, uint64_t, int32_t) { STUB_LOG("moonshine_create_embedding_model_from_memory"); return MOONSHINE_ERROR_UNKNOWN; }
void moonshine_free_embedding_model(int32_t) {}
int32_t moonshine_calculate_embedding(int32_t, const char*, float**, uint64_t*, const char*) { return MOONSHINE_ERROR_UNKNOWN; }
void moonshine_free_embedding(float* e) { std::free(e); }
int32_t moonshine_calculate_embedding_distance(int32_t, const float*, const float*, uint64_t, float*) { return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_extract_speech_clip(const float*, uint64_t, int32_t, int32_t, const moonshine_option_t*, uint64_t, moonshine_speech_clip_t*) { STUB_LOG("moonshine_extract_speech_clip"); return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_create_tts_synthesizer_from_files(const char*, const char**, uint64_t, const moonshine_option_t*, uint64_t, int32_t) { STUB_LOG("moonshine_create_tts_synthesizer_from_files"); return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_create_tts_synthesizer_from_memory(const char*, const char**, const uint64_t, const uint8_t**, const uint64_t*, const moonshine_option_t*, uint64_t, int32_t) { STUB_LOG("moonshine_create_tts_synthesizer_from_memory"); return MOONSHINE_ERROR_UNKNOWN; }
void moonshine_free_tts_synthesizer(int32_t) {}
int32_t moonshine_get_g2p_dependencies(const char*, const moonshine_option_t*, uint64_t, char**) { return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_get_tts_dependencies(const char*, const moonshine_option_t*, uint64_t, char**) { return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_get_tts_voices(const char*, const moonshine_option_t*, uint64_t, char**) { return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_get_stt_dependencies(const char*, const moonshine_option_t*, uint64_t, char**) { return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_get_embedding_dependencies(const char*, const moonshine_option_t*, uint64_t, char**) { return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_get_diarization_dependencies(char**) { return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_get_stt_catalog(char**) { return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_get_embedding_catalog(char**) { return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_text_to_speech(int32_t, const char*, const moonshine_option_t*, uint64_t, float**, uint64_t*, int32_t*) { STUB_LOG("moonshine_text_to_speech"); return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_phonemes_to_speech(int32_t, const char*, const moonshine_option_t*, uint64_t, float**, uint64_t*, int32_t*) { STUB_LOG("moonshine_phonemes_to_speech"); return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_create_grapheme_to_phonemizer_from_files(const char*, const char**, uint64_t, const moonshine_option_t*, uint64_t, int32_t) { return MOONSHINE_ERROR_UNKNOWN; }
int32_t moonshine_create_grapheme_to_phonemizer_from_memory(const char*, const char**, const uint64_t, const uint8_t**, const uint64_t*, const moonshine_option_t*, uint64_t, int32_t) { return MOONSHINE_ERROR_UNKNOWN; }
void moonshine_free_grapheme_to_phonemizer(int32_t) {}
int32_t moonshine_text_to_phonemes(int32_t, const char*, const moonshine_option_t*, uint64_t, const char**, uint64_t*) { return MOONSHINE_ERROR_UNKNOWN; }

}  // extern "C"
