/* moonshine_b200.h -- C ABI of the B200-native Moonshine transcription runtime.
 *
 * Part 1 re-declares, name for name and byte for byte, the transcription slice
 * of the reference's `core/moonshine-c-api.h` (cited per symbol as
 * c-api.h:LINE) so that every binding that links against libmoonshine
 * (ctypes `language-bindings/python/src/moonshine_voice/moonshine_api.py`,
 * JNI, Swift module map, embind, `core/moonshine-cpp.h`) resolves the same
 * symbols with the same struct layouts (24 / 40 / 88 / 16 bytes).
 * Part 2 declares the remaining reference exports; they exist so `dlopen`
 * succeeds but report MOONSHINE_ERROR_UNKNOWN (out of scope: TTS, G2P,
 * embeddings, catalogs, speech clips).
 * Part 3 is additive: batched and device-resident entry points the reference
 * does not have (it never runs batch > 1: core/transcriber.cpp:997-1081).
 *
 * No ONNX Runtime and no CPU fallback sit behind this ABI: load fails with
 * MOONSHINE_ERROR_UNKNOWN when no sm_100a device is present.
 *
 * Model directory contract: the reference loads `encoder_model.ort` +
 * `decoder_model_merged.ort` (+ `tokenizer.bin`) (c-api.h:366-371); this
 * runtime loads `model.msw` (float weights, see moonshine_b200/weights.py)
 * + `tokenizer.bin` from the same directory / the same in-memory key list.
 */
#ifndef MOONSHINE_B200_H
#define MOONSHINE_B200_H

#include <stddef.h>
#include <stdint.h>

#define MOONSHINE_EXPORT __attribute__((visibility("default")))

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants (c-api.h:95-135) ---- */
#define MOONSHINE_HEADER_VERSION (30000)
#define MOONSHINE_FROM_MEMORY_REMOVED_VERSION (30000)
#define MOONSHINE_MODEL_ARCH_TINY (0)
#define MOONSHINE_MODEL_ARCH_BASE (1)
#define MOONSHINE_MODEL_ARCH_TINY_STREAMING (2)
#define MOONSHINE_MODEL_ARCH_BASE_STREAMING (3)
#define MOONSHINE_MODEL_ARCH_SMALL_STREAMING (4)
#define MOONSHINE_MODEL_ARCH_MEDIUM_STREAMING (5)
#define MOONSHINE_ERROR_NONE (0)
#define MOONSHINE_ERROR_UNKNOWN (-1)
#define MOONSHINE_ERROR_INVALID_HANDLE (-2)
#define MOONSHINE_ERROR_INVALID_ARGUMENT (-3)
#define MOONSHINE_FLAG_FORCE_UPDATE (1 << 0)
#define MOONSHINE_FLAG_SPELLING_MODE (1 << 1)

/* ---- structs (layout is ABI) ---- */
struct moonshine_option_t { /* c-api.h:146 */
  const char *name;
  const char *value;
};
struct transcript_word_t { /* c-api.h:203, 24 bytes */
  const char *text;
  float start;
  float end;
  float confidence;
};
struct speaker_span_t { /* c-api.h:221, 40 bytes */
  float start_time;
  float duration;
  uint64_t speaker_id;
  uint32_t speaker_index;
  uint64_t start_char;
  uint64_t end_char;
};
struct transcript_line_t { /* c-api.h:240-282, 88 bytes */
  const char *text;
  const float *audio_data;
  size_t audio_data_count;
  float start_time;
  float duration;
  uint64_t id;
  int8_t is_complete;
  int8_t is_updated;
  int8_t is_new;
  int8_t has_text_changed;
  int8_t have_speakers_changed;
  const struct speaker_span_t *speaker_spans;
  uint64_t speaker_span_count;
  uint32_t last_transcription_latency_ms;
  const struct transcript_word_t *words;
  uint64_t word_count;
};
struct transcript_t { /* c-api.h:285, 16 bytes */
  struct transcript_line_t *lines;
  uint64_t line_count;
};

/* ---- Part 1: transcription path ---- */
MOONSHINE_EXPORT int32_t moonshine_get_version(void);                   /* c-api.h:295 */
MOONSHINE_EXPORT const char *moonshine_error_to_string(int32_t error);  /* c-api.h:299 */
MOONSHINE_EXPORT void moonshine_free_buffer(void *ptr);                 /* c-api.h:315 */
/* Key-term biasing exists only on the reference's streaming architectures;
   TINY/BASE return an error there too (c-api.h:326-331,354-359). */
MOONSHINE_EXPORT int32_t moonshine_transcriber_set_keyterms(int32_t transcriber_handle,
                                                            const char *keyterms);
MOONSHINE_EXPORT int32_t moonshine_transcriber_set_context(int32_t transcriber_handle,
                                                           const char *context, int32_t max_terms);
MOONSHINE_EXPORT const char *moonshine_transcript_to_string(
    const struct transcript_t *transcript); /* c-api.h:364 */
/* Replaces c-api.h:453.  `path` holds model.msw + tokenizer.bin.  Returns a
   handle >= 0 or a negative error code.  Unknown option names fail the load
   (moonshine-c-api.cpp:193-196). */
MOONSHINE_EXPORT int32_t moonshine_load_transcriber_from_files(
    const char *path, uint32_t model_arch, const struct moonshine_option_t *options,
    uint64_t options_count, int32_t moonshine_version);
/* c-api.h:482: refused (-3) for callers built against header >= 30000;
   older callers hand over ORT graph bytes, which this runtime cannot run. */
MOONSHINE_EXPORT int32_t moonshine_load_transcriber_from_memory(
    const uint8_t *encoder_model_data, size_t encoder_model_data_size,
    const uint8_t *decoder_model_data, size_t decoder_model_data_size,
    const uint8_t *tokenizer_data, size_t tokenizer_data_size,
    const uint8_t *spelling_model_data, size_t spelling_model_data_size, uint32_t model_arch,
    const struct moonshine_option_t *options, uint64_t options_count, int32_t moonshine_version);
/* c-api.h:532.  Recognised keys: the reference's canonical file names (accepted
   and ignored when this runtime has no use for them) plus `model.msw`.
   Unknown key => -3.  Buffers are parsed during the call and need not
   outlive it. */
MOONSHINE_EXPORT int32_t moonshine_load_transcriber_from_memory_files(
    const char **filenames, const uint8_t **memory, const uint64_t *memory_sizes,
    uint64_t file_count, uint32_t model_arch, const struct moonshine_option_t *options,
    uint64_t options_count, int32_t moonshine_version);
MOONSHINE_EXPORT void moonshine_free_transcriber(int32_t transcriber_handle); /* c-api.h:541 */
/* c-api.h:578.  Output is owned by the transcriber, valid until the next call
   on it or until it is freed. */
MOONSHINE_EXPORT int32_t moonshine_transcribe_without_streaming(
    int32_t transcriber_handle, float *audio_data, uint64_t audio_length, int32_t sample_rate,
    uint32_t flags, struct transcript_t **out_transcript);
MOONSHINE_EXPORT int32_t moonshine_create_stream(int32_t transcriber_handle, uint32_t flags); /* :664 */
MOONSHINE_EXPORT int32_t moonshine_free_stream(int32_t transcriber_handle, int32_t stream_handle);
MOONSHINE_EXPORT int32_t moonshine_start_stream(int32_t transcriber_handle, int32_t stream_handle);
MOONSHINE_EXPORT int32_t moonshine_stop_stream(int32_t transcriber_handle, int32_t stream_handle);
MOONSHINE_EXPORT int32_t moonshine_transcribe_add_audio_to_stream(
    int32_t transcriber_handle, int32_t stream_handle, const float *new_audio_data,
    uint64_t audio_length, int32_t sample_rate, uint32_t flags); /* c-api.h:721 */
MOONSHINE_EXPORT int32_t moonshine_transcribe_stream(int32_t transcriber_handle,
                                                     int32_t stream_handle, uint32_t flags,
                                                     struct transcript_t **out_transcript); /* :754 */

/* ---- Part 2: exports kept for link compatibility (always fail) ---- */
struct moonshine_speech_clip_t {
  float *audio_data;
  uint64_t audio_length;
  float start_time;
  float speech_duration;
  int32_t is_complete;
  char *transcript;
};
MOONSHINE_EXPORT int32_t moonshine_create_embedding_model(const char *, uint32_t, const char *);
MOONSHINE_EXPORT int32_t moonshine_create_embedding_model_from_memory(
    uint32_t, const char *, const char **, uint64_t, const uint8_t **, const uint64_t *,
    const struct moonshine_option_t *, uint64_t, int32_t);
MOONSHINE_EXPORT void moonshine_free_embedding_model(int32_t);
MOONSHINE_EXPORT int32_t moonshine_calculate_embedding(int32_t, const char *, float **, uint64_t *,
                                                       const char *);
MOONSHINE_EXPORT void moonshine_free_embedding(float *);
MOONSHINE_EXPORT int32_t moonshine_calculate_embedding_distance(int32_t, const float *,
                                                                const float *, uint64_t, float *);
MOONSHINE_EXPORT int32_t moonshine_extract_speech_clip(const float *, uint64_t, int32_t, int32_t,
                                                       const struct moonshine_option_t *, uint64_t,
                                                       struct moonshine_speech_clip_t *);
MOONSHINE_EXPORT int32_t moonshine_create_tts_synthesizer_from_files(
    const char *, const char **, uint64_t, const struct moonshine_option_t *, uint64_t, int32_t);
MOONSHINE_EXPORT int32_t moonshine_create_tts_synthesizer_from_memory(
    const char *, const char **, const uint64_t, const uint8_t **, const uint64_t *,
    const struct moonshine_option_t *, uint64_t, int32_t);
MOONSHINE_EXPORT void moonshine_free_tts_synthesizer(int32_t);
MOONSHINE_EXPORT int32_t moonshine_get_g2p_dependencies(const char *, const struct moonshine_option_t *,
                                                        uint64_t, char **);
MOONSHINE_EXPORT int32_t moonshine_get_tts_dependencies(const char *, const struct moonshine_option_t *,
                                                        uint64_t, char **);
MOONSHINE_EXPORT int32_t moonshine_get_tts_voices(const char *, const struct moonshine_option_t *,
                                                  uint64_t, char **);
MOONSHINE_EXPORT int32_t moonshine_get_stt_dependencies(const char *, const struct moonshine_option_t *,
                                                        uint64_t, char **);
MOONSHINE_EXPORT int32_t moonshine_get_embedding_dependencies(const char *,
                                                              const struct moonshine_option_t *,
                                                              uint64_t, char **);
MOONSHINE_EXPORT int32_t moonshine_get_diarization_dependencies(char **);
MOONSHINE_EXPORT int32_t moonshine_get_stt_catalog(char **);
MOONSHINE_EXPORT int32_t moonshine_get_embedding_catalog(char **);
MOONSHINE_EXPORT int32_t moonshine_text_to_speech(int32_t, const char *,
                                                  const struct moonshine_option_t *, uint64_t,
                                                  float **, uint64_t *, int32_t *);
MOONSHINE_EXPORT int32_t moonshine_phonemes_to_speech(int32_t, const char *,
                                                      const struct moonshine_option_t *, uint64_t,
                                                      float **, uint64_t *, int32_t *);
MOONSHINE_EXPORT int32_t moonshine_create_grapheme_to_phonemizer_from_files(
    const char *, const char **, uint64_t, const struct moonshine_option_t *, uint64_t, int32_t);
MOONSHINE_EXPORT int32_t moonshine_create_grapheme_to_phonemizer_from_memory(
    const char *, const char **, const uint64_t, const uint8_t **, const uint64_t *,
    const struct moonshine_option_t *, uint64_t, int32_t);
MOONSHINE_EXPORT void moonshine_free_grapheme_to_phonemizer(int32_t);
MOONSHINE_EXPORT int32_t moonshine_text_to_phonemes(int32_t, const char *,
                                                    const struct moonshine_option_t *, uint64_t,
                                                    const char **, uint64_t *);

/* ---- Part 3: additive entry points (not in the reference) ---- */
/* Batched moonshine_transcribe_without_streaming: utterance i is
   audio[i][0 .. lengths[i]).  Result i is exactly what the single call returns
   for that clip.  *out_transcripts points at `count` transcript_t structs owned
   by the transcriber (valid until its next call). */
MOONSHINE_EXPORT int32_t moonshine_transcribe_batch_without_streaming(
    int32_t transcriber_handle, const float *const *audio, const uint64_t *lengths, uint64_t count,
    int32_t sample_rate, uint32_t flags, struct transcript_t **out_transcripts);
/* Device-resident 16 kHz PCM (row i at d_pcm + i * stride, stride % 4 == 0):
   encoder + greedy decode only, no segmentation / detokenisation.
   out_tokens[i * out_stride ..] receives the ids (start token first),
   out_counts[i] how many.  The model's MoonshineModel::transcribe contract
   (core/moonshine-model.cpp:215) for a batch. */
MOONSHINE_EXPORT int32_t moonshine_b200_transcribe_device(
    int32_t transcriber_handle, const float *d_pcm, int64_t stride, const uint64_t *lengths,
    uint64_t count, int32_t *out_tokens, int32_t out_stride, int32_t *out_counts);
/* cudaStream_t the transcriber launches on (for CUDA-event timing). */
MOONSHINE_EXPORT void *moonshine_b200_get_stream(int32_t transcriber_handle);
/* Enables per-stage CUDA-event timing; out[0..7] = frontend_ms, encoder_ms,
   cross_kv_ms, decode_ms, decode_steps, kernel_launches, weight_bytes, decoder kernel version (1-4). */
MOONSHINE_EXPORT int32_t moonshine_b200_set_timing(int32_t transcriber_handle, int32_t enabled);
MOONSHINE_EXPORT int32_t moonshine_b200_last_timings(int32_t transcriber_handle, double *out8);
/* Parity hook: runs host PCM utterances and returns intermediate tensors.
   enc_out:   [sum_i frames_i][D] floats (may be NULL), enc_frames[count]
   forced:    [count][forced_stride] teacher-forced ids incl. start (may be NULL)
   logits:    [logits_steps][count][V] (may be NULL)
   out_tokens/out_counts as above (ids are the model's own argmax). */
MOONSHINE_EXPORT int32_t moonshine_b200_debug_run(
    int32_t transcriber_handle, const float *const *audio, const uint64_t *lengths, uint64_t count,
    float *enc_out, uint64_t enc_out_capacity, int32_t *enc_frames, const int32_t *forced,
    int32_t forced_stride, float *logits, int32_t logits_steps, int32_t *out_tokens,
    int32_t out_stride, int32_t *out_counts);
/* Verify-then-continue decode of a batch (reference: MoonshineStreamingModel::decode_full with speculative_tokens,
   core/moonshine-streaming-model.cpp:1192-1397; the multi-token decoder run underneath it is decode_tokens'
   run_decoder_with_cross_kv, :1136-1190).  drafts: [count][draft_stride] ids without BOS / EOS, draft_lens[count]
   (0 = plain greedy for that utterance).  Eight draft positions per utterance are verified per decoder launch; the
   returned ids (BOS first, EOS kept when emitted) equal the greedy decode of the same audio whatever the draft holds.
   out_launches (may be NULL): decoder launches the call took. */
MOONSHINE_EXPORT int32_t moonshine_b200_decode_with_drafts(
    int32_t transcriber_handle, const float *const *audio, const uint64_t *lengths, uint64_t count,
    const int32_t *drafts, int32_t draft_stride, const int32_t *draft_lens, int32_t *out_tokens,
    int32_t out_stride, int32_t *out_counts, int32_t *out_launches);
/* Teacher-forced multi-token decoder runs (reference: MoonshineStreamingModel::decode_tokens,
   core/moonshine-streaming-model.cpp:1136-1190, batched): tokens [count][tokens_stride] ids incl. the start id; the
   first n_steps positions of every utterance go through the decoder rows_per_launch (2..16) positions per launch;
   logits_out [n_steps][count][V] receives the logits of every position (rows beyond an utterance's budget stay 0). */
MOONSHINE_EXPORT int32_t moonshine_b200_decode_tokens(
    int32_t transcriber_handle, const float *const *audio, const uint64_t *lengths, uint64_t count,
    const int32_t *tokens, int32_t tokens_stride, int32_t n_steps, int32_t rows_per_launch, float *logits_out);
/* Host-only self-test of the persistent worker pool behind the batch path (segmentation, staging): returns 0 when every
   item of every parallel_for ran exactly once under `callers` concurrent callers and an exception reached its caller. */
MOONSHINE_EXPORT int32_t moonshine_b200_debug_pool_selftest(int32_t callers, int32_t n, int32_t rounds);
/* Parity hook for the streaming architectures: when enabled, moonshine_b200_debug_run /
   moonshine_b200_transcribe_device treat each utterance as a NON-final update of its segment
   (the encoder's look-ahead features are held back, core/moonshine-streaming-model.cpp:624-626). */
MOONSHINE_EXPORT int32_t moonshine_b200_debug_stream_partial(int32_t transcriber_handle, int32_t enabled);
/* Host-only parity hooks (no GPU): this library's detokeniser (reference: BinTokenizer::tokens_to_text,
   core/bin-tokenizer/bin-tokenizer.cpp:406-425) and resampler (core/resampler.cpp:5-86), so CPU tests can
   compare them with the reference's own compiled sources (oracle/_ref).  Return the full output length. */
MOONSHINE_EXPORT int64_t moonshine_b200_debug_tokens_to_text(const uint8_t *tokenizer, uint64_t tokenizer_size,
                                                             const int32_t *ids, int32_t n, char *out,
                                                             int64_t cap);
/* Word alignment (reference: align_words, core/word-alignment.cpp:181-394) on caller-supplied cross-attention
   [heads_total][steps][frames]; word texts are written NUL-separated.  Returns the word count. */
MOONSHINE_EXPORT int32_t moonshine_b200_debug_align_words(
    const uint8_t *tokenizer, uint64_t tokenizer_size, const float *xattn, int32_t heads_total, int32_t steps,
    int32_t frames, const int32_t *tokens, int32_t n_tokens, float time_per_frame, float *starts, float *ends,
    char *text_out, int64_t text_cap, int32_t max_words);
/* Text -> ids (reference: BinTokenizer::text_to_tokens, bin-tokenizer.cpp:274-404; bpe != 0 = the streaming
   models' byte-pair mode).  Returns the id count, -1 on failure. */
MOONSHINE_EXPORT int32_t moonshine_b200_debug_text_to_tokens(const uint8_t *tokenizer, uint64_t tokenizer_size,
                                                             const char *text, int32_t bpe, int32_t *out,
                                                             int32_t cap);
/* Key-term biaser (reference: ContextBiaser, core/context-biaser.cpp): trie from the given sequences, walked
   along `path`, bonuses added to `logits` in place. */
MOONSHINE_EXPORT int32_t moonshine_b200_debug_biaser_apply(const int32_t *seqs, const int32_t *seq_lens,
                                                           int32_t n_seqs, float boost, const int32_t *path,
                                                           int32_t n_path, float *logits, int32_t vocab);
/* The same bonuses through the sparse form the on-device path uploads (dense root bonuses + per-step pairs). */
MOONSHINE_EXPORT int32_t moonshine_b200_debug_biaser_apply_sparse(const int32_t *seqs, const int32_t *seq_lens,
                                                                  int32_t n_seqs, float boost, const int32_t *path,
                                                                  int32_t n_path, float *logits, int32_t vocab);
/* Key terms of a passage (reference: ContextExtractor::extract, core/context-extractor.cpp), NUL-separated. */
MOONSHINE_EXPORT int32_t moonshine_b200_debug_extract_terms(const uint8_t *tokenizer, uint64_t tokenizer_size,
                                                            const char *context, int32_t max_terms, char *out,
                                                            int64_t cap);
MOONSHINE_EXPORT int64_t moonshine_b200_debug_resample(const float *in, int64_t n, float in_rate,
                                                       float out_rate, float *out, int64_t cap);
/* Microbenchmark of the decoder's operand ring (cp.async.bulk + mbarrier stages): milliseconds for `grid` CTAs
   to each stream bytes_per_cta; shared_src != 0 makes all CTAs read the same 2 MB (L2-resident) span. */
MOONSHINE_EXPORT float moonshine_b200_test_ring_bandwidth(int64_t bytes_per_cta, int32_t stage_bytes, int32_t stages,
                                                          int32_t nsub, int32_t shared_src, int32_t grid);
/* Standalone grouped-GEMM hook used by the kernel unit tests (device pointers). */
MOONSHINE_EXPORT int32_t moonshine_b200_test_gemm(const float *dA, const float *dW, float *dC,
                                                  int32_t M, int32_t N, int32_t K, int32_t lda,
                                                  int32_t ldw, int32_t ldc, const float *d_bias,
                                                  int32_t act, int32_t accumulate, int32_t impl);

#ifdef __cplusplus
}
#endif
#endif /* MOONSHINE_B200_H */
