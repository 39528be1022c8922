"""ctypes binding over libmoonshine.so -- the host-side mirror of the reference's
Python binding for the transcription path
(language-bindings/python/src/moonshine_voice/{moonshine_api,transcriber}.py):
same class / method names (`Transcriber.transcribe_without_streaming`,
`create_stream`, `Stream.start/add_audio/update_transcription/stop`), same
struct layouts and the same struct-size guard (24 / 40 / 88 / 16 bytes).

There is no fallback: if the CUDA library is missing this module raises.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from enum import IntEnum
from typing import Dict, List, Optional, Sequence

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmoonshine.so")

MOONSHINE_HEADER_VERSION = 30000
MOONSHINE_FLAG_FORCE_UPDATE = 1 << 0


class ModelArch(IntEnum):
    TINY = 0
    BASE = 1
    TINY_STREAMING = 2
    BASE_STREAMING = 3
    SMALL_STREAMING = 4
    MEDIUM_STREAMING = 5
    # reduced-size architectures for unit tests (not in the reference)
    TEST = 100
    TEST2 = 101
    TEST_STREAMING = 102
    TEST_STREAMING2 = 103


class MoonshineError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{message} (error {code})")
        self.code = code


class TranscriptWordC(ctypes.Structure):
    _fields_ = [("text", ctypes.c_char_p), ("start", ctypes.c_float), ("end", ctypes.c_float),
                ("confidence", ctypes.c_float)]


class SpeakerSpanC(ctypes.Structure):
    _fields_ = [("start_time", ctypes.c_float), ("duration", ctypes.c_float),
                ("speaker_id", ctypes.c_uint64), ("speaker_index", ctypes.c_uint32),
                ("start_char", ctypes.c_uint64), ("end_char", ctypes.c_uint64)]


class TranscriptLineC(ctypes.Structure):
    _fields_ = [
        ("text", ctypes.c_char_p), ("audio_data", ctypes.POINTER(ctypes.c_float)),
        ("audio_data_count", ctypes.c_size_t), ("start_time", ctypes.c_float),
        ("duration", ctypes.c_float), ("id", ctypes.c_uint64), ("is_complete", ctypes.c_int8),
        ("is_updated", ctypes.c_int8), ("is_new", ctypes.c_int8), ("has_text_changed", ctypes.c_int8),
        ("have_speakers_changed", ctypes.c_int8), ("speaker_spans", ctypes.POINTER(SpeakerSpanC)),
        ("speaker_span_count", ctypes.c_uint64), ("last_transcription_latency_ms", ctypes.c_uint32),
        ("words", ctypes.POINTER(TranscriptWordC)), ("word_count", ctypes.c_uint64),
    ]


class TranscriptC(ctypes.Structure):
    _fields_ = [("lines", ctypes.POINTER(TranscriptLineC)), ("line_count", ctypes.c_uint64)]


class TranscriberOptionC(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("value", ctypes.c_char_p)]


def _require_struct_size(name, cls, expected):
    actual = ctypes.sizeof(cls)
    if actual != expected:
        raise ImportError(f"moonshine_b200 ABI mismatch: {name} is {actual} bytes, the C ABI expects {expected}")


_require_struct_size("TranscriptWordC", TranscriptWordC, 24)
_require_struct_size("SpeakerSpanC", SpeakerSpanC, 40)
_require_struct_size("TranscriptLineC", TranscriptLineC, 88)
_require_struct_size("TranscriptC", TranscriptC, 16)


# every symbol include/moonshine_b200.h declares
EXPORTED_SYMBOLS = [
    "moonshine_get_version", "moonshine_error_to_string", "moonshine_free_buffer",
    "moonshine_transcriber_set_keyterms", "moonshine_transcriber_set_context",
    "moonshine_transcript_to_string", "moonshine_load_transcriber_from_files",
    "moonshine_load_transcriber_from_memory", "moonshine_load_transcriber_from_memory_files",
    "moonshine_free_transcriber", "moonshine_transcribe_without_streaming", "moonshine_create_stream",
    "moonshine_free_stream", "moonshine_start_stream", "moonshine_stop_stream",
    "moonshine_transcribe_add_audio_to_stream", "moonshine_transcribe_stream",
    "moonshine_create_embedding_model", "moonshine_create_embedding_model_from_memory",
    "moonshine_free_embedding_model", "moonshine_calculate_embedding", "moonshine_free_embedding",
    "moonshine_calculate_embedding_distance", "moonshine_extract_speech_clip",
    "moonshine_create_tts_synthesizer_from_files", "moonshine_create_tts_synthesizer_from_memory",
    "moonshine_free_tts_synthesizer", "moonshine_get_g2p_dependencies", "moonshine_get_tts_dependencies",
    "moonshine_get_tts_voices", "moonshine_get_stt_dependencies", "moonshine_get_embedding_dependencies",
    "moonshine_get_diarization_dependencies", "moonshine_get_stt_catalog", "moonshine_get_embedding_catalog",
    "moonshine_text_to_speech", "moonshine_phonemes_to_speech",
    "moonshine_create_grapheme_to_phonemizer_from_files", "moonshine_create_grapheme_to_phonemizer_from_memory",
    "moonshine_free_grapheme_to_phonemizer", "moonshine_text_to_phonemes",
    "moonshine_transcribe_batch_without_streaming", "moonshine_b200_transcribe_device",
    "moonshine_b200_get_stream", "moonshine_b200_set_timing", "moonshine_b200_last_timings",
    "moonshine_b200_debug_run", "moonshine_b200_debug_stream_partial", "moonshine_b200_decode_with_drafts", "moonshine_b200_decode_tokens", "moonshine_b200_debug_pool_selftest",
    "moonshine_b200_debug_tokens_to_text", "moonshine_b200_debug_resample", "moonshine_b200_debug_align_words", "moonshine_b200_test_ring_bandwidth",
    "moonshine_b200_debug_text_to_tokens", "moonshine_b200_debug_biaser_apply", "moonshine_b200_debug_biaser_apply_sparse", "moonshine_b200_debug_extract_terms", "moonshine_b200_test_gemm",
]

_lib = None


def lib_path() -> str:
    return _LIB_PATH


def load_library() -> ctypes.CDLL:
    """dlopen libmoonshine.so and bind argtypes (no compute happens here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing: build it with `python -m moonshine_b200.build` "
                          "(moonshine_b200 has no CPU fallback)")
    lib = ctypes.CDLL(_LIB_PATH)
    c = ctypes
    f32p, u64p, i32p = c.POINTER(c.c_float), c.POINTER(c.c_uint64), c.POINTER(c.c_int32)
    optp = c.POINTER(TranscriberOptionC)
    tpp = c.POINTER(c.POINTER(TranscriptC))
    lib.moonshine_get_version.restype = c.c_int32
    lib.moonshine_error_to_string.restype = c.c_char_p
    lib.moonshine_error_to_string.argtypes = [c.c_int32]
    lib.moonshine_transcript_to_string.restype = c.c_char_p
    lib.moonshine_transcript_to_string.argtypes = [c.POINTER(TranscriptC)]
    lib.moonshine_load_transcriber_from_files.restype = c.c_int32
    lib.moonshine_load_transcriber_from_files.argtypes = [c.c_char_p, c.c_uint32, optp, c.c_uint64, c.c_int32]
    lib.moonshine_load_transcriber_from_memory_files.restype = c.c_int32
    lib.moonshine_load_transcriber_from_memory_files.argtypes = [
        c.POINTER(c.c_char_p), c.POINTER(c.c_void_p), u64p, c.c_uint64, c.c_uint32, optp, c.c_uint64, c.c_int32]
    lib.moonshine_load_transcriber_from_memory.restype = c.c_int32
    lib.moonshine_load_transcriber_from_memory.argtypes = [
        c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t,
        c.c_uint32, optp, c.c_uint64, c.c_int32]
    lib.moonshine_free_transcriber.restype = None
    lib.moonshine_free_transcriber.argtypes = [c.c_int32]
    lib.moonshine_transcribe_without_streaming.restype = c.c_int32
    lib.moonshine_transcribe_without_streaming.argtypes = [c.c_int32, f32p, c.c_uint64, c.c_int32, c.c_uint32, tpp]
    lib.moonshine_transcribe_batch_without_streaming.restype = c.c_int32
    lib.moonshine_transcribe_batch_without_streaming.argtypes = [
        c.c_int32, c.POINTER(f32p), u64p, c.c_uint64, c.c_int32, c.c_uint32, c.POINTER(c.POINTER(TranscriptC))]
    for name in ("moonshine_create_stream",):
        getattr(lib, name).restype = c.c_int32
        getattr(lib, name).argtypes = [c.c_int32, c.c_uint32]
    for name in ("moonshine_free_stream", "moonshine_start_stream", "moonshine_stop_stream"):
        getattr(lib, name).restype = c.c_int32
        getattr(lib, name).argtypes = [c.c_int32, c.c_int32]
    lib.moonshine_transcribe_add_audio_to_stream.restype = c.c_int32
    lib.moonshine_transcribe_add_audio_to_stream.argtypes = [c.c_int32, c.c_int32, f32p, c.c_uint64, c.c_int32, c.c_uint32]
    lib.moonshine_transcribe_stream.restype = c.c_int32
    lib.moonshine_transcribe_stream.argtypes = [c.c_int32, c.c_int32, c.c_uint32, tpp]
    lib.moonshine_transcriber_set_keyterms.restype = c.c_int32
    lib.moonshine_transcriber_set_keyterms.argtypes = [c.c_int32, c.c_char_p]
    lib.moonshine_transcriber_set_context.restype = c.c_int32
    lib.moonshine_transcriber_set_context.argtypes = [c.c_int32, c.c_char_p, c.c_int32]
    lib.moonshine_b200_transcribe_device.restype = c.c_int32
    lib.moonshine_b200_transcribe_device.argtypes = [c.c_int32, c.c_void_p, c.c_int64, u64p, c.c_uint64, i32p, c.c_int32, i32p]
    lib.moonshine_b200_get_stream.restype = c.c_void_p
    lib.moonshine_b200_get_stream.argtypes = [c.c_int32]
    lib.moonshine_b200_set_timing.restype = c.c_int32
    lib.moonshine_b200_set_timing.argtypes = [c.c_int32, c.c_int32]
    lib.moonshine_b200_last_timings.restype = c.c_int32
    lib.moonshine_b200_last_timings.argtypes = [c.c_int32, c.POINTER(c.c_double)]
    lib.moonshine_b200_debug_tokens_to_text.restype = c.c_int64
    lib.moonshine_b200_debug_tokens_to_text.argtypes = [c.c_void_p, c.c_uint64, c.POINTER(c.c_int32), c.c_int32,
                                                        c.c_char_p, c.c_int64]
    lib.moonshine_b200_debug_align_words.restype = c.c_int32
    lib.moonshine_b200_debug_align_words.argtypes = [
        c.c_void_p, c.c_uint64, c.POINTER(c.c_float), c.c_int32, c.c_int32, c.c_int32, c.POINTER(c.c_int32),
        c.c_int32, c.c_float, c.POINTER(c.c_float), c.POINTER(c.c_float), c.c_char_p, c.c_int64, c.c_int32]
    lib.moonshine_b200_test_ring_bandwidth.restype = c.c_float
    lib.moonshine_b200_test_ring_bandwidth.argtypes = [c.c_int64, c.c_int32, c.c_int32, c.c_int32, c.c_int32, c.c_int32]
    lib.moonshine_b200_debug_text_to_tokens.restype = c.c_int32
    lib.moonshine_b200_debug_text_to_tokens.argtypes = [c.c_void_p, c.c_uint64, c.c_char_p, c.c_int32,
                                                        c.POINTER(c.c_int32), c.c_int32]
    lib.moonshine_b200_debug_biaser_apply.restype = c.c_int32
    lib.moonshine_b200_debug_biaser_apply.argtypes = [c.POINTER(c.c_int32), c.POINTER(c.c_int32), c.c_int32, c.c_float,
                                                      c.POINTER(c.c_int32), c.c_int32, c.POINTER(c.c_float), c.c_int32]
    lib.moonshine_b200_debug_biaser_apply_sparse.restype = c.c_int32
    lib.moonshine_b200_debug_biaser_apply_sparse.argtypes = lib.moonshine_b200_debug_biaser_apply.argtypes
    lib.moonshine_b200_debug_extract_terms.restype = c.c_int32
    lib.moonshine_b200_debug_extract_terms.argtypes = [c.c_void_p, c.c_uint64, c.c_char_p, c.c_int32, c.c_char_p, c.c_int64]
    lib.moonshine_b200_debug_resample.restype = c.c_int64
    lib.moonshine_b200_debug_resample.argtypes = [c.POINTER(c.c_float), c.c_int64, c.c_float, c.c_float,
                                                  c.POINTER(c.c_float), c.c_int64]
    lib.moonshine_b200_debug_stream_partial.restype = c.c_int32
    lib.moonshine_b200_debug_stream_partial.argtypes = [c.c_int32, c.c_int32]
    lib.moonshine_b200_debug_run.restype = c.c_int32
    lib.moonshine_b200_debug_run.argtypes = [
        c.c_int32, c.POINTER(f32p), u64p, c.c_uint64, f32p, c.c_uint64, i32p, i32p, c.c_int32, f32p,
        c.c_int32, i32p, c.c_int32, i32p]
    lib.moonshine_b200_decode_with_drafts.restype = c.c_int32
    lib.moonshine_b200_decode_with_drafts.argtypes = [
        c.c_int32, c.POINTER(f32p), u64p, c.c_uint64, i32p, c.c_int32, i32p, i32p, c.c_int32, i32p, i32p]
    lib.moonshine_b200_debug_pool_selftest.restype = c.c_int32
    lib.moonshine_b200_debug_pool_selftest.argtypes = [c.c_int32, c.c_int32, c.c_int32]
    lib.moonshine_b200_decode_tokens.restype = c.c_int32
    lib.moonshine_b200_decode_tokens.argtypes = [
        c.c_int32, c.POINTER(f32p), u64p, c.c_uint64, i32p, c.c_int32, c.c_int32, c.c_int32, f32p]
    lib.moonshine_b200_test_gemm.restype = c.c_int32
    lib.moonshine_b200_test_gemm.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p] + [c.c_int32] * 6 + [c.c_void_p] + [c.c_int32] * 3
    _lib = lib
    return lib


def _check(code: int, what: str):
    if code < 0:
        msg = load_library().moonshine_error_to_string(code).decode()
        raise MoonshineError(code, f"{what}: {msg}")


@dataclass
class WordTiming:
    """Mirror of the reference binding's word record (moonshine_voice/transcriber.py)."""
    word: str
    start: float
    end: float
    confidence: float = 1.0


@dataclass
class TranscriptLine:
    text: Optional[str]
    start_time: float
    duration: float
    line_id: int
    is_complete: bool
    is_updated: bool = False
    is_new: bool = False
    has_text_changed: bool = False
    audio_data: Optional[np.ndarray] = None
    last_transcription_latency_ms: int = 0
    words: Optional[List[WordTiming]] = None


@dataclass
class Transcript:
    lines: List[TranscriptLine] = field(default_factory=list)

    def __str__(self):
        return "\n".join(f"[{l.start_time:.2f}s] {l.text}" for l in self.lines)


def _options_array(options: Optional[Dict[str, str]]):
    options = options or {}
    arr = (TranscriberOptionC * max(len(options), 1))()
    keep = []
    for i, (k, v) in enumerate(options.items()):
        kb, vb = str(k).encode(), str(v).encode()
        keep += [kb, vb]
        arr[i].name, arr[i].value = kb, vb
    return arr, len(options), keep


def _as_f32(audio) -> np.ndarray:
    return np.ascontiguousarray(audio, dtype=np.float32)


def _parse_transcript(tc: TranscriptC) -> Transcript:
    out = Transcript()
    for i in range(tc.line_count):
        l = tc.lines[i]
        audio = None
        if l.audio_data and l.audio_data_count:
            audio = np.ctypeslib.as_array(l.audio_data, shape=(l.audio_data_count,)).copy()
        out.lines.append(TranscriptLine(
            text=l.text.decode("utf-8", errors="replace") if l.text is not None else None,
            start_time=l.start_time, duration=l.duration, line_id=l.id,
            is_complete=bool(l.is_complete), is_updated=bool(l.is_updated), is_new=bool(l.is_new),
            has_text_changed=bool(l.has_text_changed), audio_data=audio,
            last_transcription_latency_ms=l.last_transcription_latency_ms,
            words=[WordTiming(l.words[k].text.decode("utf-8", errors="replace"), l.words[k].start, l.words[k].end,
                              l.words[k].confidence) for k in range(l.word_count)] if l.words and l.word_count else None))
    return out


class Transcriber:
    """Loads a model directory (model.msw + tokenizer.bin) onto the GPU."""

    def __init__(self, model_path: Optional[str] = None, model_arch: ModelArch = ModelArch.TINY,
                 options: Optional[Dict[str, str]] = None, *, memory_files: Optional[Dict[str, bytes]] = None):
        self._lib = load_library()
        self._handle = -1
        arr, n, keep = _options_array(options)
        if memory_files is not None:
            names = list(memory_files.keys())
            cn = (ctypes.c_char_p * len(names))(*[s.encode() for s in names])
            bufs = [np.frombuffer(memory_files[k], dtype=np.uint8) for k in names]
            cm = (ctypes.c_void_p * len(names))(*[b.ctypes.data for b in bufs])
            cs = (ctypes.c_uint64 * len(names))(*[b.size for b in bufs])
            h = self._lib.moonshine_load_transcriber_from_memory_files(
                cn, cm, cs, len(names), int(model_arch), arr, n, MOONSHINE_HEADER_VERSION)
        else:
            h = self._lib.moonshine_load_transcriber_from_files(
                str(model_path).encode(), int(model_arch), arr, n, MOONSHINE_HEADER_VERSION)
        _check(h, "Failed to load transcriber")
        self._handle = h
        self._default_stream = None

    # -- lifetime --
    def close(self):
        if getattr(self, "_handle", -1) >= 0:
            self._lib.moonshine_free_transcriber(self._handle)
            self._handle = -1

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> int:
        return self._handle

    def get_version(self) -> int:
        return self._lib.moonshine_get_version()

    # -- non-streaming --
    def transcribe_without_streaming(self, audio_data, sample_rate: int = 16000, flags: int = 0) -> Transcript:
        a = _as_f32(audio_data)
        out = ctypes.POINTER(TranscriptC)()
        rc = self._lib.moonshine_transcribe_without_streaming(
            self._handle, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), a.size, sample_rate, flags,
            ctypes.byref(out))
        _check(rc, "Failed to transcribe")
        return _parse_transcript(out.contents)

    def transcribe_batch_without_streaming(self, audios: Sequence, sample_rate: int = 16000,
                                           flags: int = 0) -> List[Transcript]:
        """Additive: one call for many utterances (host PCM)."""
        arrs = [_as_f32(a) for a in audios]
        n = len(arrs)
        ptrs = (ctypes.POINTER(ctypes.c_float) * max(n, 1))(
            *[a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) for a in arrs])
        lens = (ctypes.c_uint64 * max(n, 1))(*[a.size for a in arrs])
        out = ctypes.POINTER(TranscriptC)()
        rc = self._lib.moonshine_transcribe_batch_without_streaming(
            self._handle, ptrs, lens, n, sample_rate, flags, ctypes.byref(out))
        _check(rc, "Failed to transcribe batch")
        return [_parse_transcript(out[i]) for i in range(n)]

    def set_keyterms(self, keyterms):
        s = ",".join(keyterms) if keyterms else ""
        _check(self._lib.moonshine_transcriber_set_keyterms(self._handle, s.encode()), "Failed to set keyterms")

    def set_context(self, context: str, max_terms: int = 0):
        """Key terms picked out of a free-form passage (reference: moonshine_transcriber_set_context)."""
        _check(self._lib.moonshine_transcriber_set_context(self._handle, (context or "").encode("utf-8"), int(max_terms)),
               "Failed to set context")

    # -- streaming --
    def create_stream(self, flags: int = 0) -> "Stream":
        h = self._lib.moonshine_create_stream(self._handle, flags)
        _check(h, "Failed to create stream")
        return Stream(self, h)

    # -- additive: model-level calls --
    def transcribe_device(self, d_pcm_ptr: int, stride: int, lengths: Sequence[int], max_tokens: int = 128):
        """Device-resident PCM (pointer from e.g. torch.Tensor.data_ptr()). Returns list of id lists."""
        n = len(lengths)
        lens = (ctypes.c_uint64 * n)(*[int(x) for x in lengths])
        toks = np.zeros((n, max_tokens), np.int32)
        cnt = np.zeros(n, np.int32)
        rc = self._lib.moonshine_b200_transcribe_device(
            self._handle, ctypes.c_void_p(d_pcm_ptr), stride, lens, n,
            toks.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), max_tokens,
            cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        _check(rc, "Failed to transcribe device batch")
        return [toks[i, :min(cnt[i], max_tokens)].tolist() for i in range(n)]

    def cuda_stream_ptr(self) -> int:
        return int(self._lib.moonshine_b200_get_stream(self._handle) or 0)

    def set_timing(self, on: bool = True):
        _check(self._lib.moonshine_b200_set_timing(self._handle, 1 if on else 0), "set_timing")

    def last_timings(self) -> Dict[str, float]:
        out = (ctypes.c_double * 8)()
        _check(self._lib.moonshine_b200_last_timings(self._handle, out), "last_timings")
        keys = ["frontend_ms", "encoder_ms", "cross_kv_ms", "decode_ms", "decode_steps", "kernel_launches",
                "weight_bytes", "decoder_version"]
        return {k: out[i] for i, k in enumerate(keys)}

    def debug_stream_partial(self, enabled: bool):
        """Streaming archs: model-level calls act as a non-final update (look-ahead features held back)."""
        _check(self._lib.moonshine_b200_debug_stream_partial(self._handle, int(bool(enabled))),
               "debug_stream_partial failed")

    def debug_run(self, audios: Sequence, dim: int, vocab: int, forced: Optional[np.ndarray] = None,
                  logits_steps: int = 0, want_encoder: bool = True, max_tokens: int = 128):
        """Parity hook: returns (enc_out list per utterance, logits[steps,B,V] or None, tokens list)."""
        arrs = [_as_f32(a) for a in audios]
        n = len(arrs)
        ptrs = (ctypes.POINTER(ctypes.c_float) * n)(*[a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) for a in arrs])
        lens = (ctypes.c_uint64 * n)(*[a.size for a in arrs])
        frames = np.zeros(n, np.int32)
        cap = sum((a.size // 320 + 2) for a in arrs) * dim
        enc = np.zeros(cap, np.float32) if want_encoder else None
        lg = np.zeros((logits_steps, n, vocab), np.float32) if logits_steps > 0 else None
        toks = np.zeros((n, max_tokens), np.int32)
        cnt = np.zeros(n, np.int32)
        f32p, i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
        fptr, fstride = None, 0
        if forced is not None:
            forced = np.ascontiguousarray(forced, np.int32)
            fptr, fstride = forced.ctypes.data_as(i32p), forced.shape[1]
        rc = self._lib.moonshine_b200_debug_run(
            self._handle, ptrs, lens, n, enc.ctypes.data_as(f32p) if enc is not None else None, cap,
            frames.ctypes.data_as(i32p), fptr, fstride, lg.ctypes.data_as(f32p) if lg is not None else None,
            logits_steps, toks.ctypes.data_as(i32p), max_tokens, cnt.ctypes.data_as(i32p))
        _check(rc, "debug_run failed")
        encs = []
        if enc is not None:
            o = 0
            for i in range(n):
                encs.append(enc[o:o + frames[i] * dim].reshape(frames[i], dim).copy())
                o += frames[i] * dim
        return encs, lg, [toks[i, :min(cnt[i], max_tokens)].tolist() for i in range(n)]


    def decode_with_drafts(self, audios: Sequence, drafts: Sequence[Sequence[int]], max_tokens: int = 128):
        """Verify-then-continue decode (reference: decode_full with speculative tokens).  Returns (ids per utterance,
        decoder launches)."""
        arrs = [_as_f32(a) for a in audios]
        n = len(arrs)
        ptrs = (ctypes.POINTER(ctypes.c_float) * n)(*[a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) for a in arrs])
        lens = (ctypes.c_uint64 * n)(*[a.size for a in arrs])
        stride = max([len(d) for d in drafts] + [1])
        dr = np.zeros((n, stride), np.int32)
        dl = np.zeros(n, np.int32)
        for i, d in enumerate(drafts):
            dr[i, :len(d)] = d
            dl[i] = len(d)
        toks = np.zeros((n, max_tokens), np.int32)
        cnt = np.zeros(n, np.int32)
        launches = ctypes.c_int32(0)
        i32p = ctypes.POINTER(ctypes.c_int32)
        rc = self._lib.moonshine_b200_decode_with_drafts(
            self._handle, ptrs, lens, n, dr.ctypes.data_as(i32p), stride, dl.ctypes.data_as(i32p),
            toks.ctypes.data_as(i32p), max_tokens, cnt.ctypes.data_as(i32p), ctypes.byref(launches))
        _check(rc, "decode_with_drafts failed")
        return [toks[i, :min(cnt[i], max_tokens)].tolist() for i in range(n)], int(launches.value)


    def decode_tokens(self, audios: Sequence, tokens: np.ndarray, vocab: int, n_steps: int, rows_per_launch: int = 8):
        """Teacher-forced multi-token decoder runs (reference: decode_tokens): logits [n_steps, B, V]."""
        arrs = [_as_f32(a) for a in audios]
        n = len(arrs)
        ptrs = (ctypes.POINTER(ctypes.c_float) * n)(*[a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) for a in arrs])
        lens = (ctypes.c_uint64 * n)(*[a.size for a in arrs])
        tokens = np.ascontiguousarray(tokens, np.int32)
        lg = np.zeros((n_steps, n, vocab), np.float32)
        rc = self._lib.moonshine_b200_decode_tokens(
            self._handle, ptrs, lens, n, tokens.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), tokens.shape[1], n_steps,
            rows_per_launch, lg.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        _check(rc, "decode_tokens failed")
        return lg


class Stream:
    def __init__(self, transcriber: Transcriber, handle: int):
        self._t = transcriber
        self._handle = handle

    def start(self):
        _check(self._t._lib.moonshine_start_stream(self._t._handle, self._handle), "Failed to start stream")

    def stop(self):
        _check(self._t._lib.moonshine_stop_stream(self._t._handle, self._handle), "Failed to stop stream")

    def add_audio(self, audio_data, sample_rate: int = 16000):
        a = _as_f32(audio_data)
        rc = self._t._lib.moonshine_transcribe_add_audio_to_stream(
            self._t._handle, self._handle, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), a.size, sample_rate, 0)
        _check(rc, "Failed to add audio to stream")

    def update_transcription(self, flags: int = 0) -> Transcript:
        out = ctypes.POINTER(TranscriptC)()
        rc = self._t._lib.moonshine_transcribe_stream(self._t._handle, self._handle, flags, ctypes.byref(out))
        _check(rc, "Failed to transcribe stream")
        return _parse_transcript(out.contents)

    def close(self):
        if self._handle >= 0:
            self._t._lib.moonshine_free_stream(self._t._handle, self._handle)
            self._handle = -1


def write_model_dir(path: str, arch, weights: dict, tokenizer_bytes: Optional[bytes] = None) -> str:
    """Writes a loadable model directory: model.msw + tokenizer.bin."""
    from .arch import dims_for_arch
    from .weights import synth_tokenizer_bin, write_msw
    os.makedirs(path, exist_ok=True)
    write_msw(os.path.join(path, "model.msw"), arch, weights)
    if tokenizer_bytes is None:
        tokenizer_bytes = synth_tokenizer_bin(dims_for_arch(arch).vocab)
    with open(os.path.join(path, "tokenizer.bin"), "wb") as f:
        f.write(tokenizer_bytes)
    return path
