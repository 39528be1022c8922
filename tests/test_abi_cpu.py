"""CPU-only: the C-ABI library loads, exports every declared symbol, keeps the
reference's struct layouts and error behaviour that needs no device."""
import ctypes
import re
import os

import numpy as np
import pytest

from moonshine_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from moonshine_b200.build import build
    build(verbose=False)
    return api.load_library()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "moonshine_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(moonshine_[a-z0-9_]+)\s*\(", header))
    declared -= {"moonshine_option_t", "moonshine_speech_clip_t"}
    assert declared == set(api.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_sizes_match_reference_abi():
    # language-bindings/python/src/moonshine_voice/moonshine_api.py:137-146
    assert ctypes.sizeof(api.TranscriptWordC) == 24
    assert ctypes.sizeof(api.SpeakerSpanC) == 40
    assert ctypes.sizeof(api.TranscriptLineC) == 88
    assert ctypes.sizeof(api.TranscriptC) == 16


def test_version_and_error_strings(lib):
    assert lib.moonshine_get_version() == 30000
    assert lib.moonshine_error_to_string(0) == b"Success"
    assert lib.moonshine_error_to_string(-2) == b"Invalid handle"
    assert lib.moonshine_error_to_string(-3) == b"Invalid argument"
    assert lib.moonshine_error_to_string(-1) == b"Unknown error"
    assert lib.moonshine_error_to_string(-77) == b"Unknown error"


def test_invalid_handle_codes(lib):
    out = ctypes.POINTER(api.TranscriptC)()
    a = np.zeros(16, np.float32)
    rc = lib.moonshine_transcribe_without_streaming(12345, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 16, 16000, 0, ctypes.byref(out))
    assert rc == -2
    assert lib.moonshine_create_stream(-1, 0) == -2
    assert lib.moonshine_start_stream(999, 1) == -2
    lib.moonshine_free_transcriber(424242)  # no-op, must not crash


def test_unknown_option_fails_load(lib):
    # core/moonshine-c-api-test.cpp:538: invalid option => negative handle
    with pytest.raises(api.MoonshineError):
        api.Transcriber("/nonexistent", api.ModelArch.TINY, {"definitely_not_an_option": "1"})


def test_from_memory_refused_for_new_headers(lib):
    rc = lib.moonshine_load_transcriber_from_memory(None, 0, None, 0, None, 0, None, 0, 0, None, 0, 30000)
    assert rc == -3


def test_memory_files_rejects_unknown_key(lib):
    with pytest.raises(api.MoonshineError) as e:
        api.Transcriber(model_arch=api.ModelArch.TINY, memory_files={"not_a_model_file.bin": b"x"})
    assert e.value.code == -3


def test_stub_exports_report_failure(lib):
    lib.moonshine_get_stt_catalog.restype = ctypes.c_int32
    p = ctypes.c_char_p()
    assert lib.moonshine_get_stt_catalog(ctypes.byref(p)) == -1


def test_skip_transcription_segments_without_a_device(lib):
    """ModelSource::NONE path of the reference (moonshine-c-api-test.cpp:328-370):
    VAD + line plumbing only, text == NULL.  vad_threshold=0 => exactly one line
    spanning the clip (hop-truncated: 512-sample hops, remainder dropped)."""
    t = api.Transcriber(None, api.ModelArch.TINY, {"skip_transcription": "true", "vad_threshold": "0"})
    audio = (np.random.default_rng(0).standard_normal(16000 * 3 + 100) * 0.05).astype(np.float32)
    tr = t.transcribe_without_streaming(audio)
    assert len(tr.lines) == 1
    l = tr.lines[0]
    assert l.text is None and l.is_complete and l.is_new and l.is_updated
    assert l.start_time < 1e-3
    assert l.audio_data.size == (audio.size // 512) * 512
    assert abs(int(l.audio_data.size) - audio.size) <= 512
    np.testing.assert_array_equal(l.audio_data, audio[: l.audio_data.size])
    # resampled input: 48 kHz -> 16 kHz box average
    tr2 = t.transcribe_without_streaming(np.repeat(audio, 3), sample_rate=48000)
    assert len(tr2.lines) == 1
    # ids are unique across calls
    assert tr2.lines[0].line_id != l.line_id
    t.close()


def test_skip_transcription_default_vad_splits_long_audio(lib):
    """With the Silero model out of scope every hop counts as speech, so the
    reference's max-segment fade rule (voice-activity-detector.cpp:161-169)
    cuts a long clip at ~2/3 of vad_max_segment_duration."""
    t = api.Transcriber(None, api.ModelArch.TINY, {"skip_transcription": "true"})
    audio = np.zeros(16000 * 25, np.float32)
    tr = t.transcribe_without_streaming(audio)
    assert len(tr.lines) >= 2
    assert all(l.is_complete for l in tr.lines)
    assert max(l.duration for l in tr.lines) <= 15.0 + 1e-3
    ids = [l.line_id for l in tr.lines]
    assert len(set(ids)) == len(ids)
    t.close()


def test_default_vad_look_behind_content(lib):
    """Default threshold (0.5): the 16-hop smoothing window opens the first segment at hop 9
    and the 8192-sample look-behind reaches back to sample 0; a later segment starts with the
    look-behind taken across the cut (voice-activity-detector.cpp:171-190)."""
    t = api.Transcriber(None, api.ModelArch.TINY, {"skip_transcription": "true"})
    rng = np.random.default_rng(7)
    audio = (rng.standard_normal(16000 * 23) * 0.05).astype(np.float32)
    tr = t.transcribe_without_streaming(audio)
    first = tr.lines[0]
    assert first.start_time < 1e-3
    np.testing.assert_array_equal(first.audio_data, audio[: first.audio_data.size])
    # every line is a contiguous slice of the input that ends on a hop boundary
    for l in tr.lines:
        n = l.audio_data.size
        end = int(round((l.start_time + l.duration) * 16000))
        assert end % 512 == 0
        np.testing.assert_array_equal(l.audio_data, audio[end - n: end])
    t.close()


def test_streaming_line_invariants_without_a_device(lib):
    """core/transcriber-test.cpp:199-403 style invariants on the stream API."""
    t = api.Transcriber(None, api.ModelArch.TINY, {"skip_transcription": "true", "vad_threshold": "0"})
    s = t.create_stream()
    s.start()
    rng = np.random.default_rng(1)
    fed = np.zeros(0, np.float32)
    for i in range(6):
        chunk = (rng.standard_normal(4000) * 0.05).astype(np.float32)
        fed = np.concatenate([fed, chunk])
        s.add_audio(chunk)
        tr = s.update_transcription(api.MOONSHINE_FLAG_FORCE_UPDATE)
        assert len(tr.lines) == 1
        # hops that straddle two add_audio calls are assembled: the line holds every whole hop so far
        np.testing.assert_array_equal(tr.lines[0].audio_data, fed[: fed.size // 512 * 512])
        assert not tr.lines[0].is_complete  # only the last line may be incomplete
        assert tr.lines[0].is_new == (i == 0)
    # no new audio, no force: cached transcript, flags cleared
    tr = s.update_transcription()
    assert len(tr.lines) == 1 and not tr.lines[0].is_updated
    s.stop()
    tr = s.update_transcription()
    assert tr.lines[0].is_complete
    s.close()
    t.close()
