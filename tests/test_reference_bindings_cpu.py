"""CPU: the REFERENCE's own Python binding (language-bindings/python/src/moonshine_voice, pure ctypes), imported
unmodified from /root/reference, driven against THIS repo's libmoonshine.so.

What it pins: the struct-size guard of the binding passes (moonshine_api.py:137-146: 24 / 40 / 88 / 16 bytes),
every symbol the binding declares resolves (moonshine_api.py:864-970 `_setup_function_signatures` touches each
one), and the reference's `Transcriber` class works end to end through the ABI on the segmentation-only path
(`skip_transcription`, no GPU needed): option passing, handles, transcript_t parsing, streams, error codes.
The binding dlopens "libmoonshine.so" by name; this library carries that SONAME, so preloading it by path makes
the binding bind to it (the same thing an installed wheel does by placing the file next to the module).
Skipped where /root/reference does not exist (the GPU box)."""
import ctypes
import os
import sys

import numpy as np
import pytest

REF_PY = "/root/reference/language-bindings/python/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_PY, "moonshine_voice")),
                                reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_binding():
    from moonshine_b200 import api
    path = api.lib_path()
    assert os.path.exists(path)
    ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)          # same SONAME as the name the binding dlopens
    sys.path.insert(0, REF_PY)
    try:
        import moonshine_voice  # noqa: F401  (runs the struct-size guard at import)
        from moonshine_voice import moonshine_api, transcriber
        lib = moonshine_api._MoonshineLib().lib          # resolves every declared symbol
        # the binding must be talking to this repo's library, not to some other libmoonshine
        assert hasattr(lib, "moonshine_b200_transcribe_device")
        yield moonshine_api, transcriber, lib
    finally:
        sys.path.remove(REF_PY)


def test_struct_sizes_and_symbols(ref_binding):
    moonshine_api, _, lib = ref_binding
    assert ctypes.sizeof(moonshine_api.TranscriptWordC) == 24
    assert ctypes.sizeof(moonshine_api.SpeakerSpanC) == 40
    assert ctypes.sizeof(moonshine_api.TranscriptLineC) == 88
    assert ctypes.sizeof(moonshine_api.TranscriptC) == 16
    assert lib.moonshine_get_version() == moonshine_api.MOONSHINE_HEADER_VERSION
    assert lib.moonshine_error_to_string(-2) == b"Invalid handle"


def test_reference_transcriber_class_segments_audio(ref_binding):
    moonshine_api, transcriber, _ = ref_binding
    t = transcriber.Transcriber("/nonexistent-model-dir", moonshine_api.ModelArch.TINY,
                                options={"skip_transcription": "true", "vad_threshold": "0"})
    rng = np.random.default_rng(3)
    audio = (rng.standard_normal(16000 * 21) * 0.05).astype(np.float32)
    tr = t.transcribe_without_streaming(audio.tolist(), 16000)
    # vad_threshold=0 is the documented bypass: everything is voice (core/voice-activity-detector.cpp:152-157);
    # whole hops only, so the line holds the input up to one hop (512 samples)
    assert len(tr.lines) == 1
    assert all(ln.is_complete and ln.is_new and ln.is_updated for ln in tr.lines)
    assert tr.lines[0].start_time < 1e-3
    assert 0 <= len(audio) - len(tr.lines[0].audio_data) < 512
    np.testing.assert_array_equal(np.asarray(tr.lines[0].audio_data, np.float32), audio[:len(tr.lines[0].audio_data)])
    # default threshold: the max-segment fade cuts the clip (two lines, distinct ids)
    t2 = transcriber.Transcriber("/nonexistent-model-dir", moonshine_api.ModelArch.TINY, options={"skip_transcription": "true"})
    tr2 = t2.transcribe_without_streaming(audio.tolist(), 16000)
    assert len(tr2.lines) >= 2 and len({ln.line_id for ln in tr2.lines}) == len(tr2.lines)
    t2.close()
    t.close()


def test_reference_stream_class(ref_binding):
    moonshine_api, transcriber, _ = ref_binding
    t = transcriber.Transcriber("/nonexistent-model-dir", moonshine_api.ModelArch.TINY,
                                options={"skip_transcription": "true", "vad_threshold": "0"})
    s = t.create_stream(update_interval=0.1)
    s.start()
    rng = np.random.default_rng(4)
    total = 0
    for _ in range(5):
        chunk = (rng.standard_normal(16000) * 0.05).astype(np.float32)
        s.add_audio(chunk.tolist(), 16000)
        total += len(chunk)
        tr = s.update_transcription(moonshine_api.MOONSHINE_FLAG_FORCE_UPDATE)
        assert len(tr.lines) == 1 and not tr.lines[0].is_complete
    s.stop()
    tr = s.update_transcription(moonshine_api.MOONSHINE_FLAG_FORCE_UPDATE)
    assert len(tr.lines) == 1 and tr.lines[0].is_complete
    s.close()
    t.close()


def test_reference_binding_error_paths(ref_binding):
    moonshine_api, transcriber, lib = ref_binding
    from moonshine_voice.errors import MoonshineError
    with pytest.raises(MoonshineError):   # unknown option key must fail the load (moonshine-c-api.cpp:193-196)
        transcriber.Transcriber("/nonexistent-model-dir", moonshine_api.ModelArch.TINY,
                                options={"skip_transcription": "true", "no_such_option": "1"})
    out = ctypes.POINTER(moonshine_api.TranscriptC)()
    buf = (ctypes.c_float * 16)()
    assert lib.moonshine_transcribe_without_streaming(12345, buf, 16, 16000, 0, ctypes.byref(out)) == -2
