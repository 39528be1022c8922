"""GPU: clips longer than 10.8 s (more than 448 encoder frames).  On the classic encoder these take the unfused
attention path (scores through HBM) and the decoder's full-CTA cross-attention with multi-chunk K/V; on the
streaming encoder the banded fused kernel with several query tiles.  15 s is the segmenter's maximum.
(Runs last: the file name sorts after the other GPU tests.)"""
import pytest

from moonshine_b200.weights import synth_audio
from tests.test_parity_gpu import check_case
from tests.test_streaming_gpu import check_stream_case

pytestmark = pytest.mark.gpu


def test_small_arch_long_and_ragged():
    audios = [synth_audio(70, 16000 * 14 + 321), synth_audio(71, 16000 * 12), synth_audio(72, 30000)]
    check_case("test", 0, "scaled", audios)


def test_tiny_thirteen_seconds():
    check_case("tiny", 0, "scaled", [synth_audio(73, 16000 * 13)])


def test_streaming_fifteen_seconds():
    """750 features: needs the real archs' 4096-row adapter position table (the toy streaming archs stop at 512)."""
    check_stream_case("tiny_streaming", 0, "scaled", [synth_audio(74, 16000 * 15), synth_audio(75, 16000 * 11 + 7)])


def test_small_streaming_multi_tile_band():
    """500 / 470 features: more than one 128-query tile and more than 448 keys, inside the toy arch's 512-row table."""
    check_stream_case("test_streaming", 0, "scaled", [synth_audio(76, 16000 * 10), synth_audio(77, 320 * 470 + 11)])


def test_streaming_segment_beyond_position_table_is_a_clean_error():
    """A segment longer than the adapter position table must fail with an error code, not crash."""
    from tests.test_streaming_gpu import make_transcriber
    from moonshine_b200.arch import ARCHS
    d = ARCHS["test_streaming"]
    t = make_transcriber("test_streaming")
    with pytest.raises(Exception):
        t.debug_run([synth_audio(78, 16000 * 15)], d.dim, d.vocab, max_tokens=300)
    t.close()
