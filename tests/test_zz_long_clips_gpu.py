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
    check_stream_case("test_streaming", 0, "scaled", [synth_audio(74, 16000 * 15), synth_audio(75, 16000 * 11 + 7)])
