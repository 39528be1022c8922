"""Model dimensions per architecture id.

Follows the reference's arch table (core/moonshine-model.cpp:41-79:
TINY = 6 layers / 8 heads / head_dim 36, BASE = 8 / 8 / 52) and the HF
configuration the shipped graphs were exported from
(transformers/models/moonshine/configuration_moonshine.py).
"""
from dataclasses import dataclass
from typing import Optional, Tuple

MOONSHINE_MODEL_ARCH_TINY = 0
MOONSHINE_MODEL_ARCH_BASE = 1
# streaming family (core/moonshine-c-api.h:109-112); their dimensions are data, read from the
# model directory by the reference (streaming_config.json, core/moonshine-streaming-model.cpp:75-116)
# and from the weight container's "streaming.config" record here
MOONSHINE_MODEL_ARCH_TINY_STREAMING = 2
MOONSHINE_MODEL_ARCH_BASE_STREAMING = 3
MOONSHINE_MODEL_ARCH_SMALL_STREAMING = 4
MOONSHINE_MODEL_ARCH_MEDIUM_STREAMING = 5
STREAMING_ARCHS = (2, 3, 4, 5)

STREAM_CHUNK = 1280      # samples per frontend call (core/transcriber.cpp:1342)
STREAM_FRAME = 80        # samples per frame (5 ms)
STREAM_FEATURE = 320     # samples per encoder feature (two stride-2 causal convs)


@dataclass(frozen=True)
class ModelDims:
    name: str
    arch: int
    dim: int          # D, encoder == decoder hidden size
    enc_layers: int
    dec_layers: int
    heads: int
    head_dim: int
    ffn: int          # I (decoder fc1 is D -> 2I, gated)
    vocab: int = 32768
    rope_factor: float = 0.9
    rope_theta: float = 10000.0
    bos: int = 1
    eos: int = 2
    # --- streaming family only (HF MoonshineStreamingConfig / lora/export.py) ---
    streaming: bool = False
    enc_dim: int = 0                 # encoder hidden size (adapter projects to `dim` when different)
    enc_ffn: int = 0
    windows: Optional[Tuple[Tuple[int, int], ...]] = None   # per encoder layer (past, future), inclusive
    tied: bool = True                # logits read off embed_tokens (True) or proj_out (False)
    max_seq_len: int = 448           # decoder position limit (streaming_config.json "max_seq_len")
    max_pos_emb: int = 4096          # adapter position table rows

    @property
    def enc_head_dim(self) -> int:
        return (self.enc_dim or self.dim) // self.heads

    @property
    def total_lookahead(self) -> int:
        return sum(f for _, f in (self.windows or ()))

    @property
    def rot_dim(self) -> int:
        # HF: dim = int(head_dim * partial_rotary_factor); arange(0, dim, 2)
        # gives ceil(dim/2) frequencies, each rotating one interleaved pair.
        d = int(self.head_dim * self.rope_factor)
        return 2 * ((d + 1) // 2)

    @property
    def rope_denominator(self) -> int:
        return int(self.head_dim * self.rope_factor)


ARCHS = {
    "tiny": ModelDims("tiny", MOONSHINE_MODEL_ARCH_TINY, 288, 6, 6, 8, 36, 1152),
    "base": ModelDims("base", MOONSHINE_MODEL_ARCH_BASE, 416, 8, 8, 8, 52, 1664),
    # HF MoonshineStreamingConfig() defaults == UsefulSensors/moonshine-streaming-tiny
    "tiny_streaming": ModelDims("tiny_streaming", MOONSHINE_MODEL_ARCH_TINY_STREAMING, 320, 6, 6, 8, 40, 1280,
                                rope_factor=0.8, streaming=True, enc_dim=320, enc_ffn=1280,
                                windows=((16, 4), (16, 4), (16, 0), (16, 0), (16, 4), (16, 4)), tied=False),
    # synthetic "base-streaming" of BASELINE config #4 (core/moonshine-streaming-model.cpp:35-39 names only the
    # decoder side; encoder depth 8, lookahead kept at 16 frames)
    "base_streaming": ModelDims("base_streaming", MOONSHINE_MODEL_ARCH_BASE_STREAMING, 416, 8, 8, 8, 52, 1664,
                                rope_factor=0.8, streaming=True, enc_dim=416, enc_ffn=1664,
                                windows=((16, 4), (16, 4), (16, 0), (16, 0), (16, 0), (16, 0), (16, 4), (16, 4)),
                                tied=False),
    # Reduced-size configs used only by unit tests (fast oracle runs).
    "test_streaming": ModelDims("test_streaming", 102, 64, 3, 2, 4, 16, 96, vocab=512, rope_factor=0.8,
                                streaming=True, enc_dim=96, enc_ffn=128, windows=((4, 2), (4, 0), (4, 2)),
                                tied=False, max_pos_emb=512),
    "test_streaming2": ModelDims("test_streaming2", 103, 96, 2, 2, 2, 48, 64, vocab=300, rope_factor=0.8,
                                 streaming=True, enc_dim=96, enc_ffn=160, windows=((3, 1), (5, 3)),
                                 tied=True, max_pos_emb=512),
    "test": ModelDims("test", 100, 64, 2, 2, 4, 16, 96, vocab=512),
    "test2": ModelDims("test2", 101, 72, 2, 3, 2, 36, 80, vocab=300),
}


def dims_for_arch(arch) -> ModelDims:
    if isinstance(arch, ModelDims):
        return arch
    if isinstance(arch, str):
        return ARCHS[arch]
    for d in ARCHS.values():
        if d.arch == arch:
            return d
    raise ValueError(f"unknown model arch {arch}")


def streaming_lengths(n_samples: int, emitted_before: int = 0, processed_before: int = 0,
                      is_final: bool = True, lookahead: int = 16):
    """Host bookkeeping of Transcriber::transcribe_segment_with_streaming_model
    (core/transcriber.cpp:1331-1372) + MoonshineStreamingModel::encode (:604-640) for one segment.

    Only whole 1280-sample chunks are analysed; a non-final update holds back `lookahead`
    features; when no new chunk arrived the encoder is not run at all (so a final call on an
    already fully analysed segment keeps the held-back features out of the memory).
    Returns (processed_samples, n_features, emitted_features)."""
    processed = processed_before
    emitted = emitted_before
    if processed < n_samples:
        processed += (n_samples - processed) // STREAM_CHUNK * STREAM_CHUNK
        n = processed // STREAM_FEATURE
        stable = n if is_final else max(0, n - lookahead)
        if n > 0 and stable > emitted:
            emitted = stable
    return processed, processed // STREAM_FEATURE, emitted


def frontend_lengths(n_samples: int):
    """Frame counts after conv1(k127,s64) / conv2(k7,s3) / conv3(k3,s2); no
    padding anywhere (HF modeling_moonshine.py:501-509)."""
    t1 = (n_samples - 127) // 64 + 1 if n_samples >= 127 else 0
    t2 = (t1 - 7) // 3 + 1 if t1 >= 7 else 0
    t3 = (t2 - 3) // 2 + 1 if t2 >= 3 else 0
    return t1, t2, t3
