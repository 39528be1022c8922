"""Model dimensions per architecture id.

Follows the reference's arch table (core/moonshine-model.cpp:41-79:
TINY = 6 layers / 8 heads / head_dim 36, BASE = 8 / 8 / 52) and the HF
configuration the shipped graphs were exported from
(transformers/models/moonshine/configuration_moonshine.py).
"""
from dataclasses import dataclass

MOONSHINE_MODEL_ARCH_TINY = 0
MOONSHINE_MODEL_ARCH_BASE = 1


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
    # Reduced-size configs used only by unit tests (fast oracle runs).
    "test": ModelDims("test", 100, 64, 2, 2, 4, 16, 96, vocab=512),
    "test2": ModelDims("test2", 101, 72, 2, 3, 2, 36, 80, vocab=300),
}


def dims_for_arch(arch) -> ModelDims:
    if isinstance(arch, str):
        return ARCHS[arch]
    for d in ARCHS.values():
        if d.arch == arch:
            return d
    raise ValueError(f"unknown model arch {arch}")


def frontend_lengths(n_samples: int):
    """Frame counts after conv1(k127,s64) / conv2(k7,s3) / conv3(k3,s2); no
    padding anywhere (HF modeling_moonshine.py:501-509)."""
    t1 = (n_samples - 127) // 64 + 1 if n_samples >= 127 else 0
    t2 = (t1 - 7) // 3 + 1 if t1 >= 7 else 0
    t3 = (t2 - 3) // 2 + 1 if t2 >= 3 else 0
    return t1, t2, t3
