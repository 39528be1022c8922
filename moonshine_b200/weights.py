"""Weight tooling: canonical tensor list, seeded synthetic weights, and the
``.msw`` container the CUDA library loads.

Tensor names are the HF ``MoonshineForConditionalGeneration`` state-dict keys
(transformers/models/moonshine/modeling_moonshine.py) so a real checkpoint maps
1:1.  The reference itself ships no weights in-tree (they are downloaded,
scripts/fetch-voice-assets.sh:85-112), so every test / bench weight here is
synthetic and seeded.

``.msw`` layout (little endian):
    magic  "MSW1"            4 bytes
    u32    version (=1)
    u32    arch id
    u32    tensor count
    per tensor: u16 name_len, name, u8 dtype(0=f32), u8 ndim, u32 dims[ndim],
                u64 data_offset (from file start, 256-byte aligned), u64 nbytes
    data region
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

from .arch import ModelDims, dims_for_arch

MSW_MAGIC = b"MSW1"


def tensor_specs(d: ModelDims) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) in canonical order.  kind in
    {weight, bias, norm_w, norm_b, embed}; fan_in is shape[1:] product."""
    if d.streaming:
        return streaming_tensor_specs(d)
    D, I, V = d.dim, d.ffn, d.vocab
    specs: List[Tuple[str, Tuple[int, ...], str]] = []
    e = "model.encoder."
    specs += [
        (e + "conv1.weight", (D, 1, 127), "weight"),
        (e + "conv2.weight", (2 * D, D, 7), "weight"),
        (e + "conv2.bias", (2 * D,), "bias"),
        (e + "conv3.weight", (D, 2 * D, 3), "weight"),
        (e + "conv3.bias", (D,), "bias"),
        (e + "groupnorm.weight", (D,), "norm_w"),
        (e + "groupnorm.bias", (D,), "norm_b"),
    ]
    for l in range(d.enc_layers):
        p = f"{e}layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            specs.append((p + f"self_attn.{n}.weight", (D, D), "weight"))
        specs += [
            (p + "mlp.fc1.weight", (I, D), "weight"),
            (p + "mlp.fc1.bias", (I,), "bias"),
            (p + "mlp.fc2.weight", (D, I), "weight"),
            (p + "mlp.fc2.bias", (D,), "bias"),
            (p + "input_layernorm.weight", (D,), "norm_w"),
            (p + "post_attention_layernorm.weight", (D,), "norm_w"),
        ]
    specs.append((e + "layer_norm.weight", (D,), "norm_w"))
    dd = "model.decoder."
    specs.append((dd + "embed_tokens.weight", (V, D), "embed"))
    for l in range(d.dec_layers):
        p = f"{dd}layers.{l}."
        for a in ("self_attn", "encoder_attn"):
            for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
                specs.append((p + f"{a}.{n}.weight", (D, D), "weight"))
        specs += [
            (p + "mlp.fc1.weight", (2 * I, D), "weight"),
            (p + "mlp.fc1.bias", (2 * I,), "bias"),
            (p + "mlp.fc2.weight", (D, I), "weight"),
            (p + "mlp.fc2.bias", (D,), "bias"),
            (p + "input_layernorm.weight", (D,), "norm_w"),
            (p + "post_attention_layernorm.weight", (D,), "norm_w"),
            (p + "final_layernorm.weight", (D,), "norm_w"),
        ]
    specs.append((dd + "norm.weight", (D,), "norm_w"))
    return specs


STREAMING_CONFIG_NAME = "streaming.config"


def streaming_config_record(d: ModelDims) -> np.ndarray:
    """The dimensions the reference reads from streaming_config.json
    (core/moonshine-streaming-model.cpp:75-116) plus the encoder-side ones its graphs bake in
    (lora/export.py:100-127 windows), as one float32 vector stored beside the tensors."""
    w = [x for pf in d.windows for x in pf]
    return np.asarray([1, d.enc_dim, d.dim, d.enc_layers, d.dec_layers, d.heads, d.head_dim, d.enc_ffn, d.ffn,
                       d.vocab, d.rope_denominator, d.rot_dim, d.rope_theta, 1.0 if d.tied else 0.0,
                       d.max_seq_len, d.max_pos_emb, d.bos, d.eos, len(d.windows)] + w, np.float32)


def streaming_tensor_specs(d: ModelDims) -> List[Tuple[str, Tuple[int, ...], str]]:
    """HF ``MoonshineStreamingForConditionalGeneration`` state-dict keys
    (transformers/models/moonshine_streaming/modeling_moonshine_streaming.py)."""
    E, EI, D, I, V = d.enc_dim, d.enc_ffn, d.dim, d.ffn, d.vocab
    specs: List[Tuple[str, Tuple[int, ...], str]] = [
        (STREAMING_CONFIG_NAME, (19 + 2 * len(d.windows),), "config")]
    e = "model.encoder."
    specs += [
        (e + "embedder.comp.log_k", (1,), "log_k"),
        (e + "embedder.linear.weight", (E, 80), "weight"),
        (e + "embedder.conv1.weight", (2 * E, E, 5), "weight"),
        (e + "embedder.conv1.bias", (2 * E,), "bias"),
        (e + "embedder.conv2.weight", (E, 2 * E, 5), "weight"),
        (e + "embedder.conv2.bias", (E,), "bias"),
    ]
    for l in range(d.enc_layers):
        p = f"{e}layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            specs.append((p + f"self_attn.{n}.weight", (E, E), "weight"))
        specs += [
            (p + "mlp.fc1.weight", (EI, E), "weight"),
            (p + "mlp.fc1.bias", (EI,), "bias"),
            (p + "mlp.fc2.weight", (E, EI), "weight"),
            (p + "mlp.fc2.bias", (E,), "bias"),
            (p + "input_layernorm.gamma", (E,), "gamma0"),
            (p + "post_attention_layernorm.gamma", (E,), "gamma0"),
        ]
    specs.append((e + "final_norm.gamma", (E,), "gamma0"))
    dd = "model.decoder."
    specs.append((dd + "embed_tokens.weight", (V, D), "embed"))
    specs.append((dd + "pos_emb.weight", (d.max_pos_emb, E), "weight02"))
    if E != D:
        specs.append((dd + "proj.weight", (D, E), "weight"))
    for l in range(d.dec_layers):
        p = f"{dd}layers.{l}."
        for a in ("self_attn", "encoder_attn"):
            for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
                specs.append((p + f"{a}.{n}.weight", (D, D), "weight"))
        specs += [
            (p + "mlp.fc1.weight", (2 * I, D), "weight"),
            (p + "mlp.fc1.bias", (2 * I,), "bias"),
            (p + "mlp.fc2.weight", (D, I), "weight"),
            (p + "mlp.fc2.bias", (D,), "bias"),
            (p + "input_layernorm.weight", (D,), "norm_w"),
            (p + "post_attention_layernorm.weight", (D,), "norm_w"),
            (p + "final_layernorm.weight", (D,), "norm_w"),
        ]
    specs.append((dd + "norm.weight", (D,), "norm_w"))
    if not d.tied:
        specs.append(("proj_out.weight", (V, D), "embed"))
    return specs


def synth_weights(arch, seed: int = 0, init: str = "scaled") -> Dict[str, np.ndarray]:
    """Deterministic synthetic weights.

    init="hf":     HF ``_init_weights``: N(0, 0.02) weights, zero biases, unit
                   norms (what ``MoonshineForConditionalGeneration(cfg)``
                   gives) -- used for throughput runs.
    init="scaled": N(0, 1/sqrt(fan_in)) weights, small random biases, norms
                   1 + 0.1*N(0,1) -- keeps activations O(1) through the stack
                   so attention is far from uniform and every bias / gamma
                   path is exercised; used for parity tests.
    Each tensor draws from its own ``default_rng([seed, index])`` so values do
    not depend on the order tensors are generated in.
    """
    d = dims_for_arch(arch)
    out: Dict[str, np.ndarray] = {}
    for idx, (name, shape, kind) in enumerate(tensor_specs(d)):
        rng = np.random.default_rng([seed, idx])
        if kind == "config":
            out[name] = streaming_config_record(d)
            continue
        if kind == "log_k":   # HF init: log(0.75)
            out[name] = np.asarray([np.log(0.75) + (0.0 if init == "hf" else 0.3)], np.float32)
            continue
        if kind == "gamma0":  # unit-offset norm: effective scale is gamma + 1, HF init 0
            out[name] = (np.zeros(shape, np.float32) if init == "hf"
                         else (0.1 * rng.standard_normal(shape, dtype=np.float32)))
            continue
        if kind == "weight02":  # position table: same scale in both inits
            out[name] = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02 if init == "hf" else 0.1)
            continue
        if kind in ("weight", "embed"):
            if init == "hf":
                std = 0.02
            else:
                fan_in = int(np.prod(shape[1:]))
                std = 1.0 / np.sqrt(fan_in)
                if kind == "embed":
                    std = 0.1  # small vs. the layer outputs, else the tied head
                               # just re-predicts the input token
                if name.endswith("conv1.weight"):
                    std = 4.0 / np.sqrt(fan_in)  # audio RMS is ~0.05
            w = rng.standard_normal(shape, dtype=np.float32) * np.float32(std)
        elif kind == "bias":
            w = (np.zeros(shape, np.float32) if init == "hf"
                 else rng.standard_normal(shape, dtype=np.float32) * np.float32(0.05))
        elif kind == "norm_w":
            w = (np.ones(shape, np.float32) if init == "hf"
                 else 1.0 + 0.1 * rng.standard_normal(shape, dtype=np.float32))
        elif kind == "norm_b":
            w = (np.zeros(shape, np.float32) if init == "hf"
                 else 0.05 * rng.standard_normal(shape, dtype=np.float32))
        else:
            raise AssertionError(kind)
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


def pack_msw(arch, weights: Dict[str, np.ndarray]) -> bytes:
    d = dims_for_arch(arch)
    specs = tensor_specs(d)
    table = bytearray()
    entries = []
    for name, shape, _ in specs:
        w = weights[name]
        if tuple(w.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {w.shape} != {shape}")
        entries.append((name.encode(), np.ascontiguousarray(w, np.float32)))
    # first pass: table size
    tsize = 16
    for nb, w in entries:
        tsize += 2 + len(nb) + 2 + 4 * w.ndim + 16
    off = (tsize + 255) // 256 * 256
    blobs = []
    table += MSW_MAGIC + struct.pack("<III", 1, d.arch, len(entries))
    for nb, w in entries:
        table += struct.pack("<H", len(nb)) + nb + struct.pack("<BB", 0, w.ndim)
        table += struct.pack(f"<{w.ndim}I", *w.shape)
        table += struct.pack("<QQ", off, w.nbytes)
        blobs.append((off, w))
        off = (off + w.nbytes + 255) // 256 * 256
    buf = bytearray(off)
    buf[: len(table)] = table
    for o, w in blobs:
        buf[o: o + w.nbytes] = w.tobytes()
    return bytes(buf)


def write_msw(path: str, arch, weights: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(pack_msw(arch, weights))


def read_msw(path_or_bytes) -> Tuple[int, Dict[str, np.ndarray]]:
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        buf = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            buf = f.read()
    if buf[:4] != MSW_MAGIC:
        raise ValueError("not an MSW1 container")
    ver, arch, n = struct.unpack_from("<III", buf, 4)
    p = 16
    out = {}
    for _ in range(n):
        (ln,) = struct.unpack_from("<H", buf, p); p += 2
        name = buf[p: p + ln].decode(); p += ln
        dt, nd = struct.unpack_from("<BB", buf, p); p += 2
        dims = struct.unpack_from(f"<{nd}I", buf, p); p += 4 * nd
        off, nb = struct.unpack_from("<QQ", buf, p); p += 16
        out[name] = np.frombuffer(buf, np.float32, nb // 4, off).reshape(dims).copy()
    return arch, out


def synth_tokenizer_bin(vocab: int) -> bytes:
    """A synthetic ``tokenizer.bin`` in the reference's record format
    (core/bin-tokenizer/bin-tokenizer.cpp:46-66): id 0/1/2 = <unk>/<s>/</s>,
    then printable pieces.  Only used when no real tokenizer.bin is at hand."""
    recs = [b"<unk>", b"<s>", b"</s>"]
    i = 0
    while len(recs) < vocab:
        piece = ("▁" if i % 3 == 0 else "") + _b26(i)
        recs.append(piece.encode("utf-8"))
        i += 1
    out = bytearray()
    for r in recs:
        n = len(r)
        if n == 0:
            out.append(0)
        elif n < 128:
            out.append(n)
        else:
            out.append(128 + n % 128); out.append(n // 128)
        out += r
    return bytes(out)


def _b26(i: int) -> str:
    s = ""
    i += 1
    while i > 0:
        i, r = divmod(i - 1, 26)
        s = chr(ord("a") + r) + s
    return s


def synth_audio(i: int, n_samples: int = 160000) -> np.ndarray:
    """Synthetic utterance i: Gaussian noise, 3-tap
    smooth, RMS 0.05, clipped to [-1, 1]."""
    rng = np.random.default_rng(1234 + i)
    x = rng.standard_normal(n_samples).astype(np.float32)
    y = x.copy()
    y[1:-1] = (x[:-2] + x[1:-1] + x[2:]) / np.float32(3.0)
    rms = np.sqrt(np.mean(y.astype(np.float64) ** 2))
    y = (y * np.float32(0.05 / rms)).astype(np.float32)
    return np.clip(y, -1.0, 1.0)
