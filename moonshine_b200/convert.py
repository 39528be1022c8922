"""Real-weight ingestion: Hugging Face `model.safetensors`
checkpoints of `MoonshineForConditionalGeneration` (moonshine-ai/moonshine-tiny | -base)
-> the `.msw` container libmoonshine.so loads.

    python -m moonshine_b200.convert /path/to/hf_checkpoint_dir /path/to/model_dir --arch tiny

The safetensors format is parsed directly (8-byte little-endian header length, JSON
header {name: {dtype, shape, data_offsets}}, raw little-endian tensor bytes), so no
extra dependency is needed.  `tokenizer.bin` (the reference's own format) is copied
alongside if given.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import struct
from typing import Dict

import numpy as np

from .arch import ARCHS, ModelDims, dims_for_arch
from .weights import STREAMING_CONFIG_NAME, streaming_config_record, tensor_specs, write_msw

_DTYPES = {"F32": np.float32, "F16": np.float16, "F64": np.float64}


def read_safetensors(path: str) -> Dict[str, np.ndarray]:
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(n).decode("utf-8"))
        base = 8 + n
        out = {}
        for name, info in header.items():
            if name == "__metadata__":
                continue
            a, b = info["data_offsets"]
            f.seek(base + a)
            raw = f.read(b - a)
            dt = info["dtype"]
            if dt == "BF16":  # upper 16 bits of an fp32
                u16 = np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16
                arr = u16.view(np.float32)
            elif dt in _DTYPES:
                arr = np.frombuffer(raw, dtype=np.dtype(_DTYPES[dt]).newbyteorder("<")).astype(np.float32)
            else:
                raise ValueError(f"{name}: unsupported safetensors dtype {dt}")
            out[name] = arr.reshape(info["shape"]).astype(np.float32)
    return out


def write_safetensors(path: str, tensors: Dict[str, np.ndarray]) -> None:
    """Minimal writer (fp32) -- used by the tests to fabricate a checkpoint."""
    header, blobs, off = {}, [], 0
    for name, a in tensors.items():
        raw = np.ascontiguousarray(a, dtype="<f4").tobytes()
        header[name] = {"dtype": "F32", "shape": list(a.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for r in blobs:
            f.write(r)


def hf_to_msw_tensors(arch, tensors: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Selects / validates the tensors this runtime needs.  `proj_out.weight` is tied to the
    embedding (tie_word_embeddings=True) and rotary buffers are recomputed, so both are dropped."""
    d = dims_for_arch(arch)
    if d.streaming:
        return _hf_streaming_to_msw_tensors(d, tensors)
    out = {}
    for name, shape, _ in tensor_specs(d):
        src = name
        if src not in tensors and name == "model.decoder.embed_tokens.weight" and "proj_out.weight" in tensors:
            src = "proj_out.weight"
        if src not in tensors:
            raise KeyError(f"checkpoint is missing '{name}' (wrong --arch?)")
        a = tensors[src]
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"'{name}' has shape {tuple(a.shape)}, expected {tuple(shape)} (wrong --arch?)")
        out[name] = a
    if "proj_out.weight" in tensors and "model.decoder.embed_tokens.weight" in tensors:
        if not np.array_equal(tensors["proj_out.weight"], tensors["model.decoder.embed_tokens.weight"]):
            raise ValueError("proj_out.weight is not tied to the embedding; this runtime assumes the tied head")
    return out


def _hf_streaming_to_msw_tensors(d: ModelDims, tensors: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """`MoonshineStreamingForConditionalGeneration` checkpoint -> container tensors.  The scalar
    `comp.log_k` becomes a 1-vector, `proj_out.weight` is kept only when it differs from the embedding
    (lora/export.py:206-211 makes the same comparison), and the dimension record is synthesised."""
    out = {}
    for name, shape, kind in tensor_specs(d):
        if kind == "config":
            out[name] = streaming_config_record(d)
            continue
        if name not in tensors:
            raise KeyError(f"checkpoint is missing '{name}' (wrong --arch / config.json?)")
        a = np.asarray(tensors[name], np.float32)
        if kind == "log_k":
            a = a.reshape(1)
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"'{name}' has shape {tuple(a.shape)}, expected {tuple(shape)} (wrong config.json?)")
        out[name] = a
    return out


_STREAMING_ARCH_IDS = {"tiny_streaming": 2, "base_streaming": 3, "small_streaming": 4, "medium_streaming": 5}


def streaming_dims_from_hf_config(cfg: dict, arch: str, tensors: Dict[str, np.ndarray]) -> ModelDims:
    """ModelDims from an HF `config.json` of model_type moonshine_streaming (the reference reads the same
    numbers from streaming_config.json, core/moonshine-streaming-model.cpp:75-116)."""
    enc = cfg.get("encoder_config") or {}
    heads = int(cfg.get("num_attention_heads", 8))
    dim = int(cfg.get("hidden_size", 320))
    head_dim = int(cfg.get("head_dim") or dim // heads)
    if cfg.get("pad_head_dim_to_multiple_of"):
        raise ValueError("pad_head_dim_to_multiple_of checkpoints are not supported")
    rope = cfg.get("rope_parameters") or {}
    windows = tuple(tuple(int(x) for x in w) for w in enc.get(
        "sliding_windows", [[16, 4], [16, 4], [16, 0], [16, 0], [16, 4], [16, 4]]))
    emb, head = tensors.get("model.decoder.embed_tokens.weight"), tensors.get("proj_out.weight")
    tied = bool(cfg.get("tie_word_embeddings", False)) or head is None or (
        emb is not None and np.array_equal(emb, head))
    arch_id = _STREAMING_ARCH_IDS.get(arch, ARCHS[arch].arch if arch in ARCHS else None)
    if arch_id is None:
        raise ValueError(f"unknown streaming arch '{arch}'")
    return ModelDims(
        arch, arch_id, dim, int(enc.get("num_hidden_layers", 6)), int(cfg.get("num_hidden_layers", 6)), heads,
        head_dim, int(cfg.get("intermediate_size", 1280)), vocab=int(cfg.get("vocab_size", 32768)),
        rope_factor=float(rope.get("partial_rotary_factor", 0.8)), rope_theta=float(rope.get("rope_theta", 10000.0)),
        bos=int(cfg.get("bos_token_id", 1)), eos=int(cfg.get("eos_token_id", 2)), streaming=True,
        enc_dim=int(enc.get("hidden_size", 320)), enc_ffn=int(enc.get("intermediate_size", 1280)),
        windows=windows, tied=tied, max_pos_emb=int(cfg.get("max_position_embeddings", 4096)))


def convert(checkpoint: str, out_dir: str, arch: str, tokenizer: str | None = None) -> str:
    st = checkpoint if checkpoint.endswith(".safetensors") else os.path.join(checkpoint, "model.safetensors")
    raw = read_safetensors(st)
    if arch.endswith("streaming"):
        with open(os.path.join(os.path.dirname(st), "config.json")) as f:
            arch = streaming_dims_from_hf_config(json.load(f), arch, raw)
    tensors = hf_to_msw_tensors(arch, raw)
    os.makedirs(out_dir, exist_ok=True)
    write_msw(os.path.join(out_dir, "model.msw"), arch, tensors)
    if tokenizer:
        shutil.copyfile(tokenizer, os.path.join(out_dir, "tokenizer.bin"))
    return out_dir


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("checkpoint")
    ap.add_argument("out_dir")
    ap.add_argument("--arch", required=True,
                    choices=["tiny", "base", "tiny_streaming", "base_streaming", "small_streaming", "medium_streaming"],
                    help="streaming archs also read config.json next to model.safetensors")
    ap.add_argument("--tokenizer", help="tokenizer.bin to copy next to model.msw")
    a = ap.parse_args()
    print(convert(a.checkpoint, a.out_dir, a.arch, a.tokenizer))


if __name__ == "__main__":
    main()
