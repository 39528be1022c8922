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

from .arch import dims_for_arch
from .weights import tensor_specs, write_msw

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


def convert(checkpoint: str, out_dir: str, arch: str, tokenizer: str | None = None) -> str:
    st = checkpoint if checkpoint.endswith(".safetensors") else os.path.join(checkpoint, "model.safetensors")
    tensors = hf_to_msw_tensors(arch, read_safetensors(st))
    os.makedirs(out_dir, exist_ok=True)
    write_msw(os.path.join(out_dir, "model.msw"), arch, tensors)
    if tokenizer:
        shutil.copyfile(tokenizer, os.path.join(out_dir, "tokenizer.bin"))
    return out_dir


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("checkpoint")
    ap.add_argument("out_dir")
    ap.add_argument("--arch", choices=["tiny", "base"], required=True)
    ap.add_argument("--tokenizer", help="tokenizer.bin to copy next to model.msw")
    a = ap.parse_args()
    print(convert(a.checkpoint, a.out_dir, a.arch, a.tokenizer))


if __name__ == "__main__":
    main()
