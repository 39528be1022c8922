#!/usr/bin/env python
"""Golden fixtures for the streaming architectures, from the Hugging Face modules themselves.

Run in the BUILD container only (needs transformers):
    python tests/golden/make_golden_streaming.py

The reference's streaming graphs are produced by its in-tree exporter
(language-bindings/python/src/moonshine_voice/lora/export.py) from HF
``MoonshineStreamingForConditionalGeneration`` modules; the exporter changes ONE thing relative to a
plain HF forward: each encoder layer gets the inclusive sliding-window mask
``-future <= q - k <= past`` (export.py:110-125) -- HF itself builds no mask without an attention mask.
This script therefore calls the HF submodules (embedder, encoder layers with that additive mask,
final norm, decoder with its internal pos_emb + proj adapter, proj_out) with this repo's seeded
synthetic weights, on the samples the reference would analyse (whole 1280-sample chunks).

Writes tests/golden/hfs_<arch>_<init>_s<seed>_<input>_<final|partial>.npz:
    mem_sub (every 4th memory row = adapter output), mem_shape, tokens (greedy ids incl. BOS),
    logits_sub / top_idx / top_val / margin for the decoded steps, n_samples, emitted.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from moonshine_b200.arch import ARCHS, streaming_lengths  # noqa: E402
from moonshine_b200.weights import STREAMING_CONFIG_NAME, synth_audio, synth_weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def hf_model(d, weights):
    from transformers import MoonshineStreamingConfig, MoonshineStreamingForConditionalGeneration
    from transformers.models.moonshine_streaming.configuration_moonshine_streaming import (
        MoonshineStreamingEncoderConfig)
    enc = MoonshineStreamingEncoderConfig(
        hidden_size=d.enc_dim, intermediate_size=d.enc_ffn, num_hidden_layers=d.enc_layers,
        num_attention_heads=d.heads, num_key_value_heads=d.heads, sliding_windows=[list(w) for w in d.windows],
        max_position_embeddings=d.max_pos_emb)
    cfg = MoonshineStreamingConfig(
        encoder_config=enc, vocab_size=d.vocab, hidden_size=d.dim, intermediate_size=d.ffn,
        num_hidden_layers=d.dec_layers, num_attention_heads=d.heads, num_key_value_heads=d.heads,
        max_position_embeddings=d.max_pos_emb,
        rope_parameters={"rope_type": "default", "rope_theta": d.rope_theta,
                         "partial_rotary_factor": d.rope_factor},
        tie_word_embeddings=False, attn_implementation="eager")
    assert cfg.head_dim == d.head_dim
    m = MoonshineStreamingForConditionalGeneration(cfg).eval()
    sd = {k: torch.from_numpy(v) for k, v in weights.items() if k != STREAMING_CONFIG_NAME}
    sd["model.encoder.embedder.comp.log_k"] = sd["model.encoder.embedder.comp.log_k"].reshape(())
    if d.tied:
        sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in k or "inv_freq" in k for k in missing), missing
    return m


@torch.no_grad()
def hf_memory_inputs(m, d, pcm):
    """Encoder output (pre-adapter) over the analysed samples, export.py Encoder semantics."""
    enc = m.model.encoder
    feats, _ = enc.embedder(torch.from_numpy(pcm)[None])
    T = feats.shape[1]
    pos = torch.arange(T)
    dist = pos[:, None] - pos[None, :]
    h = feats
    for layer, (past, future) in zip(enc.layers, d.windows):
        allowed = (dist >= -future) & (dist <= past)
        mask = torch.zeros(1, 1, T, T).masked_fill(~allowed[None, None], torch.finfo(torch.float32).min)
        h = layer(h, attention_mask=mask)
    return enc.final_norm(h)


@torch.no_grad()
def hf_greedy(m, encoded, max_tokens):
    """Greedy ids + logits; the HF decoder applies pos_emb + proj to its encoder states itself."""
    dec = m.model.decoder
    tokens, logits = [1], []
    for _ in range(max_tokens):
        hid = dec(input_ids=torch.tensor([tokens]), encoder_hidden_states=encoded.clone(),
                  use_cache=False).last_hidden_state
        lg = m.proj_out(hid)[0, -1]
        logits.append(lg.numpy().copy())
        nxt = int(torch.argmax(lg))
        tokens.append(nxt)
        if nxt == 2:
            break
    return tokens, np.stack(logits)


def main():
    torch.set_num_threads(8)
    cases = [
        # arch, init, seed, input name, n_samples, final
        ("tiny_streaming", "scaled", 0, "synth0", 16000 * 3 + 700, True),
        ("tiny_streaming", "scaled", 0, "synth0", 16000 * 3 + 700, False),
        ("tiny_streaming", "hf", 0, "synth2", 16000 * 5, True),
        ("base_streaming", "scaled", 1, "synth1", 16000 * 2 + 11, True),
        ("test_streaming", "scaled", 0, "synth1", 16000 * 2 + 999, True),
        ("test_streaming", "scaled", 0, "synth1", 16000 * 2 + 999, False),
        ("test_streaming2", "scaled", 3, "synth3", 16000 * 4, True),
    ]
    for arch, init, seed, inp, n, final in cases:
        d = ARCHS[arch]
        w = synth_weights(arch, seed, init)
        m = hf_model(d, w)
        pcm = synth_audio(int(inp[5:]), n)
        processed, nfeat, emitted = streaming_lengths(n, is_final=final, lookahead=d.total_lookahead)
        encoded = hf_memory_inputs(m, d, pcm[:processed])
        assert encoded.shape[1] == nfeat
        enc_e = encoded[:, :emitted]
        dec = m.model.decoder
        with torch.no_grad():
            memory = dec.proj(enc_e + dec.pos_emb(torch.arange(emitted)))[0].numpy()
        dur = np.float32(n) / np.float32(16000.0)
        max_tokens = min(int(math.ceil(float(dur * np.float32(6.5)))), 256)
        tokens, logits = hf_greedy(m, enc_e, max_tokens)
        srt = np.sort(logits, axis=1)
        top_idx = np.argsort(-logits, axis=1, kind="stable")[:, :8]
        name = f"hfs_{arch}_{init}_s{seed}_{inp}_{n}_{'final' if final else 'partial'}.npz"
        np.savez_compressed(
            os.path.join(OUT, name),
            mem_sub=memory[::4].astype(np.float32), mem_shape=np.array(memory.shape),
            mem_absmax=np.float32(np.abs(memory).max()),
            tokens=np.array(tokens, np.int32),
            logits_sub=logits[:, ::64].astype(np.float32),
            logits_absmax=np.abs(logits).max(axis=1).astype(np.float32),
            top_idx=top_idx.astype(np.int32),
            top_val=np.take_along_axis(logits, top_idx, 1).astype(np.float32),
            margin=(srt[:, -1] - srt[:, -2]).astype(np.float32),
            n_samples=np.int64(n), emitted=np.int64(emitted), final=np.int64(final),
        )
        print(name, "memory", memory.shape, "steps", len(tokens) - 1, "min margin/absmax",
              float(((srt[:, -1] - srt[:, -2]) / np.abs(logits).max(axis=1)).min()), "tokens", tokens[:8])


if __name__ == "__main__":
    main()
