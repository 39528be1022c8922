#!/usr/bin/env python
"""Generate golden fixtures from the Hugging Face float implementation.

Run in the BUILD container only (needs transformers + /root/reference):
    python tests/golden/make_golden.py

The reference tree holds no arithmetic and no weights for this path (see
oracle/moonshine_oracle.py header), so the pin is the HF float model the
reference's graphs were exported from, loaded with this repo's seeded synthetic
weights (moonshine_b200.weights.synth_weights) -- reproducible anywhere from
(arch, seed, init) alone, which is why no weight file is committed.

Writes (all small):
  tests/golden/beckett_pcm16.npy      the reference's test-assets/beckett.wav
                                      samples (int16 mono 16 kHz) -- BASELINE
                                      config #1's input
  tests/golden/hf_<arch>_<init>_s<seed>_<input>.npz   per case:
      enc_out[::4]  (every 4th encoder frame, fp32), enc_shape,
      tokens        (greedy ids incl. start token),
      logits_sub    (steps x every-64th-vocab-entry),
      top_idx/top_val (top-8 logits per step),
      margin        (top1-top2 per step)
"""
import os
import sys
import wave

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from moonshine_b200.arch import ARCHS  # noqa: E402
from moonshine_b200.weights import synth_audio, synth_weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def load_wav_int16(path):
    with wave.open(path, "rb") as w:
        assert w.getnchannels() == 1 and w.getsampwidth() == 2 and w.getframerate() == 16000
        return np.frombuffer(w.readframes(w.getnframes()), np.int16).copy()


def hf_model(dims, weights):
    from transformers import MoonshineConfig, MoonshineForConditionalGeneration
    cfg = MoonshineConfig(
        vocab_size=dims.vocab, hidden_size=dims.dim, intermediate_size=dims.ffn,
        encoder_num_hidden_layers=dims.enc_layers, decoder_num_hidden_layers=dims.dec_layers,
        encoder_num_attention_heads=dims.heads, decoder_num_attention_heads=dims.heads,
        partial_rotary_factor=dims.rope_factor,
        rope_parameters={"rope_type": "default", "rope_theta": dims.rope_theta,
                         "partial_rotary_factor": dims.rope_factor},
        attn_implementation="eager",
    )
    assert cfg.hidden_size // cfg.encoder_num_attention_heads == dims.head_dim
    m = MoonshineForConditionalGeneration(cfg).eval()
    sd = {k: torch.from_numpy(v) for k, v in weights.items()}
    sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in k or "inv_freq" in k for k in missing), missing
    return m


@torch.no_grad()
def hf_greedy(m, pcm, max_len):
    x = torch.from_numpy(pcm)[None]
    enc_out = m.model.encoder(x)
    enc = enc_out.last_hidden_state
    ids = torch.tensor([[1]])
    past = None
    tokens, logits = [1], []
    for _ in range(max_len):
        out = m(encoder_outputs=enc_out, decoder_input_ids=ids, past_key_values=past, use_cache=True)
        past = out.past_key_values
        lg = out.logits[0, -1]
        logits.append(lg.numpy().copy())
        nxt = int(torch.argmax(lg))
        tokens.append(nxt)
        if nxt == 2:
            break
        ids = torch.tensor([[nxt]])
    return enc[0].numpy(), tokens, np.stack(logits)


def main():
    torch.set_num_threads(8)
    beck = load_wav_int16("/root/reference/test-assets/beckett.wav")
    np.save(os.path.join(OUT, "beckett_pcm16.npy"), beck)
    inputs = {
        "beckett": beck.astype(np.float32) / np.float32(32768.0),
        "synth0": synth_audio(0),
        "synth1short": synth_audio(1, 48000 + 333),
    }
    cases = [
        ("tiny", "scaled", 0, "beckett"), ("tiny", "scaled", 0, "synth0"),
        ("tiny", "hf", 0, "synth0"), ("tiny", "scaled", 1, "synth1short"),
        ("base", "scaled", 0, "synth0"), ("base", "scaled", 0, "beckett"),
        ("test", "scaled", 0, "synth1short"), ("test2", "scaled", 3, "synth1short"),
    ]
    import math
    for arch, init, seed, inp in cases:
        dims = ARCHS[arch]
        w = synth_weights(arch, seed, init)
        m = hf_model(dims, w)
        pcm = inputs[inp]
        dur = np.float32(len(pcm)) / np.float32(16000.0)
        max_len = int(math.ceil(float(dur * np.float32(6.5))))
        enc, tokens, logits = hf_greedy(m, pcm, max_len)
        srt = np.sort(logits, axis=1)
        top_idx = np.argsort(-logits, axis=1, kind="stable")[:, :8]
        name = f"hf_{arch}_{init}_s{seed}_{inp}.npz"
        np.savez_compressed(
            os.path.join(OUT, name),
            enc_sub=enc[::4].astype(np.float32), enc_shape=np.array(enc.shape),
            enc_absmax=np.float32(np.abs(enc).max()),
            tokens=np.array(tokens, np.int32),
            logits_sub=logits[:, ::64].astype(np.float32),
            logits_absmax=np.abs(logits).max(axis=1).astype(np.float32),
            top_idx=top_idx.astype(np.int32),
            top_val=np.take_along_axis(logits, top_idx, 1).astype(np.float32),
            margin=(srt[:, -1] - srt[:, -2]).astype(np.float32),
            n_samples=np.int64(len(pcm)),
        )
        print(name, "enc", enc.shape, "steps", len(tokens) - 1, "min margin/absmax",
              float(((srt[:, -1] - srt[:, -2]) / np.abs(logits).max(axis=1)).min()),
              "tokens", tokens[:8])


if __name__ == "__main__":
    main()
