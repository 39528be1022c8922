#!/usr/bin/env python
"""Golden fixtures for the streaming architectures from the REFERENCE's own graph modules, driven the way the
reference runtime drives its graphs: chunk by chunk, with carried state.

Run in the BUILD container only (needs transformers + /root/reference):
    python tests/golden/make_golden_streaming_ref.py

The five ONNX graphs the reference executes (frontend / encoder / adapter / cross_kv / decoder_kv) are traced from
the torch modules `Frontend`, `Encoder`, `Adapter`, `CrossKV`, `DecoderKV` of
language-bindings/python/src/moonshine_voice/lora/export.py:53-256.  This script IMPORTS those modules from
/root/reference (nothing is copied), wraps them around an HF `MoonshineStreamingForConditionalGeneration` carrying
this repo's seeded synthetic weights, and drives them with the host control of
core/moonshine-streaming-model.cpp (restated here in a few lines, cited per step):
  * process_audio_chunk (:441-602): 1280-sample chunks through Frontend with sample / conv1 / conv2 carry-over state;
  * encode (:604-772): stable = total - lookahead (all when final); window_start = max(0, emitted - 16 * depth);
    Encoder over [window_start, total); Adapter on the new frames with pos_offset = emitted; memory append;
  * compute_cross_kv (:779-860) once per decode; run_decoder_with_cross_kv (:867-1082) one token at a time with
    the stacked self K/V growing by one position per step.
The audio arrives in uneven pieces (several non-final updates, then the final one), so the fixtures pin
"one stateless pass == the reference's chunked state machine" to reference code, not to this repo's oracle.

Writes tests/golden/refx_<arch>_<init>_s<seed>_<input>_<n>.npz:
    mem_sub (every 4th memory row), mem_shape, mem_absmax, tokens, logits_sub / top_idx / top_val / margin /
    logits_absmax, n_samples, pieces (the update pattern), mem_len_after_update (memory rows after every update),
    verify_logits_sub (ONE decoder_kv call over the first 6 ids = decode_tokens, :1136-1190).
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_PY = "/root/reference/language-bindings/python/src"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF_PY)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from moonshine_b200.arch import ARCHS  # noqa: E402
from moonshine_b200.weights import synth_audio, synth_weights  # noqa: E402
from make_golden_streaming import hf_model  # noqa: E402
from moonshine_voice.lora import export as ref_export  # noqa: E402  (the reference's own modules)

OUT = os.path.join(ROOT, "tests", "golden")
CHUNK = 1280


class ReferenceStreamingSession:
    """Host control of MoonshineStreamingModel around the reference's graph modules (torch, no ONNX)."""

    def __init__(self, m, d):
        enc, dec = m.model.encoder, m.model.decoder
        self.d = d
        self.frontend = ref_export.Frontend(enc.embedder).eval()
        self.encoder = ref_export.Encoder(enc, [tuple(w) for w in d.windows]).eval()
        self.adapter = ref_export.Adapter(dec).eval()
        self.cross_kv = ref_export.CrossKV(dec.layers, d.heads, d.head_dim).eval()
        self.decoder_kv = ref_export.DecoderKV(dec, m.proj_out, d.heads, d.head_dim, dec.rotary_emb).eval()
        self.depth = d.dec_layers            # streaming_config.json "depth" (export.py:370, :456)
        self.lookahead = sum(f for _, f in d.windows)
        E = d.enc_dim
        shapes = ref_export.frames_state_shapes(enc.embedder, E)
        self.sample_buffer = torch.zeros(*shapes["sample_buffer"])
        self.sample_len = torch.zeros(1, dtype=torch.long)
        self.conv1_buffer = torch.zeros(*shapes["conv1_buffer"])
        self.conv2_buffer = torch.zeros(*shapes["conv2_buffer"])
        self.frame_count = torch.zeros(1, dtype=torch.long)
        self.features = torch.zeros(1, 0, E)
        self.memory = torch.zeros(1, 0, d.dim)
        self.emitted = 0
        self.processed = 0

    @torch.no_grad()
    def add_chunk(self, chunk):
        f, self.sample_buffer, self.sample_len, self.conv1_buffer, self.conv2_buffer, self.frame_count = self.frontend(
            torch.from_numpy(chunk)[None], self.sample_buffer, self.sample_len, self.conv1_buffer, self.conv2_buffer,
            self.frame_count)
        self.features = torch.cat([self.features, f], dim=1)

    @torch.no_grad()
    def encode(self, is_final):
        total = self.features.shape[1]
        if total == 0:
            return
        stable = total if is_final else max(0, total - self.lookahead)
        new = stable - self.emitted
        if new <= 0:
            return
        window_start = max(0, self.emitted - 16 * self.depth)
        encoded = self.encoder(self.features[:, window_start:])
        start = self.emitted - window_start
        mem = self.adapter(encoded[:, start:start + new], torch.tensor([self.emitted]))
        self.memory = torch.cat([self.memory, mem], dim=1)
        self.emitted += new

    def update(self, audio_so_far, is_final):
        """Transcriber::transcribe_segment_with_streaming_model (core/transcriber.cpp:1331-1372): only whole new
        1280-sample chunks are analysed; no new chunk -> the encoder does not run."""
        n = len(audio_so_far)
        if self.processed < n:
            new_chunks = (n - self.processed) // CHUNK
            if new_chunks > 0:
                for i in range(new_chunks):
                    self.add_chunk(audio_so_far[self.processed + i * CHUNK: self.processed + (i + 1) * CHUNK])
                self.processed += new_chunks * CHUNK
                self.encode(is_final)

    @torch.no_grad()
    def greedy(self, max_tokens):
        d = self.d
        k_cross, v_cross = self.cross_kv(self.memory)
        k_self = torch.zeros(d.dec_layers, 1, d.heads, 0, d.head_dim)
        v_self = torch.zeros(d.dec_layers, 1, d.heads, 0, d.head_dim)
        tokens, logits = [d.bos], []
        cur = d.bos
        for _ in range(max_tokens):
            lg, k_self, v_self, _, _ = self.decoder_kv(torch.tensor([[cur]]), k_self, v_self, k_cross, v_cross)
            row = lg[0, -1].numpy().copy()
            logits.append(row)
            cur = int(np.argmax(row))
            tokens.append(cur)
            if cur == d.eos:
                break
        return tokens, np.stack(logits)

    @torch.no_grad()
    def decode_tokens(self, ids):
        """ONE decoder_kv call over several ids from an empty cache (teacher-forced logits of every position)."""
        d = self.d
        k_cross, v_cross = self.cross_kv(self.memory)
        k0 = torch.zeros(d.dec_layers, 1, d.heads, 0, d.head_dim)
        lg, _, _, _, _ = self.decoder_kv(torch.tensor([ids]), k0, k0.clone(), k_cross, v_cross)
        return lg[0].numpy().copy()


def main():
    torch.set_num_threads(8)
    cases = [
        # arch, init, seed, input, n_samples, piece boundaries (fractions of the clip; the last update is final)
        ("tiny_streaming", "scaled", 0, "synth0", 16000 * 4 + 700, (0.17, 0.5, 0.52, 0.9)),
        ("base_streaming", "scaled", 1, "synth1", 16000 * 3 + 11, (0.4, 0.8)),
        ("test_streaming", "scaled", 0, "synth1", 16000 * 6 + 999, (0.1, 0.3, 0.31, 0.6, 0.95)),
        ("test_streaming2", "scaled", 3, "synth3", 16000 * 4, (0.25, 0.5, 0.75)),
    ]
    for arch, init, seed, inp, n, cuts in cases:
        d = ARCHS[arch]
        m = hf_model(d, synth_weights(arch, seed, init))
        pcm = synth_audio(int(inp[5:]), n)
        sess = ReferenceStreamingSession(m, d)
        ends = [int(n * c) for c in cuts] + [n]
        mem_after = []
        for i, e in enumerate(ends):
            sess.update(pcm[:e], is_final=(i == len(ends) - 1))
            mem_after.append(sess.memory.shape[1])
        memory = sess.memory[0].numpy()
        dur = np.float32(n) / np.float32(16000.0)
        max_tokens = min(int(math.ceil(float(dur * np.float32(6.5)))), 256)   # core/transcriber.cpp:1386-1390
        tokens, logits = sess.greedy(max_tokens)
        verify = sess.decode_tokens(tokens[:6])
        srt = np.sort(logits, axis=1)
        top_idx = np.argsort(-logits, axis=1, kind="stable")[:, :8]
        name = f"refx_{arch}_{init}_s{seed}_{inp}_{n}.npz"
        np.savez_compressed(
            os.path.join(OUT, name),
            mem_sub=memory[::4].astype(np.float32), mem_shape=np.array(memory.shape),
            mem_absmax=np.float32(np.abs(memory).max()), tokens=np.array(tokens, np.int32),
            logits_sub=logits[:, ::64].astype(np.float32), logits_absmax=np.abs(logits).max(axis=1).astype(np.float32),
            top_idx=top_idx.astype(np.int32), top_val=np.take_along_axis(logits, top_idx, 1).astype(np.float32),
            margin=(srt[:, -1] - srt[:, -2]).astype(np.float32), n_samples=np.int64(n),
            pieces=np.array(ends, np.int64), mem_len_after_update=np.array(mem_after, np.int64),
            verify_logits_sub=verify[:, ::64].astype(np.float32), verify_absmax=np.abs(verify).max(axis=1).astype(np.float32),
        )
        print(name, "memory", memory.shape, "updates", mem_after, "steps", len(tokens) - 1, "min margin/absmax",
              float(((srt[:, -1] - srt[:, -2]) / np.abs(logits).max(axis=1)).min()), "tokens", tokens[:6])


if __name__ == "__main__":
    main()
