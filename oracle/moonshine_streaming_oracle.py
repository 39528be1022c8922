"""CPU oracle for Moonshine's *streaming* architectures (numpy).

TEST INFRASTRUCTURE ONLY (same rule as moonshine_oracle.py: only ``tests/``,
``__graft_entry__.smoke()`` and bench.py's CPU legs may use it).

What it restates
----------------
The reference runs five ONNX graphs per streaming model (frontend, encoder,
adapter, cross_kv, decoder_kv; core/moonshine-streaming-model.cpp:441-1082)
whose weights are downloaded, not in the tree.  The graphs' arithmetic is
defined in-tree by the exporter that produces them,
``language-bindings/python/src/moonshine_voice/lora/export.py`` (cited ``export.py:line``),
wrapping modules of HF Transformers 5.5.0
``transformers/models/moonshine_streaming/modeling_moonshine_streaming.py`` (cited ``HFS:line``).
Host semantics (1280-sample chunks, look-ahead hold-back, token budget, EOS handling)
follow core/transcriber.cpp:1311-1487 and core/moonshine-streaming-model.cpp:604-772,1192-1397.

Two computations of the encoder memory are provided and tested equal:
``memory_stateless`` (what the CUDA path does: one pass over the analysed audio) and
``ChunkedState`` (the reference's chunk-by-chunk state machine, export.py:53-97 +
moonshine-streaming-model.cpp:604-772), so the batched GPU path is pinned to the incremental one.

Pinning
-------
``tests/golden/make_golden_streaming.py`` runs the HF modules themselves (with export.py's inclusive
window masks) in the build container and stores encoder memory / logits / ids in
``tests/golden/hfs_*.npz``.  Parity with the shipped int8 .ort graphs is unpinned (absent files).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

try:
    from .moonshine_oracle import Oracle
except ImportError:  # run as a plain module from oracle/
    from moonshine_oracle import Oracle

CHUNK = 1280     # core/transcriber.cpp:1342
FRAME = 80       # HFS:286 frame_len = sample_rate * frame_ms / 1000
FEATURE = 320    # two stride-2 convs over 80-sample frames


class SDims:
    """Streaming dimensions (HF MoonshineStreamingConfig + encoder_config)."""

    def __init__(self, enc_dim, dim, enc_layers, dec_layers, heads, head_dim, enc_ffn, ffn, windows,
                 vocab=32768, rope_factor=0.8, rope_theta=10000.0, tied=False, bos=1, eos=2,
                 max_seq_len=448):
        self.enc_dim, self.dim, self.enc_layers, self.dec_layers = enc_dim, dim, enc_layers, dec_layers
        self.heads, self.head_dim, self.enc_ffn, self.ffn = heads, head_dim, enc_ffn, ffn
        self.windows = [tuple(w) for w in windows]
        self.vocab, self.rope_factor, self.rope_theta = vocab, rope_factor, rope_theta
        self.tied, self.bos, self.eos, self.max_seq_len = tied, bos, eos, max_seq_len
        self.lookahead = sum(f for _, f in self.windows)   # streaming_config.json total_lookahead

    @classmethod
    def from_product(cls, d):
        return cls(d.enc_dim, d.dim, d.enc_layers, d.dec_layers, d.heads, d.head_dim, d.enc_ffn, d.ffn,
                   d.windows, d.vocab, d.rope_factor, d.rope_theta, d.tied, d.bos, d.eos, d.max_seq_len)


class StreamingOracle(Oracle):
    def __init__(self, dims: SDims, weights: Dict[str, np.ndarray], dtype=np.float32, emulate=None):
        super().__init__(dims, weights, dtype, emulate)

    # ---- frontend: HFS:62-81 (CMVN, asinh), HFS:275-307 / export.py:53-97 ----------------
    def frames_to_hidden(self, frames: np.ndarray) -> np.ndarray:
        """frames [n, 80] -> SiLU(Linear(asinh(exp(log_k) * CMVN(frame)))) [n, E]."""
        x = np.asarray(frames, self.dt)
        mean = x.mean(-1, keepdims=True)
        c = x - mean
        rms = np.sqrt((c * c).mean(-1, keepdims=True) + self.dt(1e-6))
        x = c / rms
        k = np.exp(self.w["model.encoder.embedder.comp.log_k"].reshape(()))
        x = np.arcsinh(k * x).astype(self.dt)
        return self.silu(self.linear(x, "model.encoder.embedder.linear.weight"))

    def causal_conv(self, x: np.ndarray, wname: str, bname: str, left: Optional[np.ndarray] = None):
        """x [T, C_in] channel-last; kernel 5, stride 2, four frames of left context
        (zeros at the segment start, HFS:83-98; the carried buffer when chunked, export.py:75-80)."""
        w = self.w[wname]                      # [C_out, C_in, 5]
        co, ci, k = w.shape
        if left is None:
            left = np.zeros((k - 1, ci), self.dt)
        xin = np.concatenate([left, x], 0)
        t_out = (xin.shape[0] - k) // 2 + 1 if xin.shape[0] >= k else 0
        idx = np.arange(t_out)[:, None] * 2 + np.arange(k)[None, :]
        cols = xin[idx]                         # [t_out, k, ci]
        wm = w.transpose(2, 1, 0).reshape(k * ci, co)
        return (self.mm(cols.reshape(t_out, k * ci), wm) + self.w[bname]).astype(self.dt)

    def features(self, pcm: np.ndarray) -> np.ndarray:
        """Whole-utterance frontend over the given samples (all whole frames)."""
        n = len(pcm) // FRAME
        h = self.frames_to_hidden(np.asarray(pcm[: n * FRAME]).reshape(n, FRAME))
        e = "model.encoder.embedder."
        c1 = self.silu(self.causal_conv(h, e + "conv1.weight", e + "conv1.bias"))
        return self.causal_conv(c1, e + "conv2.weight", e + "conv2.bias")

    # ---- encoder: export.py:100-127 (inclusive windows), HFS:112-122 (unit-offset LN),
    #      HFS:177-272 (layer) ------------------------------------------------------------
    def unit_ln(self, x, gname, eps=1e-5):
        mu = x.mean(-1, keepdims=True)
        var = ((x - mu) ** 2).mean(-1, keepdims=True)
        return ((x - mu) / np.sqrt(var + self.dt(eps)) * (self.w[gname] + self.dt(1.0))).astype(self.dt)

    def _eheads(self, x):
        H = self.d.heads
        return x.reshape(x.shape[0], H, x.shape[1] // H).transpose(1, 0, 2)

    def encode_features(self, feats: np.ndarray) -> np.ndarray:
        h = np.asarray(feats, self.dt)
        T = h.shape[0]
        dist = np.arange(T)[:, None] - np.arange(T)[None, :]          # q - k
        ehd = self.d.enc_dim // self.d.heads
        for l, (past, future) in enumerate(self.d.windows):
            p = f"model.encoder.layers.{l}."
            allowed = (dist >= -future) & (dist <= past)
            x = self.unit_ln(h, p + "input_layernorm.gamma")
            q = self._eheads(self.linear(x, p + "self_attn.q_proj.weight"))
            k = self._eheads(self.linear(x, p + "self_attn.k_proj.weight"))
            v = self._eheads(self.linear(x, p + "self_attn.v_proj.weight"))
            s = np.einsum("hqd,hkd->hqk", self._rnd(q), self._rnd(k)).astype(self.dt) * self.dt(ehd ** -0.5)
            s = np.where(allowed[None], s, -np.inf)
            pr = self.softmax(s).astype(self.dt)
            a = np.einsum("hqk,hkd->hqd", self._rnd(pr), self._rnd(v)).astype(self.dt)
            a = a.transpose(1, 0, 2).reshape(T, -1)
            h = h + self.linear(a, p + "self_attn.o_proj.weight")
            x = self.unit_ln(h, p + "post_attention_layernorm.gamma")
            x = self.gelu(self.linear(x, p + "mlp.fc1.weight", p + "mlp.fc1.bias"))
            h = h + self.linear(x, p + "mlp.fc2.weight", p + "mlp.fc2.bias")
        return self.unit_ln(h, "model.encoder.final_norm.gamma")

    # ---- adapter: export.py:130-144 ------------------------------------------------------
    def adapt(self, encoded: np.ndarray, pos_offset: int = 0) -> np.ndarray:
        pos = np.arange(encoded.shape[0]) + pos_offset
        x = (encoded + self.w["model.decoder.pos_emb.weight"][pos]).astype(self.dt)
        if "model.decoder.proj.weight" in self.w:
            x = self.linear(x, "model.decoder.proj.weight")
        return x

    def memory_stateless(self, pcm: np.ndarray, n_features: int, emitted: int) -> np.ndarray:
        """Encoder memory of a segment in one pass: the first ``emitted`` rows of
        adapter(encoder(features of the first n_features*320 samples))."""
        feats = self.features(pcm[: n_features * FEATURE])
        assert feats.shape[0] == n_features
        return self.adapt(self.encode_features(feats))[:emitted]

    # ---- decoder: export.py:170-256 (same block as the non-streaming decoder; untied head) --
    def head(self, h):
        # export.py:206-211,253-254: embedding when tied, else the separate output projection
        if self.d.tied:
            return super().head(h)
        return self.mm(h, self.w["proj_out.weight"].T)

    # ---- host loop ---------------------------------------------------------------------------
    @staticmethod
    def max_tokens_greedy(n_samples: int, max_tokens_per_second: float = 6.5) -> int:
        # core/transcriber.cpp:1386-1390 (float duration, cap 256)
        dur = np.float32(n_samples) / np.float32(16000.0)
        return min(int(math.ceil(float(np.float32(dur) * np.float32(max_tokens_per_second)))), 256)

    def max_tokens_speculative(self, memory_len: int) -> int:
        # core/moonshine-streaming-model.cpp:1217-1219 (float duration, double 6.5)
        dur = np.float32(memory_len) * np.float32(0.020)
        return min(int(math.ceil(float(dur) * 6.5)), self.d.max_seq_len)

    def greedy_memory(self, memory: np.ndarray, max_tokens: int, forced=None, keep_logits=True,
                      append_eos=True):
        cross = self.cross_kv(memory)
        cache = self.new_self_cache()
        tokens = [self.d.bos]
        cur = self.d.bos
        logits = []
        for t in range(max_tokens):
            lg = self.decoder_step([cur], t, cache, cross)[0]
            if keep_logits:
                logits.append(lg)
            nxt = self.argmax_first(lg) if forced is None else int(forced[t])
            if nxt == self.d.eos and not append_eos:
                break
            tokens.append(nxt)
            if nxt == self.d.eos:
                break
            cur = nxt
        return tokens, (np.stack(logits) if logits else None)

    def transcribe_segment(self, pcm: np.ndarray, is_final: bool = True, state: Optional[dict] = None,
                           max_tokens_per_second: float = 6.5, speculative: bool = False, **kw):
        """One call of Transcriber::transcribe_segment_with_streaming_model on the segment's audio so far.
        ``state`` carries {"processed", "emitted", "decoded"} between calls of the same segment."""
        st = state if state is not None else {}
        processed, emitted = st.get("processed", 0), st.get("emitted", 0)
        L = len(pcm)
        if processed < L:
            processed += (L - processed) // CHUNK * CHUNK
            n = processed // FEATURE
            stable = n if is_final else max(0, n - self.d.lookahead)
            if n > 0 and stable > emitted:
                emitted = stable
        st["processed"], st["emitted"] = processed, emitted
        if emitted == 0:
            return [], None, np.zeros((0, self.d.dim), self.dt)
        mem = self.memory_stateless(pcm, processed // FEATURE, emitted)
        if speculative and st.get("decoded"):
            # decode_full: verify-then-continue == greedy in exact arithmetic; EOS is not appended and the
            # budget comes from the memory length
            toks, lg = self.greedy_memory(mem, self.max_tokens_speculative(emitted), append_eos=False, **kw)
            toks = toks[: 1 + self.max_tokens_speculative(emitted)]
        else:
            toks, lg = self.greedy_memory(mem, self.max_tokens_greedy(L, max_tokens_per_second), **kw)
        st["decoded"] = True
        return toks, lg, mem


class ChunkedState:
    """The reference's incremental state machine, restated: frontend carry-over buffers
    (export.py:53-97), accumulated features, windowed re-encode with 16*depth frames of left
    context and look-ahead hold-back (moonshine-streaming-model.cpp:604-772), running adapter offset."""

    def __init__(self, oracle: StreamingOracle):
        o = self.o = oracle
        E = o.d.enc_dim
        self.sample_buffer = np.zeros(0, o.dt)
        self.conv1_buffer = np.zeros((4, E), o.dt)
        self.conv2_buffer = np.zeros((4, 2 * E), o.dt)
        self.feats = np.zeros((0, E), o.dt)
        self.emitted = 0
        self.pos_offset = 0
        self.memory = np.zeros((0, o.d.dim), o.dt)

    def process_audio_chunk(self, chunk: np.ndarray):
        o = self.o
        e = "model.encoder.embedder."
        comb = np.concatenate([self.sample_buffer, np.asarray(chunk, o.dt)])
        nf = len(comb) // FRAME
        hidden = o.frames_to_hidden(comb[: nf * FRAME].reshape(nf, FRAME))
        c1 = o.silu(o.causal_conv(hidden, e + "conv1.weight", e + "conv1.bias", left=self.conv1_buffer))
        f = o.causal_conv(c1, e + "conv2.weight", e + "conv2.bias", left=self.conv2_buffer)
        self.sample_buffer = comb[nf * FRAME:]
        self.conv1_buffer = np.concatenate([self.conv1_buffer, hidden])[-4:]
        self.conv2_buffer = np.concatenate([self.conv2_buffer, c1])[-4:]
        self.feats = np.concatenate([self.feats, f])

    def encode(self, is_final: bool) -> int:
        o = self.o
        total = self.feats.shape[0]
        if total == 0:
            return 0
        stable = total if is_final else max(0, total - o.d.lookahead)
        new = stable - self.emitted
        if new <= 0:
            return 0
        ws = max(0, self.emitted - 16 * o.d.dec_layers)     # `16 * config.depth`, depth = decoder layers
        enc = o.encode_features(self.feats[ws:])
        s = self.emitted - ws
        mem = o.adapt(enc[s: s + new], self.pos_offset)
        self.memory = np.concatenate([self.memory, mem])
        self.pos_offset += new
        self.emitted = stable
        return new
