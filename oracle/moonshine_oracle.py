"""CPU oracle for Moonshine's encoder-decoder transcription loop (numpy).

TEST INFRASTRUCTURE ONLY.  Nothing under ``moonshine_b200/`` may import this;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg use it, and only as the checker / CPU baseline.

What it restates
----------------
The reference (``/root/reference``) holds no arithmetic for this path: its host
loops (core/moonshine-model.cpp:215-560) marshal tensors into ONNX Runtime
1.23.2 (vendored prebuilt ``libonnxruntime.so.1``, no source) executing
``encoder_model.ort`` / ``decoder_model_merged.ort`` graphs that are downloaded
at install time and are NOT in the tree (scripts/fetch-voice-assets.sh:85-112).
Those graphs were exported (Optimum) from the Hugging Face float model, which
the reference's own docs name as its accuracy oracle
(docs/models/accuracy.md:14-19).  So:

* arithmetic follows HF Transformers 5.5.0
  ``transformers/models/moonshine/modeling_moonshine.py`` (cited as ``HF:line``);
* host-loop semantics (max_len, start/EOS handling, first-max argmax,
  detokenisation, VAD-bypass segmentation) follow the reference C++ (cited as
  ``core/...:line``).

Pinning
-------
Pinned against outputs of the HF implementation itself run in the build
container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``) on the
reference's own ``test-assets/beckett.wav`` and on the synthetic BASELINE
inputs.  The reference's tests hold no numeric golden vectors for this path
(only ``"fail" in transcript`` for beckett.wav with the real, absent, weights:
language-bindings/python/tests/test_modules.py:62-69), and the shipped int8
``.ort`` graphs cannot be run here, so parity with the *ORT int8 CPU path* is
unpinned; parity with the float model the graphs were exported from is pinned.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

try:  # exact-erf GELU (HF "gelu" == torch erf GELU)
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])


# --------------------------------------------------------------------------
# dims (kept independent of the product package on purpose)
# --------------------------------------------------------------------------
class Dims:
    def __init__(self, dim, enc_layers, dec_layers, heads, head_dim, ffn,
                 vocab=32768, rope_factor=0.9, rope_theta=10000.0, bos=1, eos=2):
        self.dim, self.enc_layers, self.dec_layers = dim, enc_layers, dec_layers
        self.heads, self.head_dim, self.ffn, self.vocab = heads, head_dim, ffn, vocab
        self.rope_factor, self.rope_theta, self.bos, self.eos = rope_factor, rope_theta, bos, eos

    @classmethod
    def from_product(cls, d):
        return cls(d.dim, d.enc_layers, d.dec_layers, d.heads, d.head_dim, d.ffn,
                   d.vocab, d.rope_factor, d.rope_theta, d.bos, d.eos)


# core/moonshine-model.cpp:41-79 (set_model_options_from_arch) + HF config
TINY = Dims(288, 6, 6, 8, 36, 1152)
BASE = Dims(416, 8, 8, 8, 52, 1664)


# --------------------------------------------------------------------------
# operand-rounding emulation (precision studies only; default = exact)
# --------------------------------------------------------------------------
def _round_mantissa(x: np.ndarray, keep_bits: int) -> np.ndarray:
    """Round-to-nearest-even an fp32 array to `keep_bits` explicit mantissa
    bits (10 = tf32, 7 = bf16)."""
    x = np.ascontiguousarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    drop = 23 - keep_bits
    half = (1 << (drop - 1)) - 1
    u = u + half + ((u >> drop) & 1)
    u = (u >> drop) << drop
    return u.astype(np.uint32).view(np.float32)


class Oracle:
    """fp32 (default) or fp64 restatement.  ``emulate`` in {None,"tf32","bf16",
    "bf16x2"} rounds matmul operands to study tensor-core input precision."""

    def __init__(self, dims: Dims, weights: Dict[str, np.ndarray],
                 dtype=np.float32, emulate: Optional[str] = None):
        self.d = dims
        self.dt = dtype
        self.emulate = emulate
        self.w = {k: np.asarray(v, dtype) for k, v in weights.items()}

    # ---- primitives ------------------------------------------------------
    def _rnd(self, a):
        if self.emulate is None:
            return a
        a32 = np.asarray(a, np.float32)
        if self.emulate == "tf32":
            return _round_mantissa(a32, 10).astype(self.dt)
        if self.emulate == "bf16":
            return _round_mantissa(a32, 7).astype(self.dt)
        if self.emulate == "bf16x2":  # hi + lo split, ~16 mantissa bits
            hi = _round_mantissa(a32, 7)
            lo = _round_mantissa(a32 - hi, 7)
            return (hi.astype(np.float64) + lo.astype(np.float64)).astype(self.dt)
        raise ValueError(self.emulate)

    def mm(self, a, b):
        return (self._rnd(a) @ self._rnd(b)).astype(self.dt)

    def linear(self, x, wname, bname=None):
        y = self.mm(x, self.w[wname].T)
        if bname is not None:
            y = y + self.w[bname]
        return y

    def layernorm(self, x, wname, eps=1e-5):
        # nn.LayerNorm(D, bias=False), HF:381-382,436-438,539
        mu = x.mean(-1, keepdims=True)
        var = ((x - mu) ** 2).mean(-1, keepdims=True)
        return ((x - mu) / np.sqrt(var + self.dt(eps)) * self.w[wname]).astype(self.dt)

    def gelu(self, x):
        return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(self.dt)

    def silu(self, x):
        return (x / (1.0 + np.exp(-x))).astype(self.dt)

    @staticmethod
    def softmax(s):
        s = s - s.max(-1, keepdims=True)
        e = np.exp(s)
        return e / e.sum(-1, keepdims=True)

    # ---- RoPE: HF:92-156 (inv_freq) + HF:196-241 (interleaved apply) -----
    def rope_tables(self, positions: np.ndarray):
        d = self.d
        dim = int(d.head_dim * d.rope_factor)
        j = np.arange(0, dim, 2, dtype=np.float64)
        inv_freq = (1.0 / (d.rope_theta ** (j / dim))).astype(np.float32)
        ang = positions.astype(np.float32)[:, None] * inv_freq[None, :]  # fp32 like HF
        cos = np.repeat(np.cos(ang), 2, axis=-1).astype(self.dt)
        sin = np.repeat(np.sin(ang), 2, axis=-1).astype(self.dt)
        return cos, sin  # [n_pos, rot]

    def apply_rope(self, x, cos, sin):
        """x: [..., n_pos, head_dim]; pairs (2p, 2p+1) rotated, tail passes."""
        rot = cos.shape[-1]
        xr, xp = x[..., :rot], x[..., rot:]
        x1, x2 = xr[..., 0::2], xr[..., 1::2]
        rh = np.stack((-x2, x1), axis=-1).reshape(xr.shape)
        return np.concatenate([xr * cos + rh * sin, xp], axis=-1).astype(self.dt)

    # ---- frontend: HF:531-534,573-578 -----------------------------------
    def conv1d(self, x, wname, bname, stride):
        """x [C_in, T_in] -> [C_out, T_out], no padding (im2col + matmul)."""
        w = self.w[wname]  # [C_out, C_in, k]
        co, ci, k = w.shape
        t_out = (x.shape[1] - k) // stride + 1
        idx = np.arange(t_out)[:, None] * stride + np.arange(k)[None, :]
        cols = x[:, idx]                      # [ci, t_out, k]
        cols = cols.transpose(1, 0, 2).reshape(t_out, ci * k)
        y = self.mm(cols, w.reshape(co, ci * k).T)  # [t_out, co]
        if bname is not None:
            y = y + self.w[bname]
        return y.T.astype(self.dt)

    def frontend(self, pcm: np.ndarray) -> np.ndarray:
        e = "model.encoder."
        x = np.asarray(pcm, self.dt)[None, :]
        h = np.tanh(self.conv1d(x, e + "conv1.weight", None, 64))
        # GroupNorm(num_groups=1, eps=1e-5): stats over all C x T of the utterance
        mu = h.mean()
        var = ((h - mu) ** 2).mean()
        h = (h - mu) / np.sqrt(var + self.dt(1e-5))
        h = h * self.w[e + "groupnorm.weight"][:, None] + self.w[e + "groupnorm.bias"][:, None]
        h = self.gelu(self.conv1d(h.astype(self.dt), e + "conv2.weight", e + "conv2.bias", 3))
        h = self.gelu(self.conv1d(h, e + "conv3.weight", e + "conv3.bias", 2))
        return h.T.astype(self.dt)  # [T, D]

    # ---- attention -------------------------------------------------------
    def _heads(self, x):  # [n, D] -> [H, n, hd]
        return x.reshape(x.shape[0], self.d.heads, self.d.head_dim).transpose(1, 0, 2)

    def _attend(self, q, k, v, causal_from: Optional[int] = None):
        """q [H,nq,hd], k/v [H,nk,hd]; softmax in fp32+ (HF:185-190)."""
        s = np.einsum("hqd,hkd->hqk", self._rnd(q), self._rnd(k)).astype(self.dt)
        s = s * self.dt(self.d.head_dim ** -0.5)
        if causal_from is not None:
            nq, nk = s.shape[1], s.shape[2]
            qpos = causal_from + np.arange(nq)[:, None]
            mask = np.arange(nk)[None, :] > qpos
            s = np.where(mask[None], -np.inf, s)
        p = self.softmax(s).astype(self.dt)
        o = np.einsum("hqk,hkd->hqd", self._rnd(p), self._rnd(v)).astype(self.dt)
        return o.transpose(1, 0, 2).reshape(o.shape[1], -1), p

    # ---- encoder: HF:366-412, 573-606 -----------------------------------
    def encoder(self, pcm: np.ndarray) -> np.ndarray:
        h = self.frontend(pcm)
        T = h.shape[0]
        cos, sin = self.rope_tables(np.arange(T))
        for l in range(self.d.enc_layers):
            p = f"model.encoder.layers.{l}."
            x = self.layernorm(h, p + "input_layernorm.weight")
            q = self.apply_rope(self._heads(self.linear(x, p + "self_attn.q_proj.weight")), cos, sin)
            k = self.apply_rope(self._heads(self.linear(x, p + "self_attn.k_proj.weight")), cos, sin)
            v = self._heads(self.linear(x, p + "self_attn.v_proj.weight"))
            a, _ = self._attend(q, k, v)
            h = h + self.linear(a, p + "self_attn.o_proj.weight")
            x = self.layernorm(h, p + "post_attention_layernorm.weight")
            x = self.gelu(self.linear(x, p + "mlp.fc1.weight", p + "mlp.fc1.bias"))
            h = h + self.linear(x, p + "mlp.fc2.weight", p + "mlp.fc2.bias")
        return self.layernorm(h, "model.encoder.layer_norm.weight")

    # ---- decoder: HF:415-485, 616-712; merged-graph step-0 branch computes
    #      the cross K/V once (core/moonshine-model.cpp:462-476) ------------
    def cross_kv(self, enc_out: np.ndarray):
        ks, vs = [], []
        for l in range(self.d.dec_layers):
            p = f"model.decoder.layers.{l}.encoder_attn."
            ks.append(self._heads(self.linear(enc_out, p + "k_proj.weight")))
            vs.append(self._heads(self.linear(enc_out, p + "v_proj.weight")))
        return ks, vs  # L x [H, T, hd]

    def new_self_cache(self):
        H, hd = self.d.heads, self.d.head_dim
        return ([np.zeros((H, 0, hd), self.dt) for _ in range(self.d.dec_layers)],
                [np.zeros((H, 0, hd), self.dt) for _ in range(self.d.dec_layers)])

    def decoder_step(self, tokens: Sequence[int], pos0: int, self_cache, cross,
                     want_cross_attn=False):
        """Run len(tokens) positions starting at absolute position pos0 with
        the self-attention cache holding positions [0, pos0).  Returns logits
        [n, V] and appends to the cache (in place)."""
        d = self.d
        ks_c, vs_c = cross
        sk, sv = self_cache
        n = len(tokens)
        h = self.w["model.decoder.embed_tokens.weight"][np.asarray(tokens)]
        cos, sin = self.rope_tables(pos0 + np.arange(n))
        xattn = []
        for l in range(d.dec_layers):
            p = f"model.decoder.layers.{l}."
            x = self.layernorm(h, p + "input_layernorm.weight")
            q = self.apply_rope(self._heads(self.linear(x, p + "self_attn.q_proj.weight")), cos, sin)
            k = self.apply_rope(self._heads(self.linear(x, p + "self_attn.k_proj.weight")), cos, sin)
            v = self._heads(self.linear(x, p + "self_attn.v_proj.weight"))
            sk[l] = np.concatenate([sk[l], k], axis=1)
            sv[l] = np.concatenate([sv[l], v], axis=1)
            a, _ = self._attend(q, sk[l], sv[l], causal_from=pos0)
            h = h + self.linear(a, p + "self_attn.o_proj.weight")
            x = self.layernorm(h, p + "post_attention_layernorm.weight")
            qc = self._heads(self.linear(x, p + "encoder_attn.q_proj.weight"))
            a, pc = self._attend(qc, ks_c[l], vs_c[l])
            if want_cross_attn:
                xattn.append(pc)
            h = h + self.linear(a, p + "encoder_attn.o_proj.weight")
            x = self.layernorm(h, p + "final_layernorm.weight")
            y = self.linear(x, p + "mlp.fc1.weight", p + "mlp.fc1.bias")
            up, gate = y[:, : d.ffn], y[:, d.ffn:]          # HF:84-90: value | gate
            h = h + self.linear(self.silu(gate) * up, p + "mlp.fc2.weight", p + "mlp.fc2.bias")
        h = self.layernorm(h, "model.decoder.norm.weight")
        logits = self.head(h)
        if want_cross_attn:
            return logits, xattn
        return logits

    def head(self, h):
        return self.mm(h, self.w["model.decoder.embed_tokens.weight"].T)  # tied head, HF:843

    # ---- host loop: core/moonshine-model.cpp:347-349,370-371,380-517 ------
    @staticmethod
    def max_len(n_samples: int, max_tokens_per_second: float = 6.5) -> int:
        # float32 arithmetic like the C++ (`audio_duration` is a float)
        dur = np.float32(n_samples) / np.float32(16000.0)
        return int(math.ceil(float(np.float32(dur) * np.float32(max_tokens_per_second))))

    @staticmethod
    def argmax_first(x: np.ndarray) -> int:
        # MoonshineTensorView::argmax, core/ort-utils/moonshine-tensor-view.cpp:222-236
        # (strict '>' => lowest index wins ties) == np.argmax
        return int(np.argmax(x))

    def greedy(self, pcm: np.ndarray, max_tokens_per_second: float = 6.5,
               forced: Optional[Sequence[int]] = None, keep_logits: bool = True):
        """Returns (tokens incl. start token and EOS if produced, logits[steps,V]).
        ``forced``: teacher-force these next-tokens instead of the argmax."""
        enc = self.encoder(pcm)
        cross = self.cross_kv(enc)
        cache = self.new_self_cache()
        tokens = [self.d.bos]
        cur = self.d.bos
        all_logits = []
        for t in range(self.max_len(len(pcm), max_tokens_per_second)):
            lg = self.decoder_step([cur], t, cache, cross)[0]
            if keep_logits:
                all_logits.append(lg)
            nxt = self.argmax_first(lg) if forced is None else int(forced[t])
            tokens.append(nxt)
            if nxt == self.d.eos:
                break
            cur = nxt
        return tokens, (np.stack(all_logits) if all_logits else None), enc


# --------------------------------------------------------------------------
# tokenizer: core/bin-tokenizer/bin-tokenizer.cpp:46-66 (format), :406-425
# --------------------------------------------------------------------------
def load_tokenizer_bin(data: bytes) -> List[bytes]:
    toks, p, n = [], 0, len(data)
    while p < n:
        b0 = data[p]; p += 1
        if b0 == 0:
            toks.append(b""); continue
        if b0 < 128:
            ln = b0
        else:
            if p >= n:
                raise ValueError("truncated tokenizer data")
            ln = data[p] * 128 + b0 - 128; p += 1
        if ln > n - p:
            raise ValueError("truncated tokenizer data")
        toks.append(bytes(data[p: p + ln])); p += ln
    if not toks:
        raise ValueError("no tokens")
    return toks


_WS = b" \t\n\r\f\v"


def tokens_to_text(vocab: List[bytes], tokens: Sequence[int], skip_specials=True) -> bytes:
    out = bytearray()
    for t in tokens:
        if t < 0 or t >= len(vocab):
            raise IndexError(f"token {t} out of range")
        b = vocab[t]
        if len(b) == 0:
            raise ValueError(f"Invalid token {t}")
        if skip_specials and len(b) > 2 and b[:1] == b"<" and b[-1:] == b">":
            continue
        out += b
    s = bytes(out).replace("▁".encode("utf-8"), b" ")
    return s.strip(_WS)


def sanitize_utf8(b: bytes) -> bytes:
    """core/transcriber.cpp:1489-1541 sanitize_text: structural check only
    (lead byte class + continuation bytes 10xxxxxx); bad lead/short tail ->
    '?' and advance ONE byte."""
    out, i, n = bytearray(), 0, len(b)
    cont = lambda k: (b[k] & 0xC0) == 0x80
    while i < n:
        c, rem = b[i], n - i
        if c < 0x80:
            out.append(c); i += 1
        elif (c & 0xE0) == 0xC0:
            if rem < 2 or not cont(i + 1):
                out += b"?"; i += 1
            else:
                out += b[i:i + 2]; i += 2
        elif (c & 0xF0) == 0xE0:
            if rem < 3 or not cont(i + 1) or not cont(i + 2):
                out += b"?"; i += 1
            else:
                out += b[i:i + 3]; i += 3
        elif (c & 0xF8) == 0xF0:
            if rem < 4 or not cont(i + 1) or not cont(i + 2) or not cont(i + 3):
                out += b"?"; i += 1
            else:
                out += b[i:i + 4]; i += 4
        else:
            out += b"?"; i += 1
    return bytes(out)


# --------------------------------------------------------------------------
# VAD bypass (vad_threshold == 0): core/voice-activity-detector.cpp:68-96,
# 124-190.  Audio is consumed in 512-sample hops; every hop is "voice"; the
# sub-hop remainder never reaches the segment; one segment covers the clip.
# --------------------------------------------------------------------------
def vad_bypass_segment_length(n_samples_16k: int, hop: int = 512) -> int:
    return (n_samples_16k // hop) * hop


def resample_audio(x: np.ndarray, src_rate: int, dst_rate: int = 16000) -> np.ndarray:
    """core/resampler.cpp:5-86, float32 arithmetic as in the C++:
    downsample = box average over input indices [floor(i*r), floor((i+1)*r)]
    INCLUSIVE (end clamped to n-1); upsample = linear interpolation, last
    sample held."""
    x = np.asarray(x, np.float32)
    if float(src_rate) == float(dst_rate):
        return x.copy()
    n = len(x)
    f32 = np.float32
    # size_t * float / float: evaluated in float, truncated
    n_out = int(f32(f32(f32(n) * f32(dst_rate)) / f32(src_rate)))
    ratio = f32(src_rate) / f32(dst_rate)
    i = np.arange(n_out, dtype=np.float32)
    if src_rate > dst_rate:
        start = (i * ratio).astype(np.float32).astype(np.int64)
        end = ((i + f32(1)) * ratio).astype(np.float32).astype(np.int64)
        end = np.where(end >= n, n - 1, end)
        out = np.empty(n_out, np.float32)
        for k in range(n_out):
            s = f32(0)
            for j in range(start[k], end[k] + 1):
                s = f32(s + x[j])
            cnt = max(end[k] - start[k] + 1, 0)
            out[k] = f32(s / f32(cnt)) if cnt > 0 else f32(0)
        return out
    pos = (i * ratio).astype(np.float32)
    idx = pos.astype(np.int64)
    frac = (pos - idx.astype(np.float32)).astype(np.float32)
    out = np.empty(n_out, np.float32)
    last = idx >= n - 1
    i0 = np.minimum(idx, n - 1); i1 = np.minimum(idx + 1, n - 1)
    out[:] = (x[i0] + frac * (x[i1] - x[i0])).astype(np.float32)
    out[last] = x[n - 1]
    return out
