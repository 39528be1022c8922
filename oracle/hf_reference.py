"""CPU baseline arm: the Hugging Face float implementation (transformers `MoonshineForConditionalGeneration`)
of the model the reference's shipped ONNX graphs were exported from, loaded with this repo's seeded synthetic
weights and run on ALL host cores.

TEST / BENCHMARK INFRASTRUCTURE ONLY (see the rule in oracle/moonshine_oracle.py): imported by tests/, by
tests/golden/make_golden.py and by bench.py's reference arm -- never by the product.

It is labelled what it is: an architecture-equivalent float CPU baseline, not the shipped int8 ORT graphs (those
files are not in the reference tree; the reference publishes 161 ms per 10 s clip for them on an unnamed
machine, docs/word-level-timestamps.md:201)."""
import math
import time

import numpy as np


def hf_model(dims, weights):
    """transformers/models/moonshine/modeling_moonshine.py with the (arch, seed, init) weights of this repo."""
    import torch
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
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in weights.items()}
    sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in k or "inv_freq" in k for k in missing), missing
    return m


def max_len_for(n_samples, tps=6.5):
    """core/moonshine-model.cpp:347-349 in float32, like the reference."""
    return int(math.ceil(float(np.float32(n_samples) / np.float32(16000.0) * np.float32(tps))))


def greedy_batch(m, pcm_batch, max_len):
    """Greedy decode of a [B, S] batch of equal-length clips with the KV cache; the reference's loop rules
    (start id 1, stop on EOS 2 or after max_len steps, first-max argmax).  Returns a list of id lists."""
    import torch
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(pcm_batch))
        enc_out = m.model.encoder(x)
        B = x.shape[0]
        ids = torch.full((B, 1), 1, dtype=torch.long)
        past = None
        tokens = [[1] for _ in range(B)]
        alive = [True] * B
        for _ in range(max_len):
            out = m(encoder_outputs=enc_out, decoder_input_ids=ids, past_key_values=past, use_cache=True)
            past = out.past_key_values
            nxt = torch.argmax(out.logits[:, -1], dim=-1)
            for b in range(B):
                if alive[b]:
                    tokens[b].append(int(nxt[b]))
                    if int(nxt[b]) == 2:
                        alive[b] = False
            if not any(alive):
                break
            ids = nxt[:, None]
    return tokens


def time_serial(m, audios, threads):
    """Batch 1, one utterance after the other: the reference's only operating point
    (Transcriber::update_transcript_from_segments, core/transcriber.cpp:989-1148)."""
    import torch
    torch.set_num_threads(threads)
    toks = []
    t0 = time.perf_counter()
    for a in audios:
        toks.append(greedy_batch(m, a[None], max_len_for(len(a)))[0])
    return time.perf_counter() - t0, toks


def time_batched(m, audios, threads):
    """The same utterances as ONE batch (what a batching server would do with the float model)."""
    import torch
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    toks = greedy_batch(m, np.stack(audios), max_len_for(len(audios[0])))
    return time.perf_counter() - t0, toks


def pick_threads(m, audios, cores, batch):
    """The fastest intra-op thread count for this host, found on a SHORT sample (2 s of audio, 6 decode steps, `batch`
    clips per call): with 65 launches of ~100 small ops per utterance, more threads than the matrices can feed only adds
    barrier cost (on a 128-core host, 128 threads ran the tiny model 20x slower than 16).  Returns (threads, {t: s})."""
    import torch
    cands = sorted({t for t in (4, 8, 16, 32, 64, cores) if 1 <= t <= cores} | {min(cores, 4)})
    x = np.stack([a[:32000] for a in audios[:batch]])
    seen = {}
    for t in cands:
        torch.set_num_threads(t)
        greedy_batch(m, x, 2)
        t0 = time.perf_counter()
        greedy_batch(m, x, 6)
        seen[t] = time.perf_counter() - t0
        if seen[t] > 4 * min(seen.values()):
            break  # far past the knee: the larger counts only get slower
    best = min(seen, key=seen.get)
    torch.set_num_threads(best)
    return best, seen
