"""GPU parity: the CUDA path (through the C ABI) against the numpy oracle and
the committed HF golden vectors.

Tolerances (north_star): greedy token ids bit-exact; logits within 1e-3
relative (max |delta| / max |logit| per step); encoder output within 1e-3.
Token ids are compared exactly wherever the oracle's own top-2 margin exceeds
the logit tolerance (a near-tie can legitimately flip under any reordering of
fp32 sums; the reference itself states its decoding "is not bit-stable across
process states", core/transcriber-test.cpp:1257-1262).
"""
import glob
import os

import numpy as np
import pytest

from moonshine_b200 import api
from moonshine_b200.arch import ARCHS
from moonshine_b200.weights import synth_audio
from oracle import moonshine_oracle as orc
from tests.util import GOLD, beckett, memory_files, oracle_for, rel_err

pytestmark = pytest.mark.gpu

ARCH_ENUM = {"tiny": api.ModelArch.TINY, "base": api.ModelArch.BASE,
             "test": api.ModelArch.TEST, "test2": api.ModelArch.TEST2}
LOGIT_TOL = 1e-3
ENC_TOL = 1e-3


def make_transcriber(arch, seed=0, init="scaled", options=None, tokenizer=None):
    opts = {"vad_threshold": "0"}
    opts.update(options or {})
    return api.Transcriber(model_arch=ARCH_ENUM[arch], options=opts,
                           memory_files=memory_files(arch, seed, init, tokenizer))


def check_case(arch, seed, init, audios, logit_tol=LOGIT_TOL, only=None, logits_steps=None, stats=None):
    """`only`: indices compared with the oracle (the rest of the batch is teacher-forced with utterance only[0]'s ids
    and just has to run) -- keeps full-size batches affordable on the CPU side.  `logits_steps`: compare the logits
    of the first N steps only (a [steps, B, V] dump of 256 utterances x 65 steps would be 2 GB).
    `stats`: dict that receives how many decode steps were compared id-for-id and how many were skipped as near-ties."""
    d = ARCHS[arch]
    o = oracle_for(arch, seed, init)
    idx = list(range(len(audios))) if only is None else list(only)
    ref_by = {i: o.greedy(audios[i]) for i in idx}
    refs = [ref_by.get(i) for i in range(len(audios))]
    max_steps = max(len(r[0]) - 1 for r in ref_by.values())
    forced = np.zeros((len(audios), max_steps + 2), np.int32)
    for i in range(len(audios)):
        toks = (refs[i] or ref_by[idx[0]])[0]
        forced[i, :len(toks)] = toks
    t = make_transcriber(arch, seed, init)
    # (1) teacher-forced: logits at every step + encoder output
    n_lg = max_steps if logits_steps is None else min(logits_steps, max_steps)
    encs, logits, _ = t.debug_run(audios, d.dim, d.vocab, forced=forced, logits_steps=n_lg)
    for i in idx:
        toks, ref_logits, ref_enc = refs[i]
        assert encs[i].shape == ref_enc.shape
        assert rel_err(encs[i], ref_enc) < ENC_TOL, f"encoder utt {i}"
        n = min(len(toks) - 1, n_lg)
        for s in range(n):
            e = np.abs(logits[s, i] - ref_logits[s]).max() / np.abs(ref_logits[s]).max()
            assert e < logit_tol, f"logits utt {i} step {s}: {e}"
    # (2) free-running greedy ids
    _, _, toks_gpu = t.debug_run(audios, d.dim, d.vocab, want_encoder=False)
    # ids must agree step for step; the two may part ways only AT a step whose top-2 margin in the oracle is inside
    # the logit tolerance (a near-tie the fp32 orderings may legitimately resolve differently) -- after that the
    # contexts differ and the rest of that utterance is not comparable.
    compared = skipped = 0
    for i in idx:
        toks, ref_logits, _ = refs[i]
        srt = np.sort(ref_logits, axis=1)
        margin = (srt[:, -1] - srt[:, -2]) / np.abs(ref_logits).max(1)
        got = toks_gpu[i]
        for s in range(len(toks) - 1):
            if s + 1 < len(got) and got[s + 1] == toks[s + 1]:
                compared += 1
                continue
            assert margin[s] < 4 * logit_tol, f"utt {i} token {s + 1}: {got[s + 1] if s + 1 < len(got) else None} != {toks[s + 1]} (margin {margin[s]:.2e})"
            skipped += len(toks) - 1 - s
            break
        else:
            assert got == toks
    if stats is not None:
        stats["compared"], stats["skipped"] = compared, skipped
    t.close()
    return refs


def test_config1_tiny_batch32_headline_full_size():
    """BASELINE configs[1]: moonshine-tiny, batch 32, 10 s clips; first / middle / last utterance against the oracle."""
    st = {}
    check_case("tiny", 0, "scaled", [synth_audio(i) for i in range(32)], only=[0, 15, 31], logits_steps=8, stats=st)
    assert st["compared"] >= 0.5 * (st["compared"] + st["skipped"]), st   # most steps are compared id-for-id


def test_config2_base_batch256_full_size():
    """BASELINE configs[2]: moonshine-base, batch 256, 10 s clips (the weight-stationary kernel with 16-utterance
    attention tiles and 48-utterance GEMM groups); first / middle / last utterance against the oracle."""
    check_case("base", 0, "scaled", [synth_audio(i) for i in range(256)], only=[0, 127, 255], logits_steps=6)


def test_argmax_tie_takes_the_lowest_index():
    """MoonshineTensorView::argmax keeps the FIRST maximum (strict >, core/ort-utils/moonshine-tensor-view.cpp:222-236).
    Three vocabulary rows share one embedding, so their logits are bit-identical at every step; whenever that
    logit wins, the smallest id must be emitted -- across vocab chunks (different CTAs) and inside one chunk."""
    from moonshine_b200.weights import pack_msw, synth_tokenizer_bin, synth_weights
    arch = "test"
    d = ARCHS[arch]
    w = {k: v.copy() for k, v in synth_weights(arch, 0, "scaled").items()}
    o0 = oracle_for(arch, 0, "scaled")
    audio = synth_audio(5, 40000)
    toks0, lg0, _ = o0.greedy(audio)
    win = int(toks0[1])                                  # the id the unmodified model emits first
    lo, mid, hi = 7, 8, d.vocab - 5                      # same chunk pair (7, 8) and a far chunk
    emb = w["model.decoder.embed_tokens.weight"]
    for j in (lo, mid, hi):
        emb[j] = emb[win]
    o = orc.Oracle(orc.Dims.from_product(d), w)
    toks, lg, _ = o.greedy(audio)
    assert lg[0][lo] == lg[0][mid] == lg[0][hi] == lg[0][win]
    assert toks[1] == min(lo, win)                       # numpy argmax: first maximum
    t = api.Transcriber(model_arch=ARCH_ENUM[arch], options={"vad_threshold": "0"},
                        memory_files={"model.msw": pack_msw(arch, w), "tokenizer.bin": synth_tokenizer_bin(d.vocab)})
    _, lgg, got = t.debug_run([audio], d.dim, d.vocab, want_encoder=False, logits_steps=2)
    assert lgg[0, 0][lo] == lgg[0, 0][mid] == lgg[0, 0][hi]     # bit-identical on the device too
    assert got[0][1] == min(lo, win)
    t.close()


def test_small_arch_single():
    check_case("test", 0, "scaled", [synth_audio(1, 48333)])


def test_small_arch_ragged_batch():
    # ragged lengths incl. the shortest the frontend accepts and an odd length
    audios = [synth_audio(i, n) for i, n in enumerate([48333, 16000, 2500, 31999, 80000, 1151])]
    check_case("test", 0, "scaled", audios)


def test_small_arch2_batch_many():
    # head_dim 36 (like tiny), 3 decoder layers, vocab not a multiple of 32, B > 16
    audios = [synth_audio(100 + i, 9000 + 777 * i) for i in range(19)]
    check_case("test2", 3, "scaled", audios)


@pytest.mark.parametrize("B", [100, 130], ids=["B100_tile8", "B130_tile16"])
def test_small_arch_large_batch_tensor_core_tiles(B):
    """B >= 97 switches the decoder's layer GEMMs to the mma.sync row tiles (8 / 16
    utterances per work item); ragged short clips keep the oracle fast."""
    audios = [synth_audio(500 + i, 2500 + 137 * (i % 23) + 16 * i) for i in range(B)]
    check_case("test", 0, "scaled", audios)


def test_tiny_beckett_and_synth():
    check_case("tiny", 0, "scaled", [beckett(), synth_audio(0)])


def test_tiny_hf_init():
    check_case("tiny", 0, "hf", [synth_audio(0)])


def test_base_synth():
    check_case("base", 0, "scaled", [synth_audio(0), synth_audio(1, 48333)])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "hf_*.npz"))),
                         ids=lambda p: os.path.basename(p)[3:-4])
def test_against_hf_golden(path):
    """Directly against the HF-generated fixtures (no oracle in the loop)."""
    _, arch, init, seed, inp = os.path.basename(path)[:-4].split("_")
    seed = int(seed[1:])
    g = np.load(path)
    d = ARCHS[arch]
    audio = {"beckett": beckett(), "synth0": synth_audio(0), "synth1short": synth_audio(1, 48333)}[inp]
    toks = g["tokens"]
    forced = np.zeros((1, len(toks) + 1), np.int32)
    forced[0, :len(toks)] = toks
    t = make_transcriber(arch, seed, init)
    encs, logits, _ = t.debug_run([audio], d.dim, d.vocab, forced=forced, logits_steps=len(toks) - 1)
    assert tuple(encs[0].shape) == tuple(g["enc_shape"])
    assert np.abs(encs[0][::4] - g["enc_sub"]).max() / g["enc_absmax"] < ENC_TOL
    lg = logits[:, 0, :]
    assert (np.abs(lg[:, ::64] - g["logits_sub"]).max(1) / g["logits_absmax"]).max() < LOGIT_TOL
    top = np.take_along_axis(lg, g["top_idx"], 1)
    assert (np.abs(top - g["top_val"]).max(1) / g["logits_absmax"]).max() < LOGIT_TOL
    clear = g["margin"] / g["logits_absmax"] > 4 * LOGIT_TOL
    assert (lg.argmax(1)[clear] == toks[1:][clear]).all()
    t.close()


def test_abi_transcribe_matches_oracle_text():
    """moonshine_transcribe_without_streaming end to end: hop-truncated segment
    (voice-activity-detector.cpp:68-96), greedy ids, detokenised + sanitised text."""
    tokenizer_path = os.path.join(GOLD, "..", "..", "tests", "golden", "tokenizer_tiny_en.bin")
    audio = beckett()
    o = oracle_for("tiny", 0, "scaled")
    seg = audio[: orc.vad_bypass_segment_length(len(audio))]
    toks, ref_logits, _ = o.greedy(seg)
    from moonshine_b200.weights import synth_tokenizer_bin
    vocab = orc.load_tokenizer_bin(synth_tokenizer_bin(32768))
    want = orc.sanitize_utf8(orc.tokens_to_text(vocab, toks)).decode("utf-8")
    t = make_transcriber("tiny", 0, "scaled")
    tr = t.transcribe_without_streaming(audio)
    assert len(tr.lines) == 1
    line = tr.lines[0]
    assert line.is_complete and line.is_new and line.is_updated
    assert line.audio_data.size == len(seg)
    assert line.start_time < 1e-3 and abs(line.duration - len(seg) / 16000.0) < 1e-3
    srt = np.sort(ref_logits, axis=1)
    margin = ((srt[:, -1] - srt[:, -2]) / np.abs(ref_logits).max(1)).min()
    if margin > 4 * LOGIT_TOL:
        assert line.text == want
    # batch entry == N single calls
    audios = [audio, synth_audio(3, 40000), synth_audio(4, 70001)]
    batch = t.transcribe_batch_without_streaming(audios)
    singles = [t.transcribe_without_streaming(a) for a in audios]
    for b, s in zip(batch, singles):
        assert [l.text for l in b.lines] == [l.text for l in s.lines]
        assert [l.audio_data.size for l in b.lines] == [l.audio_data.size for l in s.lines]
    t.close()


def test_device_entry_matches_host_entry():
    import torch
    d = ARCHS["test"]
    audios = [synth_audio(i, n) for i, n in enumerate([30000, 12345, 52000])]
    t = make_transcriber("test", 0, "scaled")
    _, _, want = t.debug_run(audios, d.dim, d.vocab, want_encoder=False)
    stride = 52000
    x = torch.zeros(3, stride, device="cuda")
    for i, a in enumerate(audios):
        x[i, :len(a)] = torch.from_numpy(a).cuda()
    torch.cuda.synchronize()
    got = t.transcribe_device(x.data_ptr(), stride, [len(a) for a in audios])
    assert got == want
    t.close()


def test_too_short_audio_is_an_error_not_a_crash():
    t = make_transcriber("test", 0, "scaled")
    with pytest.raises(api.MoonshineError):
        t.debug_run([np.zeros(600, np.float32)], 64, 512)
    # the transcriber stays usable
    d = ARCHS["test"]
    t.debug_run([synth_audio(0, 20000)], d.dim, d.vocab)
    t.close()


@pytest.mark.parametrize("impl,tol", [(1, 1e-5), (2, 6e-5)], ids=["simt_fp32", "tcgen05_bf16x3"])
def test_gemm_kernel_against_torch(impl, tol):
    """Both dense kernels against a float64 torch reference (bias + exact GELU
    epilogue, ragged M/N/K incl. K % 32 != 0 and K % 4 != 0 handled by padding)."""
    import torch
    lib = api.load_library()
    torch.manual_seed(0)
    for (M, N, K) in [(300, 200, 52), (129, 257, 36), (1000, 576, 2016), (415, 415, 416), (128, 128, 32),
                      (77, 130, 288), (2049, 96, 1152)]:
        A = torch.randn(M, K, device="cuda")
        W = torch.randn(N, K, device="cuda")
        bias = torch.randn(N, device="cuda")
        C = torch.zeros(M, N, device="cuda")
        rc = lib.moonshine_b200_test_gemm(A.data_ptr(), W.data_ptr(), C.data_ptr(), M, N, K, K, K, N,
                                          bias.data_ptr(), 1, 0, impl)
        assert rc == 0
        ref = torch.nn.functional.gelu(A.double() @ W.double().T + bias.double()).float()
        err = (C - ref).abs().max().item() / ref.abs().max().item()
        assert err < tol, (M, N, K, err)
        # accumulate mode, no activation
        C2 = torch.ones(M, N, device="cuda")
        rc = lib.moonshine_b200_test_gemm(A.data_ptr(), W.data_ptr(), C2.data_ptr(), M, N, K, K, K, N,
                                          0, 0, 1, impl)
        assert rc == 0
        ref2 = (1.0 + A.double() @ W.double().T).float()
        assert (C2 - ref2).abs().max().item() / ref2.abs().max().item() < tol


@pytest.mark.parametrize("impl", [3, 4, 5, 6], ids=["tiles", "tiles_planes_out", "persistent", "persistent_planes_out"])
def test_plane_fed_gemm_against_torch(impl):
    """gemm_planes.cu (both operands pre-split to bf16 hi/lo plane tiles, bulk-copy fed tcgen05) against float64: bias +
    exact GELU, accumulate, ragged M / N, both CTA mappings (one tile per CTA / persistent macro tiles).  impl 4 / 6 also write the result as plane tiles (the fc1 -> fc2 hand-over) and
    reads it back through a second plane-fed product with the identity: hi + lo carries ~16 mantissa bits."""
    import torch
    lib = api.load_library()
    torch.manual_seed(0)
    shapes = [(300, 200, 64), (129, 256, 32), (1000, 576, 288), (415, 416, 416), (128, 128, 32), (77, 1152, 288), (2049, 288, 1152)]
    for (M, N, K) in shapes:
        if impl in (4, 6) and N % 32:
            continue
        A = torch.randn(M, K, device="cuda")
        W = torch.randn(N, K, device="cuda")
        bias = torch.randn(N, device="cuda")
        C = torch.zeros(M, N, device="cuda")
        rc = lib.moonshine_b200_test_gemm(A.data_ptr(), W.data_ptr(), C.data_ptr(), M, N, K, K, K, N, bias.data_ptr(), 1, 0, impl)
        assert rc == 0
        ref = torch.nn.functional.gelu(A.double() @ W.double().T + bias.double()).float()
        err = (C - ref).abs().max().item() / ref.abs().max().item()
        assert err < (6e-5 if impl in (3, 5) else 1e-4), (M, N, K, err)
        if impl in (3, 5):
            C2 = torch.ones(M, N, device="cuda")
            rc = lib.moonshine_b200_test_gemm(A.data_ptr(), W.data_ptr(), C2.data_ptr(), M, N, K, K, K, N, 0, 0, 1, impl)
            assert rc == 0
            ref2 = (1.0 + A.double() @ W.double().T).float()
            assert (C2 - ref2).abs().max().item() / ref2.abs().max().item() < 6e-5
