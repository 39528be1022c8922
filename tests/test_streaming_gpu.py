"""GPU parity for the streaming architectures: the CUDA path (through the C ABI) against the streaming
numpy oracle and the HF-generated fixtures.  Same tolerances as test_parity_gpu.py: logits / memory within
1e-3 relative, greedy ids exact wherever the oracle's own top-2 margin is above the tolerance."""
import glob
import os

import numpy as np
import pytest

from moonshine_b200 import api
from moonshine_b200.arch import ARCHS
from moonshine_b200.weights import synth_audio, synth_tokenizer_bin, synth_weights
from oracle import moonshine_oracle as orc
from oracle.moonshine_streaming_oracle import SDims, StreamingOracle
from tests.util import GOLD, memory_files, rel_err

pytestmark = pytest.mark.gpu

ARCH_ENUM = {"tiny_streaming": api.ModelArch.TINY_STREAMING, "base_streaming": api.ModelArch.BASE_STREAMING,
             "test_streaming": api.ModelArch.TEST_STREAMING, "test_streaming2": api.ModelArch.TEST_STREAMING2}
TOL = 1e-3


def make_transcriber(arch, seed=0, init="scaled", options=None):
    opts = {"vad_threshold": "0"}
    opts.update(options or {})
    return api.Transcriber(model_arch=ARCH_ENUM[arch], options=opts, memory_files=memory_files(arch, seed, init))


def soracle(arch, seed=0, init="scaled"):
    return StreamingOracle(SDims.from_product(ARCHS[arch]), synth_weights(arch, seed, init))


def check_stream_case(arch, seed, init, audios, final=True, only=None):
    """`only`: indices compared with the oracle (the rest of the batch is teacher-forced with
    utterance only[0]'s ids and just has to run) -- keeps full-size batches affordable on the CPU side."""
    d = ARCHS[arch]
    o = soracle(arch, seed, init)
    idx = list(range(len(audios))) if only is None else list(only)
    ref_by = {i: o.transcribe_segment(audios[i], is_final=final) for i in idx}
    refs = [ref_by.get(i) for i in range(len(audios))]
    max_steps = max(len(r[0]) - 1 for r in ref_by.values())
    forced = np.zeros((len(audios), max_steps + 2), np.int32)
    for i in range(len(audios)):
        toks = (refs[i] or ref_by[idx[0]])[0]
        forced[i, :len(toks)] = toks
    t = make_transcriber(arch, seed, init)
    t.debug_stream_partial(not final)
    mems, logits, _ = t.debug_run(audios, d.dim, d.vocab, forced=forced, logits_steps=max_steps, max_tokens=300)
    for i in idx:
        toks, ref_logits, ref_mem = refs[i]
        assert mems[i].shape == ref_mem.shape, f"memory shape utt {i}"
        assert rel_err(mems[i], ref_mem) < TOL, f"memory utt {i}"
        for s in range(len(toks) - 1):
            e = np.abs(logits[s, i] - ref_logits[s]).max() / np.abs(ref_logits[s]).max()
            assert e < TOL, f"logits utt {i} step {s}: {e}"
    _, _, toks_gpu = t.debug_run(audios, d.dim, d.vocab, want_encoder=False, max_tokens=300)
    for i in idx:
        toks, ref_logits, _ = refs[i]
        srt = np.sort(ref_logits, axis=1)
        margin = (srt[:, -1] - srt[:, -2]) / np.abs(ref_logits).max(1)
        got = toks_gpu[i]
        for s in range(len(toks) - 1):
            if margin[s] < 4 * TOL:
                break
            assert got[s + 1] == toks[s + 1], f"utt {i} token {s + 1}"
        else:
            assert got == toks
    t.close()


def test_small_streaming_single():
    check_stream_case("test_streaming", 0, "scaled", [synth_audio(1, 32999)])


def test_small_streaming_ragged_batch_final_and_partial():
    # ragged lengths: one chunk only, partial trailing chunks, lengths that are exact chunk multiples
    audios = [synth_audio(i, n) for i, n in enumerate([32999, 1280 * 5, 1280 * 18 + 1279, 48000, 25600 + 7, 64000])]
    check_stream_case("test_streaming", 0, "scaled", audios, final=True)
    # non-final: 6 look-ahead features held back (needs > 6 features per clip)
    check_stream_case("test_streaming", 0, "scaled", audios[2:], final=False)


def test_small_streaming_tied_head_other_windows():
    audios = [synth_audio(40 + i, 16000 + 3333 * i) for i in range(9)]
    check_stream_case("test_streaming2", 3, "scaled", audios)


def test_tiny_streaming():
    check_stream_case("tiny_streaming", 0, "scaled", [synth_audio(0, 48700), synth_audio(1, 80000)])
    check_stream_case("tiny_streaming", 0, "scaled", [synth_audio(0, 48700)], final=False)


def test_base_streaming():
    check_stream_case("base_streaming", 1, "scaled", [synth_audio(1, 32011), synth_audio(2, 70000)])


def test_config4_base_streaming_batch64_full_size():
    """BASELINE config #4: moonshine-base streaming, batch 64, 10 s clips (500 features, 65 tokens each);
    first / middle / last utterance against the oracle."""
    audios = [synth_audio(i) for i in range(64)]
    check_stream_case("base_streaming", 0, "scaled", audios, only=[0, 31, 63])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "hfs_*.npz"))),
                         ids=lambda p: os.path.basename(p)[4:-4])
def test_streaming_against_hf_golden(path):
    """Directly against fixtures produced by the HF streaming modules (no oracle in the loop)."""
    parts = os.path.basename(path)[4:-4].split("_")
    final, n, inp, seed, init = parts[-1] == "final", int(parts[-2]), parts[-3], int(parts[-4][1:]), parts[-5]
    arch = "_".join(parts[:-5])
    g = np.load(path)
    d = ARCHS[arch]
    audio = synth_audio(int(inp[5:]), n)
    toks = g["tokens"]
    forced = np.zeros((1, len(toks) + 1), np.int32)
    forced[0, :len(toks)] = toks
    t = make_transcriber(arch, seed, init)
    t.debug_stream_partial(not final)
    steps = len(toks) - 1
    mems, logits, _ = t.debug_run([audio], d.dim, d.vocab, forced=forced, logits_steps=steps, max_tokens=300)
    assert tuple(mems[0].shape) == tuple(g["mem_shape"])
    assert np.abs(mems[0][::4] - g["mem_sub"]).max() / g["mem_absmax"] < TOL
    lg = logits[:, 0]
    assert (np.abs(lg[:, ::64] - g["logits_sub"]).max(1) / g["logits_absmax"]).max() < TOL
    top = np.take_along_axis(lg, g["top_idx"], 1)
    assert (np.abs(top - g["top_val"]).max(1) / g["logits_absmax"]).max() < TOL
    clear = g["margin"] / g["logits_absmax"] > 4 * TOL
    assert (lg.argmax(1)[clear] == toks[1:][clear]).all()
    t.close()


def test_streaming_arch_through_the_reference_abi():
    """moonshine_transcribe_without_streaming and the stream calls on a streaming architecture: text per
    update equals the oracle run with the reference's per-segment bookkeeping
    (core/transcriber.cpp:1311-1487), including the speculative-path token budget on re-decodes."""
    arch = "test_streaming"
    d = ARCHS[arch]
    vocab = orc.load_tokenizer_bin(synth_tokenizer_bin(d.vocab))
    o = soracle(arch)
    t = make_transcriber(arch)
    audio = synth_audio(3, 16000 * 3 + 700)
    tr = t.transcribe_without_streaming(audio)
    seg = audio[: len(audio) // 512 * 512]                       # VAD bypass keeps whole 512-sample hops
    toks, _, _ = o.transcribe_segment(seg, is_final=True)
    want = orc.sanitize_utf8(orc.tokens_to_text(vocab, toks)).decode("utf-8")
    assert len(tr.lines) == 1 and tr.lines[0].text == want
    # stream: growing open segment, one forced update per piece
    s = t.create_stream()
    s.start()
    state, fed = {}, np.zeros(0, np.float32)
    for piece in np.array_split(synth_audio(4, 16000 * 4), 5):
        fed = np.concatenate([fed, piece])
        s.add_audio(piece)
        tr = s.update_transcription(api.MOONSHINE_FLAG_FORCE_UPDATE)
        seg = fed[: len(fed) // 512 * 512]
        toks, _, mem = o.transcribe_segment(seg, is_final=False, state=state, speculative=True)
        want = orc.sanitize_utf8(orc.tokens_to_text(vocab, toks)).decode("utf-8") if len(toks) else ""
        assert len(tr.lines) == 1 and not tr.lines[0].is_complete
        assert tr.lines[0].text == want
    s.stop()
    s.close()
    t.close()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "refx_*.npz"))), ids=lambda p: os.path.basename(p)[5:-4])
def test_streaming_against_reference_graph_modules(path):
    """CUDA path against fixtures produced by the reference's own export.py graph modules run chunk by chunk with
    carried state (tests/golden/make_golden_streaming_ref.py) -- no oracle in the loop."""
    parts = os.path.basename(path)[5:-4].split("_")
    n, inp, seed, init = int(parts[-1]), parts[-2], int(parts[-3][1:]), parts[-4]
    arch = "_".join(parts[:-4])
    g = np.load(path)
    d = ARCHS[arch]
    audio = synth_audio(int(inp[5:]), n)
    toks = g["tokens"]
    forced = np.zeros((1, len(toks) + 1), np.int32)
    forced[0, :len(toks)] = toks
    t = make_transcriber(arch, seed, init)
    steps = len(toks) - 1
    mems, logits, _ = t.debug_run([audio], d.dim, d.vocab, forced=forced, logits_steps=steps, max_tokens=300)
    assert tuple(mems[0].shape) == tuple(g["mem_shape"])
    assert np.abs(mems[0][::4] - g["mem_sub"]).max() / g["mem_absmax"] < TOL
    lg = logits[:, 0]
    assert (np.abs(lg[:, ::64] - g["logits_sub"]).max(1) / g["logits_absmax"]).max() < TOL
    top = np.take_along_axis(lg, g["top_idx"], 1)
    assert (np.abs(top - g["top_val"]).max(1) / g["logits_absmax"]).max() < TOL
    clear = g["margin"] / g["logits_absmax"] > 4 * TOL
    assert (lg.argmax(1)[clear] == toks[1:][clear]).all()
    t.close()
