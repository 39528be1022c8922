"""CPU: the streaming numpy oracle against golden vectors produced by the HF streaming modules
(tests/golden/make_golden_streaming.py), and its two memory computations against each other."""
import glob
import os

import numpy as np
import pytest

from moonshine_b200.arch import ARCHS, streaming_lengths
from moonshine_b200.weights import synth_audio, synth_weights
from oracle.moonshine_streaming_oracle import ChunkedState, SDims, StreamingOracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def golden_cases():
    out = []
    for f in sorted(glob.glob(os.path.join(GOLD, "hfs_*.npz"))):
        parts = os.path.basename(f)[4:-4].split("_")
        # <arch...>_<init>_s<seed>_<input>_<n>_<final|partial>
        final, n, inp, seed, init = parts[-1], int(parts[-2]), parts[-3], int(parts[-4][1:]), parts[-5]
        arch = "_".join(parts[:-5])
        out.append((arch, init, seed, inp, n, final == "final", f))
    return out


def streaming_oracle(arch, seed=0, init="scaled", **kw):
    return StreamingOracle(SDims.from_product(ARCHS[arch]), synth_weights(arch, seed, init), **kw)


@pytest.mark.parametrize("arch,init,seed,inp,n,final,path", golden_cases(),
                         ids=[os.path.basename(c[6])[4:-4] for c in golden_cases()])
def test_streaming_oracle_matches_hf_golden(arch, init, seed, inp, n, final, path):
    g = np.load(path)
    o = streaming_oracle(arch, seed, init)
    pcm = synth_audio(int(inp[5:]), n)
    toks, logits, mem = o.transcribe_segment(pcm, is_final=final, forced=g["tokens"][1:])
    assert tuple(mem.shape) == tuple(g["mem_shape"]) and mem.shape[0] == int(g["emitted"])
    assert np.abs(mem[::4] - g["mem_sub"]).max() / g["mem_absmax"] < 2e-5
    rel = np.abs(logits[:, ::64] - g["logits_sub"]).max(1) / g["logits_absmax"]
    assert rel.max() < 2e-5
    top = np.take_along_axis(logits, g["top_idx"], 1)
    assert (np.abs(top - g["top_val"]).max(1) / g["logits_absmax"]).max() < 2e-5
    am = logits.argmax(1)
    clear = g["margin"] / g["logits_absmax"] > 1e-4
    assert (am[clear] == g["tokens"][1:][clear]).all()


@pytest.mark.parametrize("arch", ["test_streaming", "test_streaming2"])
def test_one_pass_memory_equals_the_reference_state_machine(arch):
    """The CUDA path encodes a segment in one pass; the reference feeds 1280-sample chunks through carried
    frontend state and re-encodes a sliding window with look-ahead hold-back
    (moonshine-streaming-model.cpp:604-772).  Both orders must give the same memory, for any update pattern."""
    o = streaming_oracle(arch, 0, dtype=np.float64)
    pcm = synth_audio(5, 16000 * 4 + 123)
    rng = np.random.default_rng(0)
    st = ChunkedState(o)
    fed, emitted_hist = 0, []
    total_chunks = len(pcm) // 1280
    while fed < total_chunks:
        k = int(rng.integers(1, 6))
        for c in range(fed, min(total_chunks, fed + k)):
            st.process_audio_chunk(pcm[c * 1280:(c + 1) * 1280])
        fed = min(total_chunks, fed + k)
        st.encode(is_final=False)
        emitted_hist.append(st.emitted)
        ref = o.memory_stateless(pcm, fed * 4, st.emitted)
        assert np.abs(ref - st.memory).max() < 1e-11
    st.encode(is_final=True)
    assert st.emitted == total_chunks * 4
    ref = o.memory_stateless(pcm, total_chunks * 4, total_chunks * 4)
    assert np.abs(ref - st.memory).max() < 1e-11


def test_streaming_length_bookkeeping():
    # core/transcriber.cpp:1331-1372: whole 1280-sample chunks only, 16 held back unless final
    assert streaming_lengths(160000) == (160000, 500, 500)
    assert streaming_lengths(160000, is_final=False) == (160000, 500, 484)
    assert streaming_lengths(1279) == (0, 0, 0)
    assert streaming_lengths(16000 * 3 + 700) == (48640, 152, 152)
    # incremental: a second non-final call only adds the new whole chunks
    p, n, e = streaming_lengths(5000, is_final=False)
    assert (p, n, e) == (3840, 12, 0)
    p, n, e = streaming_lengths(30000, emitted_before=e, processed_before=p, is_final=False)
    assert (p, n, e) == (29440, 92, 76)
    # final call with no new audio at all: the encoder is not run, held-back features stay out
    assert streaming_lengths(29440, emitted_before=76, processed_before=29440, is_final=True) == (29440, 92, 76)
    # final call with a partial chunk of new audio: everything analysed so far is emitted
    assert streaming_lengths(29500, emitted_before=76, processed_before=29440, is_final=True) == (29440, 92, 92)
    o = StreamingOracle
    assert o.max_tokens_greedy(160000) == 65 and o.max_tokens_greedy(16000 * 60) == 256


def refx_cases():
    out = []
    for f in sorted(glob.glob(os.path.join(GOLD, "refx_*.npz"))):
        parts = os.path.basename(f)[5:-4].split("_")     # <arch...>_<init>_s<seed>_<input>_<n>
        n, inp, seed, init = int(parts[-1]), parts[-2], int(parts[-3][1:]), parts[-4]
        out.append(("_".join(parts[:-4]), init, seed, inp, n, f))
    return out


@pytest.mark.parametrize("arch,init,seed,inp,n,path", refx_cases(), ids=[os.path.basename(c[5])[5:-4] for c in refx_cases()])
def test_streaming_oracle_matches_reference_graph_modules_run_chunk_by_chunk(arch, init, seed, inp, n, path):
    """Fixtures from the REFERENCE's own export.py modules (Frontend / Encoder / Adapter / CrossKV / DecoderKV),
    driven chunk by chunk with carried state over several non-final updates and a final one
    (tests/golden/make_golden_streaming_ref.py).  The oracle's ONE stateless pass over the analysed audio must give
    the same memory, logits and ids: "stateless == chunked" pinned to reference code."""
    g = np.load(path)
    o = streaming_oracle(arch, seed, init)
    pcm = synth_audio(int(inp[5:]), n)
    toks, logits, mem = o.transcribe_segment(pcm, is_final=True, forced=g["tokens"][1:])
    assert tuple(mem.shape) == tuple(g["mem_shape"]) and mem.shape[0] == int(g["mem_len_after_update"][-1])
    assert np.abs(mem[::4] - g["mem_sub"]).max() / g["mem_absmax"] < 2e-5
    assert (np.abs(logits[:, ::64] - g["logits_sub"]).max(1) / g["logits_absmax"]).max() < 2e-5
    top = np.take_along_axis(logits, g["top_idx"], 1)
    assert (np.abs(top - g["top_val"]).max(1) / g["logits_absmax"]).max() < 2e-5
    clear = g["margin"] / g["logits_absmax"] > 1e-4
    assert (logits.argmax(1)[clear] == g["tokens"][1:][clear]).all()
    # the memory length after every non-final update follows the same bookkeeping as the product's host code
    d = ARCHS[arch]
    processed = emitted = 0
    for end, want in zip(g["pieces"], g["mem_len_after_update"]):
        processed, _, emitted = streaming_lengths(int(end), emitted, processed, is_final=(int(end) == n),
                                                  lookahead=d.total_lookahead)
        assert emitted == int(want)
    # one multi-token decoder call (decode_tokens, moonshine-streaming-model.cpp:1136-1190) == the stepped logits
    k = g["verify_logits_sub"].shape[0]
    assert (np.abs(logits[:k, ::64] - g["verify_logits_sub"]).max(1) / g["verify_absmax"]).max() < 2e-5
