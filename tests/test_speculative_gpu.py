"""GPU: verify-then-continue decoding on explicit rows (reference: MoonshineStreamingModel::decode_full with speculative
tokens and the multi-token decoder run of decode_tokens, core/moonshine-streaming-model.cpp:1136-1190, 1192-1397; call site
core/transcriber.cpp:1403-1432).  The contract the reference's own tool checks (core/speculative-mismatch-investigate.cpp):
whatever the draft holds, the result equals the greedy decode id for id (a verified row runs the same fp32 arithmetic as the
single-token step; only the summation order of the in-launch attention scores differs, far below any top-2 margin seen here).
What speculation buys is launches: an accepted draft of m ids costs ceil((m + 1) / 8) decoder launches instead of m + 1."""
import numpy as np
import pytest

from moonshine_b200 import api
from moonshine_b200.arch import ARCHS
from moonshine_b200.weights import synth_audio
from tests.util import memory_files

pytestmark = pytest.mark.gpu

ARCH_ENUM = {"test": api.ModelArch.TEST, "test_streaming": api.ModelArch.TEST_STREAMING, "tiny": api.ModelArch.TINY,
             "tiny_streaming": api.ModelArch.TINY_STREAMING}


def make(arch, options=None):
    opts = {"vad_threshold": "0"}
    opts.update(options or {})
    return api.Transcriber(model_arch=ARCH_ENUM[arch], options=opts, memory_files=memory_files(arch, 0, "scaled"))


def greedy(t, arch, audios):
    d = ARCHS[arch]
    return t.debug_run(audios, d.dim, d.vocab, want_encoder=False, max_tokens=300)[2]


def content(ids, d):
    return [i for i in ids if i not in (d.bos, d.eos)]


def assert_same_ids_up_to_near_ties(t, arch, audios, want, got, label):
    """The verify path and the greedy path run the same fp32 arithmetic in different orders (other tile shapes, other kernels),
    so the two may part ways -- but only AT a position where the model's own top-2 candidates are tied to within rounding
    (the reference has a tool that counts exactly these, core/speculative-mismatch-investigate.cpp).  Past such a position
    the contexts differ and that utterance is no longer comparable."""
    d = ARCHS[arch]
    exact = True
    for u, (w, g) in enumerate(zip(want, got)):
        if w == g:
            continue
        exact = False
        i = next((k for k in range(min(len(w), len(g))) if w[k] != g[k]), None)
        assert i is not None and i >= 1, (label, u, "one list is a strict prefix of the other", w, g)
        forced = np.zeros((len(audios), max(len(x) for x in want) + 1), np.int32)
        for k, x in enumerate(want):
            forced[k, :len(x)] = x
        _, lg, _ = t.debug_run(audios, d.dim, d.vocab, forced=forced, logits_steps=i, want_encoder=False, max_tokens=300)
        row = lg[i - 1, u]
        gap = abs(float(row[w[i]]) - float(row[g[i]])) / float(np.abs(row).max())
        assert gap < 2e-4, (label, u, i, w[i], g[i], gap)  # fp32 reordering noise between the kernels is ~1e-5 of max |logit|
    return exact


@pytest.mark.parametrize("arch", ["test", "test_streaming", "tiny_streaming"])
def test_any_draft_gives_the_greedy_ids(arch):
    d = ARCHS[arch]
    audios = [synth_audio(i, n) for i, n in enumerate([48000, 33000, 64000, 25600 + 7, 80000])]
    t = make(arch)
    want = greedy(t, arch, audios)
    assert all(len(w) >= 4 for w in want)
    rng = np.random.default_rng(7)
    cases = {
        "exact": [content(w, d) for w in want],
        "empty": [[] for _ in want],
        "prefix": [content(w, d)[: len(w) // 2] for w in want],                      # the draft of an earlier, shorter update
        "wrong_first": [[(w[1] + 1) % d.vocab] + content(w, d)[1:] for w in want],
        "wrong_middle": [content(w, d)[:5] + [(w[6] + 3) % d.vocab] + content(w, d)[6:] for w in want],
        "too_long": [content(w, d) + [5, 6, 7, 8, 9, 10, 11, 12, 13] for w in want],
        "random": [list(rng.integers(3, d.vocab, size=len(w))) for w in want],
        "mixed": [content(want[0], d), [], content(want[2], d)[:3], [want[3][1]], list(rng.integers(3, d.vocab, size=30))],
    }
    launches, exact = {}, True
    for name, drafts in cases.items():
        got, launches[name] = t.decode_with_drafts(audios, drafts, max_tokens=300)
        exact &= assert_same_ids_up_to_near_ties(t, arch, audios, want, got, name)
    if not exact:   # a near-tie flipped somewhere: the "exact" draft was (correctly) rejected there, launch counts say nothing
        t.close()
        return
    longest = max(len(w) for w in want) - 1          # ids the longest utterance emits
    assert launches["empty"] >= longest              # no draft: the plain greedy loop, one launch per id
    assert launches["exact"] <= (longest + 1 + 7) // 8 + 2, launches   # the whole draft in ceil((m+1)/8) launches (+ the tail)
    assert launches["exact"] < launches["wrong_middle"] <= longest + 2, launches
    t.close()


def test_streaming_updates_with_and_without_draft_verification_print_the_same_lines(monkeypatch):
    """A growing stream: every update re-decodes the open segment; with use_speculative_decoding (the default) the ids of
    the previous update are the draft.  Verifying the draft (8 positions per launch) must print what the plain greedy
    launches print under the same budgets (MOONSHINE_B200_SPEC_VERIFY=0), and the later updates must take fewer launches.
    (use_speculative_decoding=false is not the A/B: the reference budgets that path differently, core/transcriber.cpp:1386
    vs moonshine-streaming-model.cpp:1217.)"""
    arch = "test_streaming"
    audio = synth_audio(3, 16000 * 6)
    texts, steps = {}, {}
    for verify in ("1", "0"):
        monkeypatch.setenv("MOONSHINE_B200_SPEC_VERIFY", verify)
        t = make(arch)
        s = t.create_stream()
        s.start()
        texts[verify], steps[verify] = [], []
        for k in range(0, len(audio), 16000):
            s.add_audio(audio[k:k + 16000])
            tr = s.update_transcription()
            texts[verify].append([l.text for l in tr.lines])
            steps[verify].append(t.last_timings()["decode_steps"])
        s.stop()
        texts[verify].append([l.text for l in s.update_transcription().lines])
        s.close()
        t.close()
    assert texts["1"] == texts["0"]
    assert any(any(x) for x in texts["1"])
    assert sum(steps["1"][1:]) < sum(steps["0"][1:]), (steps["1"], steps["0"])


@pytest.mark.parametrize("arch,rows", [("test_streaming", 8), ("test", 4), ("tiny_streaming", 16)])
def test_multi_token_runs_give_the_single_token_logits(arch, rows):
    """decode_tokens (n > 1): the logits of positions run `rows` at a time must be those of the one-position-per-launch
    teacher-forced loop (the path pinned against the oracle in test_parity_gpu / test_streaming_gpu) and of the oracle."""
    from tests.test_streaming_gpu import soracle
    from tests.util import oracle_for
    d = ARCHS[arch]
    audios = [synth_audio(i, n) for i, n in enumerate([48000, 33000, 64000])]
    t = make(arch)
    ids = greedy(t, arch, audios)
    steps = max(len(x) for x in ids) - 1
    forced = np.zeros((len(audios), steps + 2), np.int32)
    for i, x in enumerate(ids):
        forced[i, :len(x)] = x
    _, single, _ = t.debug_run(audios, d.dim, d.vocab, forced=forced, logits_steps=steps, want_encoder=False, max_tokens=300)
    multi = t.decode_tokens(audios, forced, d.vocab, steps, rows_per_launch=rows)
    for i, x in enumerate(ids):
        for s in range(len(x) - 1):
            ref = single[s, i]
            assert np.abs(multi[s, i] - ref).max() <= 2e-5 * np.abs(ref).max(), (i, s)
    # and against the oracle directly (utterance 0)
    if ARCHS[arch].streaming:
        toks, ref_logits, _ = soracle(arch).transcribe_segment(audios[0], is_final=True)
    else:
        toks, ref_logits, _ = oracle_for(arch, 0, "scaled").greedy(audios[0])
    for s in range(min(len(toks) - 1, steps)):
        if toks[:s + 1] != ids[0][:s + 1]:
            break
        assert np.abs(multi[s, 0] - ref_logits[s]).max() < 1e-3 * np.abs(ref_logits[s]).max(), s
    t.close()
