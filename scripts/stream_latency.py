#!/usr/bin/env python
"""Per-update latency of ONE growing stream (the reference's live-caption operating point): a 10 s clip arrives in 0.5 s
pieces; every update re-encodes and re-decodes the open segment.  Compared: verifying the previous update's ids as a draft
(8 positions per decoder launch, the default) against plain greedy launches under the same budgets
(MOONSHINE_B200_SPEC_VERIFY=0).  Prints one JSON line per model.

    python scripts/stream_latency.py [tiny_streaming base_streaming]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moonshine_b200 import api  # noqa: E402
from moonshine_b200.arch import ARCHS  # noqa: E402
from moonshine_b200.weights import pack_msw, synth_audio, synth_tokenizer_bin, synth_weights  # noqa: E402

ENUM = {"tiny_streaming": api.ModelArch.TINY_STREAMING, "base_streaming": api.ModelArch.BASE_STREAMING}


def run(arch, verify):
    os.environ["MOONSHINE_B200_SPEC_VERIFY"] = "1" if verify else "0"
    d = ARCHS[arch]
    t = api.Transcriber(model_arch=ENUM[arch], options={"vad_threshold": "0"},
                        memory_files={"model.msw": pack_msw(arch, synth_weights(arch, 0, "hf")),
                                      "tokenizer.bin": synth_tokenizer_bin(d.vocab)})
    audio = synth_audio(0, 160000)
    lat, steps, texts = [], [], []
    for rep in range(2):  # first pass warms allocations
        s = t.create_stream()
        s.start()
        lat, steps, texts = [], [], []
        for k in range(0, len(audio), 8000):
            s.add_audio(audio[k:k + 8000])
            t0 = time.perf_counter()
            tr = s.update_transcription()
            lat.append(1000.0 * (time.perf_counter() - t0))
            steps.append(int(t.last_timings()["decode_steps"]))
            texts.append([l.text for l in tr.lines])
        s.stop()
        s.close()
    t.close()
    return lat, steps, texts


def redecode(arch):
    """Upper bound of what verification buys (what a trained model's stable transcript gives): re-decode of the full 10 s
    clip with the exact previous ids as the draft against no draft; whole call (encoder included), median of 5."""
    d = ARCHS[arch]
    t = api.Transcriber(model_arch=ENUM[arch], options={"vad_threshold": "0"},
                        memory_files={"model.msw": pack_msw(arch, synth_weights(arch, 0, "hf")),
                                      "tokenizer.bin": synth_tokenizer_bin(d.vocab)})
    audio = [synth_audio(0, 160000)]
    ids, _ = t.decode_with_drafts(audio, [[]], max_tokens=300)
    draft = [[i for i in ids[0] if i not in (d.bos, d.eos)]]
    out = {}
    for name, dr in (("exact_draft", draft), ("no_draft", [[]])):
        ms = []
        for _ in range(6):
            t0 = time.perf_counter()
            got, launches = t.decode_with_drafts(audio, dr, max_tokens=300)
            ms.append(1000.0 * (time.perf_counter() - t0))
            assert got == ids
        out[name] = {"ms_median": sorted(ms[1:])[2], "decoder_launches": launches}
    t.close()
    return out


for arch in (sys.argv[1:] or ["tiny_streaming", "base_streaming"]):
    print(json.dumps({"model": arch, "redecode_10s_clip": redecode(arch)}))
    la, sa, ta = run(arch, True)
    lb, sb, tb = run(arch, False)
    print(json.dumps({"model": arch, "updates": len(la), "audio_s_per_update": 0.5, "same_text": ta == tb,
                      "verify": {"ms_per_update_mean": sum(la) / len(la), "ms_last_update": la[-1], "decoder_launches": sa},
                      "greedy": {"ms_per_update_mean": sum(lb) / len(lb), "ms_last_update": lb[-1], "decoder_launches": sb}}))
