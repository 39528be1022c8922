"""One transcription of B synthetic 10 s clips (no timing, no checks): the workload ncu / the in-kernel
timeline are pointed at.   python scripts/prof_step.py <tiny|base|base_streaming|tiny_streaming> <B> [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moonshine_b200 import api  # noqa: E402
from moonshine_b200.arch import ARCHS  # noqa: E402
from moonshine_b200.weights import pack_msw, synth_audio, synth_tokenizer_bin, synth_weights  # noqa: E402

arch = sys.argv[1]
B = int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
d = ARCHS[arch]
enum = {"tiny": api.ModelArch.TINY, "base": api.ModelArch.BASE, "tiny_streaming": api.ModelArch.TINY_STREAMING,
        "base_streaming": api.ModelArch.BASE_STREAMING}[arch]
t = api.Transcriber(model_arch=enum, options={"vad_threshold": "0"},
                    memory_files={"model.msw": pack_msw(arch, synth_weights(arch, 0, "hf")),
                                  "tokenizer.bin": synth_tokenizer_bin(d.vocab)})
aud = [synth_audio(i) for i in range(B)]
for _ in range(reps):
    t.debug_run(aud, d.dim, d.vocab, want_encoder=False, max_tokens=300)
t.close()
