import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moonshine_b200 import api
from moonshine_b200.arch import ARCHS
from moonshine_b200.weights import pack_msw, synth_audio, synth_tokenizer_bin, synth_weights
arch='test'; d=ARCHS[arch]; w=synth_weights(arch,0,'scaled')
mf={"model.msw": pack_msw(arch,w), "tokenizer.bin": synth_tokenizer_bin(d.vocab)}
def mk(): return api.Transcriber(model_arch=api.ModelArch.TEST, options={"vad_threshold":"0"}, memory_files=mf)
a0=synth_audio(0,24000)
t=mk()
print("A first-call text:", repr(t.transcribe_without_streaming(a0).lines[0].text))
print("A debug tokens   :", t.debug_run([a0[:23552]], d.dim, d.vocab, want_encoder=False)[2])
print("A text again     :", repr(t.transcribe_without_streaming(a0).lines[0].text))
print("A B=2 tokens     :", t.debug_run([a0, synth_audio(1,17000)], d.dim, d.vocab, want_encoder=False)[2])
print("A text after B=2 :", repr(t.transcribe_without_streaming(a0).lines[0].text))
t.close()
aud=[synth_audio(i,n) for i,n in enumerate([30000,12345,52000])]
for r in range(6):
    t=mk()
    try:
        print("B", r, [x[:4] for x in t.debug_run(aud, d.dim, d.vocab, want_encoder=False)[2]])
    except Exception as e:
        print("B", r, "FAILED", e); break
    t.close()
