import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moonshine_b200 import api
from moonshine_b200.arch import ARCHS
from moonshine_b200.weights import pack_msw, synth_audio, synth_tokenizer_bin, synth_weights
arch='test'; d=ARCHS[arch]; w=synth_weights(arch,0,'scaled')
mf={"model.msw": pack_msw(arch,w), "tokenizer.bin": synth_tokenizer_bin(d.vocab)}
def mk(): return api.Transcriber(model_arch=api.ModelArch.TEST, options={"vad_threshold":"0"}, memory_files=mf)
audio=[synth_audio(0,24000), synth_audio(1,17000)]
for name, kw in [("enc+logits4", dict(logits_steps=4)), ("enc only", dict()), ("logits4 only", dict(logits_steps=4, want_encoder=False)), ("plain", dict(want_encoder=False))]:
    t=mk()
    try:
        r=t.debug_run(audio, d.dim, d.vocab, **kw)
        print(name, "tokens", [x[:5] for x in r[2]])
        tr=t.transcribe_without_streaming(audio[0])
        print(name, "-> text", repr(tr.lines[0].text))
        print(name, "-> dbg ", t.debug_run([audio[0][:23552]], d.dim, d.vocab, want_encoder=False)[2])
    except Exception as e:
        print(name, "FAILED", e)
    try: t.close()
    except Exception: pass
