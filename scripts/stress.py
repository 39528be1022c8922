import sys, os, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moonshine_b200 import api
from moonshine_b200.arch import ARCHS
from moonshine_b200.weights import pack_msw, synth_audio, synth_tokenizer_bin, synth_weights
arch=sys.argv[1] if len(sys.argv)>1 else 'test'; d=ARCHS[arch]; w=synth_weights(arch,0,'scaled')
mf={"model.msw": pack_msw(arch,w), "tokenizer.bin": synth_tokenizer_bin(d.vocab)}
AE={'test':api.ModelArch.TEST,'test2':api.ModelArch.TEST2,'tiny':api.ModelArch.TINY}[arch]
def mk(): return api.Transcriber(model_arch=AE, options={"vad_threshold":"0"}, memory_files=mf)
rng=np.random.default_rng(0)
sets=[[synth_audio(0,24000), synth_audio(1,17000)], [synth_audio(0,24000)[:23552]], [synth_audio(i,n) for i,n in enumerate([30000,12345,52000])],
      [synth_audio(10+i, 5000+3111*i) for i in range(7)]]
ref={}
bad=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 40):
    t=mk()
    order=rng.permutation(len(sets))
    for k in order:
        kw=dict(want_encoder=bool(rng.integers(2)), logits_steps=int(rng.integers(0,3)))
        try:
            encs,lg,toks=t.debug_run(sets[k], d.dim, d.vocab, **kw)
        except Exception as e:
            print("iter",it,"set",k,kw,"EXC",e); bad+=1; break
        if k not in ref: ref[k]=toks
        if toks!=ref[k]:
            bad+=1
            nan_enc = [bool(np.isnan(e).any()) for e in encs] if encs else None
            nan_lg = bool(np.isnan(lg).any()) if lg is not None else None
            print("iter",it,"set",k,kw,"MISMATCH",[x[:6] for x in toks],"ref",[x[:6] for x in ref[k]],"nan enc",nan_enc,"nan logits",nan_lg)
    try: t.close()
    except Exception: pass
print("done, bad =",bad)
