import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moonshine_b200 import api
from moonshine_b200.arch import ARCHS
from moonshine_b200.weights import pack_msw, synth_audio, synth_tokenizer_bin, synth_weights
arch=sys.argv[1]; B=int(sys.argv[2]); d=ARCHS[arch]
w=synth_weights(arch,0,'hf')
t=api.Transcriber(model_arch={'tiny':api.ModelArch.TINY,'base':api.ModelArch.BASE}[arch], options={"vad_threshold":"0"}, memory_files={"model.msw": pack_msw(arch,w), "tokenizer.bin": synth_tokenizer_bin(d.vocab)})
aud=[synth_audio(i) for i in range(B)]
t.debug_run(aud, d.dim, d.vocab, want_encoder=False)
