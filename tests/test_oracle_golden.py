"""CPU: the numpy oracle against the golden vectors produced by the HF float
implementation (tests/golden/make_golden.py).  This is the pin that lets the
GPU parity tests trust the oracle."""
import glob
import os

import numpy as np
import pytest

from moonshine_b200.arch import ARCHS
from moonshine_b200.weights import synth_audio, synth_weights
from oracle.moonshine_oracle import Dims, Oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def golden_input(name):
    if name == "beckett":
        return np.load(os.path.join(GOLD, "beckett_pcm16.npy")).astype(np.float32) / np.float32(32768.0)
    if name == "synth0":
        return synth_audio(0)
    if name == "synth1short":
        return synth_audio(1, 48000 + 333)
    raise KeyError(name)


def golden_cases():
    out = []
    for f in sorted(glob.glob(os.path.join(GOLD, "hf_*.npz"))):
        _, arch, init, seed, inp = os.path.basename(f)[:-4].split("_")
        out.append((arch, init, int(seed[1:]), inp, f))
    return out


@pytest.mark.parametrize("arch,init,seed,inp,path", golden_cases(),
                         ids=[os.path.basename(c[4])[3:-4] for c in golden_cases()])
def test_oracle_matches_hf_golden(arch, init, seed, inp, path):
    g = np.load(path)
    o = Oracle(Dims.from_product(ARCHS[arch]), synth_weights(arch, seed, init))
    pcm = golden_input(inp)
    assert len(pcm) == int(g["n_samples"])
    toks, logits, enc = o.greedy(pcm, forced=g["tokens"][1:])
    assert tuple(enc.shape) == tuple(g["enc_shape"])
    # fp32 vs fp32, different summation order only
    assert np.abs(enc[::4] - g["enc_sub"]).max() / g["enc_absmax"] < 2e-5
    rel = np.abs(logits[:, ::64] - g["logits_sub"]).max(1) / g["logits_absmax"]
    assert rel.max() < 2e-5
    top = np.take_along_axis(logits, g["top_idx"], 1)
    assert (np.abs(top - g["top_val"]).max(1) / g["logits_absmax"]).max() < 2e-5
    # greedy ids: identical wherever HF's own top-2 margin is above fp32 noise
    am = logits.argmax(1)
    clear = g["margin"] / g["logits_absmax"] > 1e-4
    assert (am[clear] == g["tokens"][1:][clear]).all()


def test_max_len_rule():
    # core/moonshine-model.cpp:347-349
    assert Oracle.max_len(160000) == 65
    assert Oracle.max_len(159414) == 65
    assert Oracle.max_len(159744) == 65
    assert Oracle.max_len(16000) == 7
    assert Oracle.max_len(48333) == 20
