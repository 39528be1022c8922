"""CPU: safetensors checkpoint -> .msw conversion (real-weight ingestion path)."""
import os

import numpy as np
import pytest

from moonshine_b200.convert import convert, hf_to_msw_tensors, read_safetensors, write_safetensors
from moonshine_b200.weights import read_msw, synth_weights


def test_safetensors_roundtrip_to_msw(tmp_path):
    w = synth_weights("test", 2, "scaled")
    ck = dict(w)
    ck["proj_out.weight"] = w["model.decoder.embed_tokens.weight"]           # tied head as HF stores it
    ck["model.encoder.rotary_emb.inv_freq"] = np.ones(7, np.float32)          # ignored buffer
    st = tmp_path / "model.safetensors"
    write_safetensors(str(st), ck)
    back = read_safetensors(str(st))
    assert set(back) == set(ck)
    out = convert(str(tmp_path), str(tmp_path / "model_dir"), "test")
    arch, got = read_msw(os.path.join(out, "model.msw"))
    assert arch == 100
    assert set(got) == set(w)
    for k in w:
        np.testing.assert_array_equal(got[k], w[k])


def test_wrong_arch_is_reported():
    w = synth_weights("test", 0, "hf")
    with pytest.raises(ValueError):
        hf_to_msw_tensors("test2", w)


def test_bf16_and_f16_tensors_are_widened(tmp_path):
    import json, struct
    a = np.array([1.0, -2.5, 0.15625, 3.0], np.float32)
    bf = (a.view(np.uint32) >> 16).astype("<u2").tobytes()
    f16 = a.astype("<f2").tobytes()
    hdr = {"x": {"dtype": "BF16", "shape": [4], "data_offsets": [0, 8]},
           "y": {"dtype": "F16", "shape": [2, 2], "data_offsets": [8, 16]}}
    hj = json.dumps(hdr).encode()
    p = tmp_path / "t.safetensors"
    p.write_bytes(struct.pack("<Q", len(hj)) + hj + bf + f16)
    t = read_safetensors(str(p))
    np.testing.assert_array_equal(t["x"], a)
    np.testing.assert_array_equal(t["y"], a.reshape(2, 2))


def test_streaming_checkpoint_with_config_json(tmp_path):
    """HF MoonshineStreamingForConditionalGeneration layout: config.json carries the dimensions,
    comp.log_k is a scalar, proj_out is a separate (untied) head."""
    import json
    from moonshine_b200.arch import ARCHS
    d = ARCHS["test_streaming"]
    w = synth_weights("test_streaming", 1, "scaled")
    ck = {k: v for k, v in w.items() if k != "streaming.config"}
    ck["model.encoder.embedder.comp.log_k"] = ck["model.encoder.embedder.comp.log_k"].reshape(())
    write_safetensors(str(tmp_path / "model.safetensors"), ck)
    cfg = {"model_type": "moonshine_streaming", "hidden_size": d.dim, "intermediate_size": d.ffn,
           "num_hidden_layers": d.dec_layers, "num_attention_heads": d.heads, "vocab_size": d.vocab,
           "max_position_embeddings": d.max_pos_emb, "tie_word_embeddings": False,
           "rope_parameters": {"rope_type": "default", "rope_theta": 10000.0, "partial_rotary_factor": 0.8},
           "encoder_config": {"hidden_size": d.enc_dim, "intermediate_size": d.enc_ffn,
                              "num_hidden_layers": d.enc_layers, "sliding_windows": [list(x) for x in d.windows]}}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    out = convert(str(tmp_path), str(tmp_path / "model_dir"), "tiny_streaming")
    arch, got = read_msw(os.path.join(out, "model.msw"))
    assert arch == 2
    assert set(got) == set(w)
    for k in w:
        np.testing.assert_array_equal(got[k], w[k])
