"""GPU: one transcriber handle driving several device contexts (the additive `devices` option).  A batch
is cut into contiguous shards, one per context, each decoded by its own host thread against a device-to-device
replica of the weights; the merged transcript must equal the single-context transcript line for line (text,
word timings).  On a one-GPU box the two contexts share the device ("0,0"), which still exercises the replica
constructor, the sharding and the merge; with two or more GPUs the same test runs across real peers."""
import numpy as np
import pytest
import torch

from moonshine_b200 import api
from moonshine_b200.weights import synth_audio
from tests.util import memory_files

pytestmark = pytest.mark.gpu


def clips(n, seed):
    rng = np.random.default_rng(seed)
    return [synth_audio(seed + i, int(16000 * rng.uniform(0.8, 3.0))) for i in range(n)]


def flat(tr):
    return [(l.text, round(l.start_time, 4), round(l.duration, 4),
             [(w.word, round(w.start, 4), round(w.end, 4), round(w.confidence, 5)) for w in (l.words or [])]) for l in tr.lines]


@pytest.mark.parametrize("arch,model_arch", [("test", api.ModelArch.TEST), ("test_streaming", api.ModelArch.TEST_STREAMING)])
@pytest.mark.parametrize("n_clips", [1, 5, 8])
def test_sharded_batch_equals_the_single_context_batch(arch, model_arch, n_clips):
    n_gpu = torch.cuda.device_count()
    devices = "0,0,0" if n_gpu < 2 else ",".join(str(i) for i in range(min(n_gpu, 4)))
    opts = {"vad_threshold": "0", "word_timestamps": "true"}
    audios = clips(n_clips, 100 + n_clips)
    with api.Transcriber(model_arch=model_arch, options=opts, memory_files=memory_files(arch, 0)) as one:
        want = [flat(t) for t in one.transcribe_batch_without_streaming(audios)]
    with api.Transcriber(model_arch=model_arch, options=dict(opts, devices=devices), memory_files=memory_files(arch, 0)) as many:
        got = [flat(t) for t in many.transcribe_batch_without_streaming(audios)]
        again = [flat(t) for t in many.transcribe_batch_without_streaming(audios[::-1])][::-1]
    assert got == want
    assert again == want
    assert any(line[0] for t in want for line in t)


def test_a_bad_device_list_is_an_error():
    with pytest.raises(Exception):
        api.Transcriber(model_arch=api.ModelArch.TEST, options={"devices": "0,99"}, memory_files=memory_files("test", 0))
    with pytest.raises(Exception):
        api.Transcriber(model_arch=api.ModelArch.TEST, options={"devices": ","}, memory_files=memory_files("test", 0))
