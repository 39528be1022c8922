"""CPU: the ABI's threading contract ("All API calls are thread-safe ... calculations on a single transcriber are
serialized", core/moonshine-c-api.h:64-67; cf. core/transcriber-concurrency-test.cpp) on the host-side state
machine: many threads create / feed / update / free streams of ONE transcriber and call the one-shot entry point
at the same time (skip_transcription: no device needed)."""
import threading

import numpy as np

from moonshine_b200 import api


def test_streams_and_one_shot_calls_from_many_threads():
    t = api.Transcriber(None, api.ModelArch.TINY, {"skip_transcription": "true", "vad_threshold": "0"})
    errors = []

    def stream_worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for _ in range(6):
                s = t.create_stream()
                s.start()
                fed = 0
                for _ in range(5):
                    n = int(rng.integers(300, 5000))
                    s.add_audio((rng.standard_normal(n) * 0.05).astype(np.float32))
                    fed += n
                    tr = s.update_transcription(api.MOONSHINE_FLAG_FORCE_UPDATE)
                    assert len(tr.lines) <= 1
                    if tr.lines:
                        assert tr.lines[0].audio_data.size == fed // 512 * 512
                s.stop()
                s.update_transcription()
                s.close()
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    def oneshot_worker(seed):
        try:
            rng = np.random.default_rng(100 + seed)
            for _ in range(10):
                n = int(rng.integers(2000, 40000))
                a = (rng.standard_normal(n) * 0.05).astype(np.float32)
                tr = t.transcribe_without_streaming(a)
                assert len(tr.lines) == 1 and tr.lines[0].is_complete
                np.testing.assert_array_equal(tr.lines[0].audio_data, a[: n // 512 * 512])
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=stream_worker, args=(i,)) for i in range(6)]
    # one thread only on the one-shot entry point: its transcript storage is per transcriber and, as in the
    # reference, valid until the NEXT such call -- two threads reading it concurrently would race by contract
    threads += [threading.Thread(target=oneshot_worker, args=(0,))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    t.close()
    assert not errors, errors


def test_worker_pool_runs_every_item_once_and_propagates_exceptions():
    """The persistent pool behind the batch path (segmentation of a batch, staging copies): several callers at once, odd
    item counts, an exception thrown by one item reaches its own caller and nobody else."""
    from moonshine_b200 import api
    lib = api.load_library()
    for callers, n, rounds in [(1, 1, 3), (1, 37, 20), (4, 257, 10), (8, 16, 50), (3, 1000, 5)]:
        assert lib.moonshine_b200_debug_pool_selftest(callers, n, rounds) == 0, (callers, n, rounds)
