#!/usr/bin/env python
"""Benchmark of the hot path: batched Moonshine transcription (conv frontend ->
encoder -> greedy decoder) on B200, against the CPU oracle port.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--model tiny|base] [--batch B]

One "step" = one pass of the hot path over one batch of B synthetic 10 s @
16 kHz utterances (BASELINE.json configs[1]: moonshine-tiny, batch 32).
`value` = whole-job utterances/s with the PCM already resident in HBM
(moonshine_b200_transcribe_device); `e2e` = the same through the
reference-facing C-ABI call with HOST buffers
(moonshine_transcribe_batch_without_streaming: segmentation, H2D, encoder,
decode, D2H, detokenisation).  N > 1: one process per GPU (torchrun), weights
NCCL-broadcast once at init, utterances sharded, no per-step collective.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from moonshine_b200.arch import ARCHS, frontend_lengths  # noqa: E402
from moonshine_b200.weights import pack_msw, synth_audio, synth_tokenizer_bin, synth_weights  # noqa: E402

N_SAMPLES = 160000  # 10 s @ 16 kHz


def decoder_step_bytes(d, B, T, t, e_w=4, e_cross=2, e_self=4):
    """Algorithmic HBM bytes of ONE decoder-step launch at decode position t
    (BASELINE.md section 3 formula with this build's element sizes: fp32
    weights, fp16 cross K/V, fp32 self K/V)."""
    D, I, L, V = d.dim, d.ffn, d.dec_layers, d.vocab
    p_step = L * (3 * D * D + D * D + D * D + D * D + 2 * I * D + 2 * I + I * D + D + 3 * D) + V * D + D
    cross = 2 * L * T * D
    self_r = 2 * L * t * D
    self_w = 2 * L * D
    return e_w * p_step + B * (e_cross * cross + e_self * (self_r + self_w) + 4)


def clocks_sampler(stop_evt, out):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    def reader():
        for line in p.stdout:
            out.append(line.strip())
    th = threading.Thread(target=reader, daemon=True)
    th.start()
    stop_evt.wait()
    p.terminate()
    th.join(timeout=2)


def summarise_clocks(lines, dev_index):
    sm, mx, reasons = [], [], set()
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 9 or not f[0].isdigit() or int(f[0]) != dev_index:
            continue
        try:
            sm.append(float(f[1])); mx.append(float(f[2]))
        except ValueError:
            continue
        for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    # under load = samples above idle clocks
    hi = [x for x in sm if x > 0.5 * max(sm)]
    return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
            "samples": len(sm)}


CPU_THREADS = int(os.environ.get("MOONSHINE_ORACLE_THREADS", "0")) or min(os.cpu_count() or 1, 16)


def oracle_baseline(model, weights, audios, budget_s=20.0, max_utts=8):
    """Times the numpy oracle.  BLAS threads are capped (default 16): the
    per-token decoder matmuls are tiny and 100+ spinning BLAS threads make the
    port slower, which would flatter the GPU."""
    from oracle.moonshine_oracle import Dims, Oracle
    from oracle.moonshine_streaming_oracle import SDims, StreamingOracle
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=CPU_THREADS)
    except Exception:  # pragma: no cover
        import contextlib
        ctx = contextlib.nullcontext()
    with ctx:
        streaming = ARCHS[model].streaming
        o = (StreamingOracle(SDims.from_product(ARCHS[model]), weights) if streaming
             else Oracle(Dims.from_product(ARCHS[model]), weights))
        t0 = time.perf_counter()
        n = 0
        toks = []
        for a in audios[:max_utts]:
            tk = (o.transcribe_segment(a, keep_logits=False) if streaming else o.greedy(a, keep_logits=False))[0]
            toks.append(tk)
            n += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return n / dt, n, dt, toks


def run_reference(args):
    """--impl reference: the reference's CPU path for this workload.  The
    reference's own ORT graphs/weights are not available,
    so this is the oracle port of the float model, on all host cores (numpy
    BLAS threads), same seeded weights / inputs / metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = args.model
    weights = synth_weights(model, 0, "hf")
    per_step = 2  # bounded sample: utterances per step
    audios = [synth_audio(i, N_SAMPLES) for i in range(per_step)]
    for _ in range(args.warmup):
        oracle_baseline(model, weights, audios[:1], max_utts=1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_baseline(model, weights, audios, budget_s=1e9, max_utts=per_step)
    dt = time.perf_counter() - t0
    ups = args.steps * per_step / dt
    cores = CPU_THREADS
    line = {
        "impl": "reference", "metric": "utterances_per_sec", "value": ups, "unit": "utt/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "rtf": (dt / (args.steps * per_step)) / 10.0,
        "config": {"workload": f"moonshine-{model} greedy transcription of synthetic 10s@16kHz utterances "
                               f"(configs[1] inputs), CPU oracle port, {per_step} utterances per step",
                   "batch": per_step, "weights": "seeded synthetic (HF init), fp32"},
        "cpu_baseline": {"value": ups, "unit": "utt/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps}x{per_step} utterances of the workload, numpy oracle, {cores} BLAS threads (host has {os.cpu_count()} cores)"},
        "e2e": {"value": ups, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--model", default="tiny")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from moonshine_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model, B = args.model, args.batch
    d = ARCHS[model]
    arch_enum = {"tiny": api.ModelArch.TINY, "base": api.ModelArch.BASE,
                 "tiny_streaming": api.ModelArch.TINY_STREAMING, "base_streaming": api.ModelArch.BASE_STREAMING}[model]

    # ---- weights: rank 0 builds the container, ONE NCCL broadcast replicates it ----
    from moonshine_b200.dist import broadcast_bytes
    weights = synth_weights(model, 0, "hf") if rank == 0 else None
    blob = pack_msw(model, weights) if rank == 0 else b""
    msw = broadcast_bytes(blob, 0, device=f"cuda:{local}")
    del blob
    tr = api.Transcriber(model_arch=arch_enum, options={"vad_threshold": "0", "device": str(local), "return_audio_data": "false"},
                         memory_files={"model.msw": msw, "tokenizer.bin": synth_tokenizer_bin(d.vocab)})
    tr.set_timing(True)

    # ---- inputs: utterances rank*B .. rank*B+B-1 ----
    audios = [synth_audio(rank * B + i, N_SAMPLES) for i in range(B)]
    host = np.stack(audios)
    dev = torch.from_numpy(host).cuda()
    lengths = [N_SAMPLES] * B
    stream = torch.cuda.ExternalStream(tr.cuda_stream_ptr(), device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(max(args.warmup, 1)):
        toks = tr.transcribe_device(dev.data_ptr(), N_SAMPLES, lengths)
    tm = tr.last_timings()
    launches_per_step = int(tm["kernel_launches"])
    n_tokens = [len(t) - 1 for t in toks]

    # ---- timed: device-resident inputs, CUDA events on the library's stream ----
    clock_lines, stop_evt = [], threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(stop_evt, clock_lines), daemon=True)
    th.start()
    time.sleep(0.3)
    sync_all()
    wall0 = time.perf_counter()
    ev_ms, dec_ms, enc_ms, fe_ms, xkv_ms = [], [], [], [], []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        tr.transcribe_device(dev.data_ptr(), N_SAMPLES, lengths)
        e1.record(stream)
        e1.synchronize()
        ev_ms.append(e0.elapsed_time(e1))
        tm = tr.last_timings()
        dec_ms.append(tm["decode_ms"]); enc_ms.append(tm["encoder_ms"])
        fe_ms.append(tm["frontend_ms"]); xkv_ms.append(tm["cross_kv_ms"])
    sync_all()
    wall_dev = time.perf_counter() - wall0
    dev_total_ms = float(sum(ev_ms))

    # ---- timed: end to end through the reference-facing ABI with host buffers ----
    for _ in range(2):
        tr.transcribe_batch_without_streaming(audios)
    sync_all()
    e2e_total = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        # the call is synchronous (it returns the transcripts), so wall clock around it is end to end;
        # the L2 flush stays outside
        w0 = time.perf_counter()
        res = tr.transcribe_batch_without_streaming(audios)
        e2e_total += time.perf_counter() - w0
    sync_all()
    stop_evt.set()
    th.join(timeout=3)

    # ---- max over ranks ----
    t = torch.tensor([dev_total_ms, e2e_total, wall_dev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_total_ms, e2e_total, wall_dev = [float(x) for x in t.tolist()]

    if rank == 0:
        K = args.steps
        utts = world * B * K
        value = utts / (dev_total_ms / 1000.0)
        e2e_value = utts / e2e_total
        ms_per_step = dev_total_ms / K
        # roofline of the dominant kernel (decoder step): algorithmic bytes / launch / measured time
        _, _, T = frontend_lengths(N_SAMPLES)
        if d.streaming:
            T = N_SAMPLES // 1280 * 4   # encoder features of the whole 1280-sample chunks (20 ms each)
        steps_run = int(tm["decode_steps"])
        bytes_per_launch = float(np.mean([decoder_step_bytes(d, B, T, s) for s in range(steps_run)]))
        launch_ms = float(np.mean(dec_ms)) / max(steps_run, 1)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
        line = {
            "metric": "utterances_per_sec", "value": value, "unit": "utt/s", "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "rtf": (dev_total_ms / 1000.0) / (utts * 10.0),
            "config": {
                "workload": f"moonshine-{model}, batch={B} synthetic 10s@16kHz utterances per GPU, "
                            + ("streaming frontend + sliding-window encoder + adapter" if d.streaming else "conv frontend + encoder")
                            + f" + greedy decoder ({steps_run} decode steps; seeded random weights never emit EOS)",
                "batch_per_gpu": B, "global_batch": B * world, "audio_seconds_per_utt": 10.0,
                "weights": "seeded synthetic (HF init std 0.02), fp32 storage; cross K/V cache fp16",
                "parallelism": f"dp{world} (utterance shards, one NCCL weight broadcast at init)",
                "timing": "CUDA events on the library stream per step; 256 MB L2 flush between steps (outside the events)",
            },
            "stage_ms": {"frontend": float(np.mean(fe_ms)), "encoder": float(np.mean(enc_ms)),
                         "cross_kv": float(np.mean(xkv_ms)), "decode": float(np.mean(dec_ms)),
                         "decode_launch_us": 1000.0 * launch_ms},
            "roofline": {"kernel": "decoder_step2_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed
                         # ncu --set full capture of this exact workload (profiles/r1c_decoder_step2_ncu_full_tiny_b32.txt)
                         "traffic": 194202624.0 if (model == "tiny" and B == 32) else None,
                         "bytes_per_launch": bytes_per_launch,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            "e2e": {"value": e2e_value, "unit": "utt/s", "h2d_bytes_per_step": B * N_SAMPLES * 4,
                    "d2h_bytes_per_step": int(sum(len(x) for x in toks) * 4 + 4 * B),
                    "ms_per_step": 1000.0 * e2e_total / K,
                    "api": "moonshine_transcribe_batch_without_streaming (host PCM -> transcript_t; options vad_threshold=0, return_audio_data=false)"},
            "gpu_launches": launches_per_step * K,
            "tokens_per_utt": float(np.mean(n_tokens)),
            "clocks": summarise_clocks(clock_lines, local),
            "wall_s_device_loop": wall_dev,
        }
        if not args.no_cpu_baseline:
            ups, n, dt, ref_toks = oracle_baseline(model, weights, audios)
            line["cpu_baseline"] = {"value": ups, "unit": "utt/s", "cores": CPU_THREADS, "kind": "port",
                                    "sample": f"first {n} utterances of the batch, numpy oracle (fp32, {CPU_THREADS} BLAS threads of {os.cpu_count()} host cores), {dt:.1f}s"}
            line["cpu_baseline"]["tokens_match_gpu"] = bool(all(ref_toks[i] == toks[i] for i in range(n)))
        print(json.dumps(line))
    tr.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
