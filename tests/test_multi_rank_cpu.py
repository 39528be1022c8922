"""world_size-2 gloo test of the N>1 host logic: sharding, the one weight
broadcast, and result collection (no GPU involved)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from moonshine_b200.dist import broadcast_bytes, gather_token_lists, shard_range
from moonshine_b200.weights import pack_msw, read_msw, synth_weights


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 32, 2048, 2049):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                assert 0 <= a <= b <= n
                seen += list(range(a, b))
            assert seen == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = pack_msw("test", synth_weights("test", 5, "scaled")) if rank == 0 else b""
    got = broadcast_bytes(blob, 0)
    arch, w = read_msw(got)
    a, b = shard_range(9, rank, world)
    toks = [[1, 100 + i, 2] for i in range(a, b)]
    allt = gather_token_lists(toks, 0)
    if rank == 0:
        q.put((len(got), arch, float(w["model.decoder.norm.weight"].sum()), allt))
    else:
        q.put((len(got), arch, float(w["model.decoder.norm.weight"].sum()), None))
    dist.destroy_process_group()


def test_weight_broadcast_and_gather_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = synth_weights("test", 5, "scaled")
    want = float(ref["model.decoder.norm.weight"].sum())
    assert all(r[1] == 100 and abs(r[2] - want) < 1e-6 for r in res)
    assert res[0][0] == res[1][0] > 1000
    gathered = [r[3] for r in res if r[3] is not None][0]
    flat = [t for part in gathered for t in part]
    assert flat == [[1, 100 + i, 2] for i in range(9)]
