"""Host-side logic of the multi-GPU frame stream on CPU: world_size-2 gloo processes shard a stream round-robin,
'match' their frames with a stand-in engine, agree on the throughput (max time over ranks) and re-order results."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import stream  # noqa: E402


def test_shard_and_merge():
    ids = list(range(10))
    parts = [stream.shard(ids, r, 4) for r in range(4)]
    assert parts == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    assert all(stream.owner(i, 4) == r for r, p in enumerate(parts) for i in p)
    merged = stream.merge_in_order([[(i, i * i) for i in p] for p in parts])
    assert [m[0] for m in merged] == ids and [m[1] for m in merged] == [i * i for i in ids]
    with pytest.raises(ValueError):
        stream.merge_in_order([[(1, 0)], [(1, 0)]])
    with pytest.raises(ValueError):
        stream.shard(ids, 4, 4)


def test_run_lanes_order_and_waits():
    log = []
    n = stream.run_lanes(3, range(7), lambda ln, fr: log.append(("enq", ln, fr)), lambda ln: log.append(("wait", ln)))
    assert n == 7
    enq = [e for e in log if e[0] == "enq"]
    assert [e[2] for e in enq] == list(range(7)) and [e[1] for e in enq] == [0, 1, 2, 0, 1, 2, 0]
    # a lane is always waited on before it is reused
    busy = set()
    for e in log:
        if e[0] == "enq":
            assert e[1] not in busy
            busy.add(e[1])
        else:
            busy.discard(e[1])
    assert not busy
    assert stream.run_lanes(4, [], lambda *a: None, lambda *a: None) == 0


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = list(range(9))
    mine = stream.shard(frames, rank, world)
    results = []

    def enqueue(lane, fid):   # stand-in engine: a deterministic per-frame checksum
        rng = np.random.default_rng(fid)
        results.append((fid, int(rng.integers(0, 1 << 30))))

    n = stream.run_lanes(2, mine, enqueue, lambda lane: None)
    elapsed = 0.5 + 0.25 * rank          # rank 1 is the slowest: whole-job rate must use ITS time
    fps, total, tmax = stream.reduce_throughput(n, elapsed, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    merged = stream.merge_in_order(gathered)
    if rank == 0:
        q.put((fps, total, tmax, merged))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_stream():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    fps, total, tmax, merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == 9 and abs(tmax - 0.75) < 1e-9 and abs(fps - 9 / 0.75) < 1e-9
    assert [m[0] for m in merged] == list(range(9))
    assert all(m[1] == int(np.random.default_rng(m[0]).integers(0, 1 << 30)) for m in merged)
