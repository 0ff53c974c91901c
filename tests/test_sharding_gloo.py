"""N>1 host logic on CPU: world_size-2 gloo run of the row sharding + whole-job aggregation used by bench.py,
with geometry-only handles standing in for the per-GPU engines (no samples are processed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import G
from gar_b200.shard import job_throughput, partition


def test_partition_covers_all_rows_once():
    for total in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, cnt = partition(total, world, r)
                seen.extend(range(first, first + cnt))
            assert seen == list(range(total))
    with pytest.raises(ValueError):
        partition(10, 2, 2)


def _worker(rank, world, port, total_rows, n_in, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, rows = partition(total_rows, world, rank)
        # each rank owns its rows' carry state; shards never exchange data
        b = G.NewBatch(48000, 16000, G.QualityMedium, max(rows, 1), np.float32, device=-1)
        n1 = sum(b.advance_geometry(n_in, stream=s) for s in range(rows))
        n2 = sum(b.advance_geometry(0, flush=True, stream=s) for s in range(rows))
        dist.barrier()
        # pretend rank r took (r+1) seconds: the job time is the max over ranks
        rate, total, tmax = job_throughput(n1 + n2, float(rank + 1))
        if rank == 0:
            q.put((rate, total, tmax, first, rows))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_sharded_job_matches_single_process():
    total_rows, n_in = 37, 48000
    single = G.NewBatch(48000, 16000, G.QualityMedium, 1, np.float32, device=-1)
    per_row = single.advance_geometry(n_in) + single.advance_geometry(0, flush=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total_rows, n_in, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rate, total, tmax, first, rows = q.get(timeout=10)
    assert total == per_row * total_rows          # SUM over ranks == whole job
    assert tmax == 2.0                            # MAX over ranks
    assert rate == total / 2.0
    assert (first, rows) == (0, 19)
