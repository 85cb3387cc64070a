"""world_size-2 gloo test (CPU) of the multi-GPU host logic: chain sharding, the optional box gather and
the max-over-ranks timing reduction bench.py uses.  The data path has no collective."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cova_b200 import shard, synth
from oracle import bboxcc_ref


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.streams_of_rank(n_streams, world, rank)
    local = {}
    for s in mine:                                   # CPU stand-in for the per-rank GPU pipeline: oracle boxes
        m = synth.mask_patterns(45, 80, seed=s)["rects_noise"]
        local[s] = [bboxcc_ref.bboxcc_transform_ref(m, 80, 45, 1)]
    merged = shard.gather_blobs(local, world)
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((sorted(merged), float(t.item()), {k: v[0] for k, v in merged.items()}))
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    world, n_streams = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    keys, tmax, blobs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert keys == list(range(n_streams))
    assert tmax == 11.0                               # max over ranks, not rank 0's own time
    for s in range(n_streams):
        m = synth.mask_patterns(45, 80, seed=s)["rects_noise"]
        assert blobs[s] == bboxcc_ref.bboxcc_transform_ref(m, 80, 45, 1)
