#!/usr/bin/env python
"""How much host->device bandwidth does the PLATFORM give N GPUs at once?  One rank per GPU (torchrun), every rank copies a
123 MB pinned buffer (one 720p benchmark batch) to its GPU in a loop, all ranks at the same time; prints per-rank and
aggregate GB/s.  The end-to-end curve of bench.py at N > 1 is bounded by this number x (1 frame / 14.4 KB).

    for n in 1 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port 29517 tools/pcie_concurrent.py; done
"""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 123494400
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(rank + 1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h_out = torch.empty(n // 3, dtype=torch.uint8).pin_memory()
d_out = torch.empty(n // 3, dtype=torch.uint8, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for name, with_d2h in (("h2d", False), ("h2d_with_d2h", True)):
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 30
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s_in):
            d.copy_(h, non_blocking=True)
        if with_d2h:
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([n / dt / 1e9], device="cuda")
    lo, hi, tot = t.clone(), t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    res[name] = {"aggregate_GBps": round(float(tot), 1), "per_rank_min_GBps": round(float(lo), 1), "per_rank_max_GBps": round(float(hi), 1)}
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_copy": n, "host_cpus": os.cpu_count(), **res,
                      "frames_per_s_cap_720p": round(res["h2d_with_d2h"]["aggregate_GBps"] * 1e9 / 14400)}))
if world > 1:
    dist.destroy_process_group()
