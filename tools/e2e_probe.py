#!/usr/bin/env python
"""Development aid: where does the end-to-end time go (submit / collect wall times)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cova_b200 import synth, weights
from cova_b200.elements import BlobPipeline, PinnedBuffer

n_streams, fps = 128, 67
chunks = int(os.environ.get("CHUNKS", 1))
frames = synth.tiled_streams(n_streams, fps, 45, 80, 1)
pins = [PinnedBuffer(frames.shape), PinnedBuffer(frames.shape)]
for pb in pins:
    pb.array[...] = frames
p = BlobPipeline(80, 45, weights.to_blob(weights.random_weights(0, head_bias=-1.0)), n_streams, fps, n_chunks=chunks)
for _ in range(2):
    t0 = time.perf_counter(); p.process(pins[0].array, raw=True); print("process", (time.perf_counter() - t0) * 1e3, "ms")
ts = []
t_all = time.perf_counter()
t0 = time.perf_counter(); p.submit(pins[0].array); ts.append(("submit0", time.perf_counter() - t0))
N = 8
for k in range(N):
    if k + 1 < N:
        t0 = time.perf_counter(); p.submit(pins[(k + 1) & 1].array); ts.append((f"submit{k+1}", time.perf_counter() - t0))
    t0 = time.perf_counter(); p.collect(raw=True); ts.append((f"collect{k}", time.perf_counter() - t0))
tot = time.perf_counter() - t_all
for n, t in ts:
    print(f"{n:10s} {t*1e3:8.3f} ms")
print("per step", tot / N * 1e3, "ms ->", p.n_windows * N / tot, "frames/s; blob bytes", p.last_blob_len)
