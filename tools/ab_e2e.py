#!/usr/bin/env python
"""Development aid: interleaved A/B of the END-TO-END loop of bench.py (pinned host frames in, boxes out, three batches
in flight) under different debug flags.  FLAGS="0,32" ROUNDS=4 STEPS=20 python tools/ab_e2e.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cova_b200 import synth, weights
from cova_b200.elements import BlobPipeline, PinnedBuffer

n_streams, fps = 128, 67
flags = [int(f) for f in os.environ.get("FLAGS", "0,32").split(",")]
rounds, steps = int(os.environ.get("ROUNDS", 4)), int(os.environ.get("STEPS", 20))
frames = synth.tiled_streams(n_streams, fps, 45, 80, 1)
AHEAD = int(os.environ.get("AHEAD", 3))
pins = [PinnedBuffer(frames.shape) for _ in range(AHEAD + 1)]
for i, pb in enumerate(pins):
    pb.array[...] = np.roll(frames, i, axis=0)
p = BlobPipeline(80, 45, weights.to_blob(weights.random_weights(0, head_bias=-1.0)), n_streams, fps, n_chunks=int(os.environ.get("CHUNKS", 1)))
p.process(pins[0].array, raw=True)
for pb in pins:
    p.submit(pb.array)
for _ in pins:
    p.collect(raw=True)
acc = {f: [] for f in flags}
for r in range(rounds):
    for f in flags:
        p.set_debug(f)
        t0 = time.perf_counter()
        for k in range(AHEAD):
            p.submit(pins[k].array)
        for k in range(steps):
            if k + AHEAD < steps:
                p.submit(pins[(k + AHEAD) % len(pins)].array)
            p.collect(raw=True)
        acc[f].append((time.perf_counter() - t0) * 1e3 / steps)
for f in flags:
    v = np.array(acc[f])
    print(f"flags {f:3d}: median {np.median(v):.4f} ms/step  min {v.min():.4f}  max {v.max():.4f} -> {p.n_windows / np.median(v) / 1e3:.3f} M frames/s end to end")
