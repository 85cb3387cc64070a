#!/usr/bin/env python
"""Development aid: interleaved A/B timing of whole steps under different debug flags (clock drift under the power cap
makes back-to-back columns of layer_timing.py incomparable).  FLAGS="0,32,64" ROUNDS=8 python tools/ab_step.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cova_b200 import synth, weights
from cova_b200.elements import BlobPipeline

n_streams, fps = int(os.environ.get("STREAMS", 128)), 67
h, w = int(os.environ.get("H", 45)), int(os.environ.get("W", 80))
flags = [int(f) for f in os.environ.get("FLAGS", "0,32,64,96").split(",")]
rounds = int(os.environ.get("ROUNDS", 8))
p = BlobPipeline(w, h, weights.to_blob(weights.random_weights(0, head_bias=-1.0)), n_streams, fps, n_chunks=int(os.environ.get("CHUNKS", 1)))
p.load_frames(synth.tiled_streams(n_streams, fps, h, w, 1))
for _ in range(5):
    p.run()
p.sync()
acc = {f: [] for f in flags}
for r in range(rounds):
    for f in flags:
        p.set_debug(f)
        p.run(); p.sync()
        t0 = time.perf_counter()
        for _ in range(5):
            p.run()
        p.sync()
        acc[f].append((time.perf_counter() - t0) * 200.0)
for f in flags:
    v = np.array(acc[f])
    print(f"flags {f:3d}: median {np.median(v):.4f} ms  min {v.min():.4f}  max {v.max():.4f}  -> {p.n_windows / np.median(v) / 1e3:.3f} M windows/s")
