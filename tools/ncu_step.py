#!/usr/bin/env python
"""Development aid: N whole-path steps on device-resident frames, for `ncu` captures (profiles/)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cova_b200 import synth, weights
from cova_b200.elements import BlobPipeline

CFG = {"c2": (45, 80, 128, -1.0), "c3": (68, 120, 64, -1.0), "c4": (135, 240, 16, 0.0)}   # bench.py CONFIGS
if os.environ.get("CONFIG"):
    h, w, n_streams, hb = CFG[os.environ["CONFIG"]]
    os.environ.setdefault("HEAD_BIAS", str(hb))
    fps = 67
else:
    n_streams, fps = int(os.environ.get("STREAMS", 128)), 67
    h, w = int(os.environ.get("H", 45)), int(os.environ.get("W", 80))
steps = int(os.environ.get("STEPS", 4))
p = BlobPipeline(w, h, weights.to_blob(weights.random_weights(0, head_bias=float(os.environ.get("HEAD_BIAS", -1.0)))), n_streams, fps, n_chunks=int(os.environ.get("CHUNKS", 1)))
p.load_frames(synth.tiled_streams(n_streams, fps, h, w, 1))
if os.environ.get("DBG"):
    p.set_debug(int(os.environ["DBG"]))       # cova_pipeline_set_debug flags (experiments)
for _ in range(steps):
    p.run()
    p.sync()
print("windows", p.n_windows, "launches", p.launch_count())
