#!/usr/bin/env python
"""Development aid: CCL kernel time on the benchmark's own masks and on the synthetic mask patterns."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cova_b200 import synth, weights
from cova_b200.elements import BlobPipeline

def time_ccl(p, masks, reps=5):
    p.load_masks(masks)
    p.set_profiling(True)
    ts = []
    for _ in range(reps):
        p.ccl(); p.sync()
        ts.append(p.last_timings().get("ccl_bbox", float("nan")))
    p.set_profiling(False)
    boxes = p.fetch_boxes_raw()[2]
    return float(np.median(ts[1:])), float(((boxes.astype(np.int64) - 8) // 24).mean())

CASES = ((45, 80, 128, -1.0), (68, 120, 64, -1.0), (135, 240, 16, -1.0), (135, 240, 16, 0.0))
if os.environ.get("CASES"):             # e.g. CASES=3 -> only the dense 4K case
    CASES = tuple(CASES[int(i)] for i in os.environ["CASES"].split(","))
for (h, w, n_streams, hb) in CASES:
    fps = 67
    p = BlobPipeline(w, h, weights.to_blob(weights.random_weights(0, head_bias=hb)), n_streams, fps, n_chunks=1)
    n = p.windows_for(n_streams, fps)
    p.load_frames(synth.tiled_streams(n_streams, fps, h, w, 1)); p.run(); p.sync()
    m = p.read_mask()
    ms, nb = time_ccl(p, m)
    if os.environ.get("ONLY_NET"):
        print(f"{h}x{w} blobnet masks hb={hb} n={n} {ms*1e3:8.1f} us  {n/ms/1e3:8.2f} M masks/s  boxes/frame {nb:.1f}  fg {m.mean():.3f}")
        continue
    print(f"{h}x{w} blobnet masks hb={hb} n={n} {ms*1e3:8.1f} us  {n/ms/1e3:8.2f} M masks/s  boxes/frame {nb:.1f}  fg {m.mean():.3f}")
    for name, pat in synth.mask_patterns(h, w, seed=1).items():
        mm = np.broadcast_to(pat, (n,) + pat.shape).copy()
        ms, nb = time_ccl(p, mm)
        print(f"{h}x{w} {name:20s} n={n} {ms*1e3:8.1f} us  {n/ms/1e3:8.2f} M masks/s  boxes/frame {nb:.1f}  fg {pat.mean():.3f}")
