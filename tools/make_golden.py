#!/usr/bin/env python
"""Generate tests/golden/*.npz - run in the BUILD container only (needs cv2; reads nothing from
/root/reference).  Committed so that the GPU box and CI never need to regenerate them.

ccl_golden.npz  outputs of cv2.connectedComponentsWithStats(mask, 8, CV_32S) - the same third-party
                entry point the reference's bboxcc calls (cova-rs/gst-plugins/src/bboxcc/process.rs:23-30)
                - on the mask families of SURVEY.md section 8d, with OpenCV run single- and
                multi-threaded (both its sequential and its striped-parallel labeller).
"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cova_b200.synth import mask_patterns  # noqa: E402


def main():
    out = {}
    meta = []
    idx = 0
    sizes = [(45, 80), (67, 120), (68, 120), (135, 240), (7, 9), (1, 17), (17, 1), (2, 2), (1, 1), (3, 64)]
    for (h, w) in sizes:
        pats = mask_patterns(h, w, seed=h * 1000 + w) if min(h, w) >= 2 else {
            "bern": (np.random.default_rng(h + w).random((h, w)) < 0.5).astype(np.uint8),
            "ones": np.ones((h, w), np.uint8), "zeros": np.zeros((h, w), np.uint8)}
        for name, m in pats.items():
            res = []
            for threads in (1, 8):
                cv2.setNumThreads(threads)
                n, labels, stats, _ = cv2.connectedComponentsWithStats(m, connectivity=8, ltype=cv2.CV_32S)
                res.append((n, labels.copy(), stats.copy()))
            assert res[0][0] == res[1][0] and (res[0][1] == res[1][1]).all() and (res[0][2][1:] == res[1][2][1:]).all()
            n, labels, stats = res[0]
            out[f"mask_{idx}"] = np.packbits(m != 0)
            out[f"raw_{idx}"] = m if name == "nonbinary" else np.zeros(0, np.uint8)
            out[f"labels_{idx}"] = labels.astype(np.uint16)
            out[f"stats_{idx}"] = stats[1:].astype(np.int32)
            meta.append(f"{h},{w},{name},{n}")
            idx += 1
    out["meta"] = np.array(meta)
    out["opencv_version"] = np.array(cv2.__version__)
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ccl_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, idx, "cases", os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
