#!/usr/bin/env python
"""Stage-by-stage diagnostic on a GPU box: prints parity numbers for every kernel against the oracle
and a first timing table.  Writes gpurun_out/check.json.  (Development aid; the assertions live in tests/.)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cova_b200 import _lib, synth, weights  # noqa: E402
from cova_b200.elements import BboxCc, BlobPipeline, MetaPreprocess  # noqa: E402
from oracle import bboxcc_ref, blobnet_ref, metapreprocess_ref as mpr  # noqa: E402

report = {}


def rel_err(a, ref):
    d = float(np.abs(a.astype(np.float64) - ref.astype(np.float64)).max())
    s = float(np.abs(ref).max()) + 1e-12
    return d, d / s


def check_ccl():
    g = np.load(os.path.join(ROOT, "tests", "golden", "ccl_golden.npz"))
    meta = [m.split(",") for m in g["meta"]]
    bad = 0
    els = {}
    for i, (h, w, name, n) in enumerate(meta):
        h, w, n = int(h), int(w), int(n)
        raw = g[f"raw_{i}"]
        m = raw.reshape(h, w) if raw.size else np.unpackbits(g[f"mask_{i}"])[: h * w].reshape(h, w)
        el = els.setdefault((h, w), BboxCc(w, h, 1))
        n2, l2, s2 = el.labels(m)
        ok = n2 == n and (l2 == g[f"labels_{i}"]).all() and (n == 1 or (s2[1:] == g[f"stats_{i}"]).all())
        for thr in (0, 1, 3, 30):
            el.set_property("cc-threshold", thr)
            ok = ok and el.transform_ip(m) == bboxcc_ref.bboxcc_transform_ref(m, w, h, thr)
        if not ok:
            bad += 1
            print("CCL MISMATCH", h, w, name, n2, n)
    report["ccl_cases"] = len(meta)
    report["ccl_bad"] = bad
    print(f"[ccl] {len(meta)} golden cases, {bad} mismatches")


def check_metapreprocess():
    fr = synth.synth_stream(13, 45, 80, seed=5)
    bad = 0
    for T, gamma in [(1, 1), (4, 1), (4, 2), (3, 3)]:
        el = MetaPreprocess(1280, 720, T, gamma)
        ref = mpr.MetaPreprocessRef(1280, 720, T, gamma)
        for f in range(fr.shape[0]):
            a, b = el.transform(fr[f]), ref.transform(fr[f])
            if a != b:
                bad += 1
    report["metapreprocess_bad"] = bad
    print(f"[metapreprocess element] mismatching buffers: {bad}")


def check_blobnet(h_mb=45, w_mb=80, n_streams=2, fps=6, seed=0):
    w = weights.random_weights(seed)
    blob = weights.to_blob(w)
    frames = synth.synth_streams(n_streams, fps, h_mb, w_mb, config_idx=1)
    # oracle
    stacked = np.concatenate([mpr.tensorise_stream(frames[s], 4, 1) for s in range(n_streams)])
    x = mpr.stacked_to_nchw(stacked, 4)
    logit_ref, inter = blobnet_ref.blobnet_forward(w, x, return_intermediates=True)
    refs = {0: np.clip(x, 0, 6), 1: inter["enc0"], 2: inter["enc1"], 3: inter["enc2"], 4: inter["enc3"][:, :, :1],
            5: np.maximum(inter["dec0"], 0)[:, :, None], 6: np.maximum(inter["dec1"], 0)[:, :, None],
            7: np.maximum(inter["dec2"], 0)[:, :, None]}
    out = {}
    # --- SIMT validation path against the oracle
    ps = BlobPipeline(w_mb, h_mb, blob, n_streams, fps, impl=_lib.IMPL_SIMT, keep_logits=True, keep_stacked=True)
    ps.load_frames(frames)
    ps.run()
    ps.sync()
    st = ps.read_stacked()
    out["stacked_bit_exact"] = bool((st == stacked).all())
    print("[tensorise] stacked RGBA bit-exact:", out["stacked_bit_exact"])
    simt_act = {}
    for layer in range(8):
        a = ps.read_activation(layer)
        simt_act[layer] = a
        d, r = rel_err(a, refs[layer])
        out[f"simt_L{layer}"] = (d, r)
        print(f"[simt vs oracle] layer {layer} shape {a.shape} max|d| {d:.4g} rel {r:.3g}")
    lg = ps.read_logits()
    d, r = rel_err(lg, logit_ref)
    flips = float(((lg > 0) != (logit_ref > 0)).mean())
    out["simt_logits"] = (d, r, flips)
    print(f"[simt vs oracle] logits max|d| {d:.4g} rel {r:.3g} mask flips {flips:.5f} fg {float((logit_ref > 0).mean()):.3f}")
    # --- tcgen05: each layer in isolation on the SIMT inputs, then the whole net
    pt = ps  # same buffers: run layer L with tcgen05 after SIMT populated everything
    for layer in range(8):
        try:
            pt.run_layer(layer, _lib.IMPL_TCGEN05)
            pt.sync()
        except Exception as e:  # noqa: BLE001
            print(f"[tc layer {layer}] FAILED: {e}")
            out[f"tc_L{layer}"] = str(e)
            return out
        if layer < 7:
            a = pt.read_activation(layer + 1)
            d, r = rel_err(a, simt_act[layer + 1])
            d2, r2 = rel_err(a, refs[layer + 1])
            out[f"tc_L{layer}"] = (d, r, d2, r2)
            print(f"[tc layer {layer} isolated] vs simt max|d| {d:.4g} rel {r:.3g} | vs oracle rel {r2:.3g}")
            # restore the validation output so the next layer sees identical inputs
            pt.run_layer(layer, _lib.IMPL_SIMT)
        else:
            lg2 = pt.read_logits()
            d, r = rel_err(lg2, lg)
            out["tc_L7"] = (d, r)
            print(f"[tc layer 7 isolated] logits vs simt max|d| {d:.4g} rel {r:.3g}")
    pf = BlobPipeline(w_mb, h_mb, blob, n_streams, fps, impl=_lib.IMPL_TCGEN05, keep_logits=True)
    boxes = pf.process(frames)
    lg3 = pf.read_logits()
    mask = pf.read_mask()
    d, r = rel_err(lg3, logit_ref)
    flips = float(((lg3 > 0) != (logit_ref > 0)).mean())
    out["tc_logits"] = (d, r, flips)
    print(f"[tc whole net vs oracle] logits max|d| {d:.4g} rel {r:.3g} mask flips {flips:.5f}")
    okb = all(b == bboxcc_ref.bboxcc_transform_ref(mask[i], w_mb, h_mb, 1) for i, b in enumerate(boxes))
    out["boxes_bit_exact_for_mask"] = okb
    print("[pipeline] boxes bit-exact for the device mask:", okb, "windows", len(boxes))
    return out


def timing(h_mb=45, w_mb=80, n_streams=64, fps=35):
    w = weights.random_weights(0)
    blob = weights.to_blob(w)
    frames = synth.tiled_streams(n_streams, fps, h_mb, w_mb, 1)
    out = {}
    for impl, name in ((_lib.IMPL_TCGEN05, "tc"), (_lib.IMPL_SIMT, "simt")):
        p = BlobPipeline(w_mb, h_mb, blob, n_streams, fps, impl=impl)
        p.load_frames(frames)
        p.set_profiling(True)
        for _ in range(3):
            p.run()
            p.sync()
        t = p.last_timings()
        nwin = p.n_windows
        tot = sum(t.values())
        out[name] = {"windows": nwin, "ms": t, "total_ms": tot, "windows_per_s": nwin / tot * 1e3}
        print(f"[timing {name}] windows {nwin} total {tot:.3f} ms -> {nwin / tot * 1e3:.0f} windows/s")
        for k, v in t.items():
            print(f"    {k:16s} {v:8.3f} ms")
        p.close()
    return out


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    t0 = time.time()
    steps = [("ccl", check_ccl), ("metapreprocess", check_metapreprocess), ("blobnet_720p", check_blobnet),
             ("blobnet_small", lambda: check_blobnet(20, 24, 1, 5, seed=1)),
             ("blobnet_1080p", lambda: check_blobnet(68, 120, 1, 5, seed=2)),
             ("timing", timing)]
    for name, fn in steps:
        try:
            report[name] = fn()
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            report[name] = f"EXC {e}"
            if "CUDA" in str(e) or "cuda" in str(e):
                break   # context is gone after a trap
    report["seconds"] = time.time() - t0
    with open(os.path.join(ROOT, "gpurun_out", "check.json"), "w") as f:
        json.dump(report, f, indent=1, default=str)
    print("done in", report["seconds"])
