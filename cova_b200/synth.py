"""Synthetic inputs for the blob-detection path (no datasets are available offline).

Streams imitate what the patched entropy decoder writes (reference
third_parties/FFmpeg/libavcodec/h264_mb.c:822-855): one 4-byte quad per macroblock -
byte0 mb_weight in {0..6}, byte1 |mv_x|, byte2 |mv_y| (quarter-pel, clipped to u8), byte3 never
written (stale garbage that must be carried bit-exactly by tensorise and ignored by BlobNet).
Value distribution follows the dump of demo/1m.mp4 described in SURVEY.md section 8c/8d:
~88 % skip MBs (weight 1), ~98 % zero MVs, an all-intra frame every 250 frames.
"""
from __future__ import annotations

import numpy as np

GOP = 250  # cova/imp.rs:258 assumes 250-frame GoPs; demo/1m.mp4 has one IDR per ~250 frames


def stream_seed(config_idx: int, stream_id: int) -> int:
    return 1000 * config_idx + stream_id


def synth_stream(n_frames: int, h_mb: int, w_mb: int, seed: int = 0, start_frame: int = 0) -> np.ndarray:
    """u8 [n_frames, h_mb, w_mb, 4]: moving rectangles + salt noise over a skip-MB background."""
    rng = np.random.default_rng(seed)
    fr = np.zeros((n_frames, h_mb, w_mb, 4), dtype=np.uint8)
    fr[..., 0] = 1
    area_scale = (h_mb * w_mb) / (45 * 80)
    k = int(rng.integers(2, 9) * max(1.0, area_scale))
    for _ in range(k):
        rh = int(rng.integers(2, max(3, h_mb // 6 + 1)))
        rw = int(rng.integers(2, max(3, w_mb // 6 + 1)))
        y = float(rng.uniform(0, h_mb - rh))
        x = float(rng.uniform(0, w_mb - rw))
        vy, vx = rng.uniform(-1, 1, 2)
        for f in range(n_frames):
            yi, xi = int(round(y)), int(round(x))
            ys, xs = slice(max(0, yi), min(h_mb, yi + rh)), slice(max(0, xi), min(w_mb, xi + rw))
            shp = fr[f, ys, xs, 0].shape
            fr[f, ys, xs, 0] = rng.integers(2, 6, shp, dtype=np.uint8)
            fr[f, ys, xs, 1] = np.minimum(255, rng.geometric(0.15, shp)).astype(np.uint8)
            fr[f, ys, xs, 2] = np.minimum(255, rng.geometric(0.25, shp)).astype(np.uint8)
            y += vy
            x += vx
            if y < 0 or y > h_mb - rh:
                vy = -vy
                y = min(max(y, 0.0), float(h_mb - rh))
            if x < 0 or x > w_mb - rw:
                vx = -vx
                x = min(max(x, 0.0), float(w_mb - rw))
    salt = rng.random((n_frames, h_mb, w_mb)) < 0.01
    n_salt = int(salt.sum())
    fr[..., 0][salt] = rng.choice(np.array([2, 3, 4, 6], dtype=np.uint8), n_salt)
    fr[..., 1][salt] = rng.integers(0, 12, n_salt, dtype=np.uint8)
    for f in range(n_frames):
        if (start_frame + f) % GOP == 0:           # I-frame: all intra, no MVs
            fr[f, ..., 0] = 6
            fr[f, ..., 1:3] = 0
    fr[..., 3] = rng.integers(0, 5, (n_frames, h_mb, w_mb), dtype=np.uint8)   # stale byte
    return fr


def synth_streams(n_streams: int, n_frames: int, h_mb: int, w_mb: int, config_idx: int = 0,
                  first_stream: int = 0) -> np.ndarray:
    """u8 [n_streams, n_frames, h_mb, w_mb, 4]."""
    return np.stack([synth_stream(n_frames, h_mb, w_mb, stream_seed(config_idx, first_stream + s))
                     for s in range(n_streams)])


def tiled_streams(n_streams: int, n_frames: int, h_mb: int, w_mb: int, config_idx: int = 0,
                  n_unique: int = 8) -> np.ndarray:
    """Large bench inputs: ``n_unique`` generated streams, repeated with a per-copy roll along x so the
    copies are not byte-identical.  Generation cost stays bounded; content statistics are unchanged."""
    base = synth_streams(min(n_unique, n_streams), n_frames, h_mb, w_mb, config_idx)
    out = np.empty((n_streams,) + base.shape[1:], dtype=np.uint8)
    for s in range(n_streams):
        out[s] = np.roll(base[s % base.shape[0]], shift=(s // base.shape[0]) * 3, axis=2)
    return out


# ----------------------------------------------------------------------------- masks for CCL
def mask_patterns(h: int, w: int, seed: int = 0) -> dict[str, np.ndarray]:
    """The CCL-only mask families of SURVEY.md section 8d (i)-(vi)."""
    rng = np.random.default_rng(seed)
    out = {}
    m = np.zeros((h, w), np.uint8)
    for _ in range(int(rng.integers(2, 9))):
        rh, rw = int(rng.integers(2, max(3, h // 6 + 1))), int(rng.integers(2, max(3, w // 6 + 1)))
        y, x = int(rng.integers(0, h - rh + 1)), int(rng.integers(0, w - rw + 1))
        m[y:y + rh, x:x + rw] = 1
    m[rng.random((h, w)) < 0.01] = 1
    out["rects_noise"] = m
    for p in (0.05, 0.2, 0.4, 0.5, 0.6, 0.8):
        out[f"bernoulli_{p}"] = (rng.random((h, w)) < p).astype(np.uint8)
    cb = np.zeros((h, w), np.uint8)
    cb[::2, ::2] = 1
    out["checker_sparse"] = cb                       # ceil(h/2)*ceil(w/2) single-pixel components
    cb2 = np.zeros((h, w), np.uint8)
    cb2[(np.add.outer(np.arange(h), np.arange(w)) % 2) == 0] = 1
    out["checker_diag"] = cb2                        # one component through diagonals
    out["serpentine"] = serpentine(h, w)
    out["spiral"] = spiral(h, w)
    out["zeros"] = np.zeros((h, w), np.uint8)
    out["ones"] = np.ones((h, w), np.uint8)
    vc = np.zeros((h, w), np.uint8); vc[:, ::2] = 1; vc[0, :] = 1
    out["comb_down"] = vc
    vc2 = np.zeros((h, w), np.uint8); vc2[:, ::2] = 1; vc2[-1, :] = 1
    out["comb_up"] = vc2
    hc = np.zeros((h, w), np.uint8); hc[::2, :] = 1; hc[:, 0] = 1
    out["comb_right"] = hc
    nz = (rng.random((h, w)) < 0.3).astype(np.uint8) * rng.integers(1, 256, (h, w)).astype(np.uint8)
    out["nonbinary"] = nz                            # any non-zero byte is foreground
    return out


def serpentine(h: int, w: int) -> np.ndarray:
    """1-px path snaking through the grid: rows 0,2,4.. full, joined alternately right/left."""
    m = np.zeros((h, w), np.uint8)
    m[::2, :] = 1
    for i, y in enumerate(range(1, h, 2)):
        if y + 1 < h:
            m[y, w - 1 if i % 2 == 0 else 0] = 1
    return m


def spiral(h: int, w: int) -> np.ndarray:
    """1-px spiral with 1-px gaps (a wall-following walker): the longest label chain a grid allows."""
    m = np.zeros((h, w), np.uint8)
    dirs = ((0, 1), (1, 0), (0, -1), (-1, 0))
    y = x = d = turns = 0
    m[0, 0] = 1
    while turns < 2:
        dy, dx = dirs[d]
        ny, nx, ay, ax = y + dy, x + dx, y + 2 * dy, x + 2 * dx
        ok = 0 <= ny < h and 0 <= nx < w and m[ny, nx] == 0
        if ok and 0 <= ay < h and 0 <= ax < w and m[ay, ax]:
            ok = False
        if ok:
            y, x = ny, nx
            m[y, x] = 1
            turns = 0
        else:
            d = (d + 1) % 4
            turns += 1
    return m
