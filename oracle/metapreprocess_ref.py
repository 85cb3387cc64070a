"""Oracle restatement of the ``metapreprocess`` element (TEST INFRASTRUCTURE).

Follows cova-rs/gst-plugins/src/metapreprocess/imp.rs:
  * caps (``transform_caps`` :247-286): RGBA, width/16, height/16*timestep (integer division)
  * ``set_caps`` :204-236: size_per_buf = out_size / timestep
  * ``transform`` :288-332: the sliding window + gamma sub-sampling

PARITY UNPINNED: the reference has no test or golden output for this element
(SURVEY.md section 8c); this is a line-by-line restatement of the Rust code.
"""
from __future__ import annotations

from collections import deque

import numpy as np

FLOW_OK = 0
FLOW_DROPPED = 1  # gst_base::BASE_TRANSFORM_FLOW_DROPPED


def mb_grid(width_px: int, height_px: int) -> tuple[int, int]:
    """imp.rs:262-268 - integer division, so 1080 rows -> 67 macroblock rows."""
    return width_px // 16, height_px // 16


class MetaPreprocessRef:
    """One element instance == one stream == one sliding window (imp.rs:38-48)."""

    def __init__(self, width_px: int, height_px: int, timestep: int = 1, gamma: int = 1):
        if timestep < 1 or gamma < 1:
            raise ValueError("timestep and gamma are u32 >= 1 (imp.rs:57-80)")
        self.timestep = int(timestep)
        self.gamma = int(gamma)
        self.w_mb, self.h_mb = mb_grid(width_px, height_px)
        # out caps: RGBA w_mb x (h_mb * timestep); VideoInfo::size() of RGBA = w*h*4 (w*4 is 4-aligned)
        self.out_size = self.w_mb * self.h_mb * self.timestep * 4
        self.size_per_buf = self.out_size // self.timestep  # imp.rs:233
        self.gamma_idx = 0
        self.prev = deque()  # push_front / pop_back == appendleft / pop

    def transform(self, inbuf) -> tuple[int, bytes | None]:
        """imp.rs:288-332.  ``inbuf`` is the I420 buffer; only its first size_per_buf bytes are read."""
        buf = bytes(memoryview(inbuf))[: self.size_per_buf] if not isinstance(inbuf, np.ndarray) \
            else inbuf.reshape(-1)[: self.size_per_buf].tobytes()
        if len(buf) < self.size_per_buf:
            raise ValueError("input buffer shorter than size_per_buf")
        S = self.size_per_buf
        if len(self.prev) < self.timestep - 1:
            self.prev.appendleft(buf)
            return FLOW_DROPPED, None
        if self.gamma_idx == 0:
            out = bytearray(self.out_size)
            out[0:S] = buf
            idx = S
            for p in self.prev:  # newest -> oldest
                out[idx: idx + S] = p
                idx += S
            self.prev.appendleft(buf)
            self.prev.pop()
            self.gamma_idx = self.gamma - 1
            return FLOW_OK, bytes(out)
        self.prev.appendleft(buf)
        self.prev.pop()
        self.gamma_idx -= 1
        return FLOW_DROPPED, None


def tensorise_stream(frames: np.ndarray, timestep: int, gamma: int = 1) -> np.ndarray:
    """Run a whole stream through the element.

    frames: u8 [F, H_mb, W_mb, 4] (first S bytes of every decoder buffer).
    returns u8 [N, timestep*H_mb, W_mb, 4] - the stacked RGBA images, row block k = frame t-k.
    """
    F, H, W, C = frames.shape
    assert C == 4
    el = MetaPreprocessRef(W * 16, H * 16, timestep, gamma)
    outs = []
    for f in range(F):
        flow, out = el.transform(frames[f])
        if flow == FLOW_OK:
            outs.append(np.frombuffer(out, dtype=np.uint8).reshape(timestep * H, W, 4))
    if not outs:
        return np.zeros((0, timestep * H, W, 4), dtype=np.uint8)
    return np.stack(outs)


def window_newest_indices(n_frames: int, timestep: int, gamma: int = 1) -> list[int]:
    """Index of the newest frame of every emitted window (closed form of transform())."""
    return [f for f in range(timestep - 1, n_frames) if (f - (timestep - 1)) % gamma == 0]


def stacked_to_nchw(stacked: np.ndarray, timestep: int) -> np.ndarray:
    """nvinfer pre-process + Reshape (config/blobnet/*.txt:7,9; utils/train-blobnet.py:113-116).

    RGBA u8 [N, T*H, W, 4] -> planar RGB float32 [N, 3, T, H, W]; byte 3 is dropped; scale 1.
    """
    N, TH, W, _ = stacked.shape
    H = TH // timestep
    x = stacked[..., :3].astype(np.float32)          # [N, T*H, W, 3]
    x = x.transpose(0, 3, 1, 2)                      # [N, 3, T*H, W]
    return np.ascontiguousarray(x.reshape(N, 3, timestep, H, W))
