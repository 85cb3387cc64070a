"""CPU tests: the oracle against the committed golden vectors and against itself (Python vs C)."""
import struct

import numpy as np
import pytest

from cova_b200 import synth
from oracle import bboxcc_ref, c_oracle, metapreprocess_ref as mpr

G = np.load(__file__.rsplit("/", 1)[0] + "/golden/ccl_golden.npz")
META = [m.split(",") for m in G["meta"]]


def golden_case(i):
    h, w, name, n = META[i]
    h, w, n = int(h), int(w), int(n)
    raw = G[f"raw_{i}"]
    m = raw.reshape(h, w) if raw.size else np.unpackbits(G[f"mask_{i}"])[: h * w].reshape(h, w)
    return h, w, name, n, m, G[f"labels_{i}"].astype(np.int32), G[f"stats_{i}"]


@pytest.mark.parametrize("i", range(len(META)))
def test_ccl_matches_opencv_golden(i):
    h, w, name, n, m, labels, stats = golden_case(i)
    for impl in (bboxcc_ref.ccl_fast, c_oracle.ccl) + ((bboxcc_ref.ccl_ref,) if h * w <= 3600 else ()):
        n2, l2, s2 = impl(m)
        assert n2 == n, (name, impl.__name__)
        assert (l2 == labels).all(), (name, impl.__name__)
        if n > 1:
            assert (s2[1:] == stats).all(), (name, impl.__name__)


def test_label_order_is_block_raster_not_pixel_raster():
    # SURVEY A7 example: pixels at (row 1, col 0) and (row 0, col 5): block order gives 1, 2
    m = np.zeros((4, 8), np.uint8)
    m[1, 0] = 1
    m[0, 5] = 1
    _, labels, _ = bboxcc_ref.ccl_ref(m)
    assert labels[1, 0] == 1 and labels[0, 5] == 2


def test_bincode_known_bytes():
    # hand-computed from the bincode 1.3 wire format: u64 len, 5 x f32, 4 x Option::None tag
    b = bboxcc_ref.serialize_vec([bboxcc_ref.bbox_new(0., 0., 2., 2.)])
    assert b == struct.pack("<Q", 1) + struct.pack("<5f", 0, 0, 2, 2, 4) + b"\0\0\0\0"
    assert len(b) == 8 + 24
    assert bboxcc_ref.serialize_vec([]) == b"\0" * 8
    # reference's own round-trip test (cova-rs/bbox/src/bbox.rs:124-130)
    assert bboxcc_ref.deserialize_vec(b) == [(0., 0., 2., 2., 4., None, None, None, None)]
    # Some(..) payloads as nvdsbbox writes them (cova-rs/nvdsbbox/src/lib.rs:28-32)
    full = bboxcc_ref.serialize_vec([(1., 2., 3., 4., 12., None, 7, 2, 0.5)])
    assert len(full) == 8 + 20 + 1 + 9 + 5 + 5
    assert bboxcc_ref.deserialize_vec(full)[0][6:] == (7, 2, 0.5)


@pytest.mark.parametrize("thr", [0, 1, 3, 30])
def test_regionprops_python_vs_c(thr):
    for name, m in synth.mask_patterns(45, 80, seed=thr).items():
        py = bboxcc_ref.bboxcc_transform_ref(m, 80, 45, thr)
        assert py == c_oracle.bboxcc(m, thr), name
        boxes = bboxcc_ref.deserialize_vec(py)
        for b in boxes:
            assert b[4] == b[2] * b[3]           # area = bbox w*h, not the pixel count


def test_metapreprocess_window_order_and_drop():
    fr = synth.synth_stream(11, 5, 6, seed=3)
    for T, gamma in [(1, 1), (4, 1), (4, 2), (4, 3), (2, 5)]:
        el = mpr.MetaPreprocessRef(6 * 16, 5 * 16, T, gamma)
        newest = []
        for f in range(fr.shape[0]):
            flow, out = el.transform(fr[f])
            if flow == mpr.FLOW_OK:
                newest.append(f)
                S = 5 * 6 * 4
                for k in range(T):               # row block k = frame f-k (newest first)
                    assert out[k * S:(k + 1) * S] == fr[f - k].tobytes()
        assert newest == mpr.window_newest_indices(fr.shape[0], T, gamma)
        st = mpr.tensorise_stream(fr, T, gamma)
        assert (st == c_oracle.metapreprocess_stream(fr, T, gamma)).all()


def test_metapreprocess_1080p_grid_is_67_rows():
    assert mpr.mb_grid(1920, 1080) == (120, 67)
