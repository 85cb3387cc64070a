"""Frame selection of the `cova` element: host C++ behind the C ABI against the
line-by-line Python restatement (oracle/cova_select_ref.py; parity unpinned - the reference has no test for it).
Host-only code: no GPU needed."""
import numpy as np
import pytest

from cova_b200 import _lib
from cova_b200.elements import CovaSelect, deserialize_vec_full
from oracle import sort_ref
from oracle.bboxcc_ref import serialize_vec
from oracle.cova_select_ref import CovaSelectRef, split_wire
from test_sort import moving_boxes

FRAME_NS = 33_333_333


def drive(sel, frames, gop, lag, is_ref):
    """Encoded frame f arrives `lag` frames before its boxes, like the decoder-side queue of the pipeline."""
    log = []
    n = len(frames)
    for f in range(n + lag):
        if f < n:
            sel.sink_enc(f, f * FRAME_NS, 0 if f % gop == 0 else 1)
        k = f - lag
        if k >= 0:
            boxes = [sort_ref.bbox(*b) for b in frames[k][1]]
            out = sel.sink_mask(boxes, k * FRAME_NS) if is_ref else sel.sink_mask(serialize_vec([b[:5] for b in boxes]), k * FRAME_NS)
            log.append([tuple(int(v) for v in p) for p in out])
    a = sel.on_eos(0) if is_ref else sel.eos(0)
    assert a is None
    b = sel.on_eos(1) if is_ref else sel.eos(1)
    log.append([tuple(int(v) for v in p) for p in b])
    return log


def counters(sel, is_ref):
    if is_ref:
        return sel.dropped, sel.decoded_dependency, sel.decoded_inference
    return tuple(sel.get_property(k) for k in ("dropped", "decoded-dependency", "decoded-inference"))


@pytest.mark.parametrize("seed,props", [
    (0, dict(sort_maxage=6, sort_minhits=10, sort_iou=0.1)),
    (1, dict(sort_maxage=6, sort_minhits=10, sort_iou=0.1, infer_i=True)),
    (2, dict(sort_maxage=8, sort_minhits=12, sort_iou=0.2, alpha=6, beta=3)),
    (3, dict(sort_maxage=6, sort_minhits=10, sort_iou=0.1, port=7000, alpha=4, beta=2, infer_i=True)),
])
def test_selection_matches_oracle(seed, props):
    """sort-minhits >= 10 throughout: the element's window `pts - (maxage + 10 frames)` (SAFETY_BUFFER, imp.rs:127-134)
    only reaches back to the start of a dead track when an active track has lived at least 10 frames; with a smaller
    minhits the reference's own assert!(track_inferenced > 0) fires (next test)."""
    frames = moving_boxes(seed, 700, 8, drop=0.03, clutter=0.2)
    got_sel, ref_sel = CovaSelect(**props), CovaSelectRef(**props)
    got = drive(got_sel, frames, gop=60, lag=5, is_ref=False)
    ref = drive(ref_sel, frames, gop=60, lag=5, is_ref=True)
    assert got == ref
    assert counters(got_sel, False) == counters(ref_sel, True)
    dropped, dep, inf = counters(got_sel, False)
    n_pushed = sum(1 for step in got for p in step if p[0] != CovaSelect.EMPTY_LIST)
    assert inf > 0 and dep > 0 and dropped > 0
    assert n_pushed == dep + inf, "every decoded frame is pushed exactly once"
    assert n_pushed + dropped <= 700  # the remainder are the buffers the reference loses (imp.rs:172-177)
    # dependency frames carry DROPPABLE, inference frames do not; GoP heads keep DISCONT
    flags = {p[0]: p[2] for step in got for p in step if p[0] != CovaSelect.EMPTY_LIST}
    assert sum(1 for f in flags.values() if f & CovaSelect.FLAG_DROPPABLE) == dep
    assert all(bool(fl & CovaSelect.FLAG_DISCONT) == (fid % 60 == 0) for fid, fl in flags.items())
    wire = got_sel.take_wire()
    if props.get("port"):
        a, b = split_wire(wire), split_wire(bytes(ref_sel.wire))
        assert len(a) == len(b) > 0
        for (rs, old, boxes), (rs2, old2, boxes2) in zip(a, b):
            assert (rs, old) == (rs2, old2) and rs == 0
            x, y = deserialize_vec_full(boxes), deserialize_vec_full(boxes2)
            assert [v[5:] for v in x] == [v[5:] for v in y]
            np.testing.assert_allclose(np.array([v[:5] for v in x], dtype=np.float64).reshape(-1, 5),
                                       np.array([v[:5] for v in y], dtype=np.float64).reshape(-1, 5), rtol=2e-3, atol=2e-3)
    else:
        assert wire == b"" and not ref_sel.wire


def test_reference_assertion_is_reported_not_aborted():
    """A track that becomes active after 2 hits and dies maxage+1 frames after its creation starts AFTER the
    selection window: the reference panics on assert!(track_inferenced > 0) (imp.rs:239); the oracle raises, the
    C ABI returns COVA_E_STATE without aborting the process."""
    props = dict(sort_maxage=6, sort_minhits=2, sort_iou=0.1)
    frames = [(f, [(10.0, 10.0, 4.0, 4.0)] if 121 <= f <= 123 else []) for f in range(140)]
    with pytest.raises(AssertionError):
        drive(CovaSelectRef(**props), frames, gop=60, lag=5, is_ref=True)
    with pytest.raises(_lib.CovaError) as e:
        drive(CovaSelect(**props), frames, gop=60, lag=5, is_ref=False)
    assert e.value.code == _lib.E_STATE


def test_properties_and_error_paths():
    s = CovaSelect()
    assert (s.get_property("sort-iou"), s.get_property("sort-maxage"), s.get_property("sort-minhits")) == (pytest.approx(0.1), 30, 30)
    assert (s.get_property("port"), s.get_property("infer-i"), s.get_property("alpha"), s.get_property("beta")) == (0, False, 0, 0)
    with pytest.raises(_lib.CovaError):
        s.set_property("dropped", 1)          # read-only counter
    with pytest.raises(_lib.CovaError):
        s.sink_enc(0, 0, CovaSelect.FLAG_DELTA_UNIT)   # delta unit before any key frame
    s.sink_enc(0, 0, 0)
    with pytest.raises(_lib.CovaError):
        s.sink_mask(b"\x01", 0)               # not bincode
    assert s.sink_mask(serialize_vec([]), 0) == []
    assert s.eos(1) is None
    assert s.eos(0) == [(CovaSelect.EMPTY_LIST, 0, 0, 0)]   # the reference pushes the (empty) list of every GoP at EOS
    assert s.get_property("dropped") == 1


def test_old_gops_are_dropped_after_250_frames():
    """No objects at all: nothing is decoded, every finalized GoP older than 250 frames is dropped
    (imp.rs:246-283), the rest at EOS."""
    s = CovaSelect()
    for f in range(400):
        s.sink_enc(f, f * FRAME_NS, 0 if f % 50 == 0 else 1)
        assert s.sink_mask(serialize_vec([]), f * FRAME_NS) == []
    # at pts 399 the limit is frame 149: GoPs [0,49], [50,99], [100,149] are gone
    assert s.get_property("dropped") == 150
    s.eos(0)
    assert len(s.eos(1)) == 5 and s.get_property("dropped") == 400
    assert s.get_property("decoded-inference") == 0 and s.get_property("decoded-dependency") == 0


def test_eos_pad_order_decides_whether_the_tracker_is_flushed():
    """Only sink_mask_event takes and flushes the tracker (imp.rs:387-390).  When the sink_enc EOS is the one that
    completes the pair (imp.rs:399-424) the GoP lists are still drained, but the tracks alive at EOS never reach the
    aggregator - a reference quirk both implementations must share."""
    props = dict(sort_maxage=6, sort_minhits=10, sort_iou=0.1, port=7000)
    frames = moving_boxes(4, 200, 4, drop=0.0, clutter=0.0)
    wires = {}
    for first_pad in (0, 1):
        got, ref = CovaSelect(**props), CovaSelectRef(**props)
        for f, (_, bx) in enumerate(frames):
            got.sink_enc(f, f * FRAME_NS, 0 if f % 50 == 0 else 1)
            ref.sink_enc(f, f * FRAME_NS, 0 if f % 50 == 0 else 1)
            boxes = [sort_ref.bbox(*b) for b in bx]
            a = got.sink_mask(serialize_vec([b[:5] for b in boxes]), f * FRAME_NS)
            b = ref.sink_mask(boxes, f * FRAME_NS)
            assert [tuple(int(v) for v in p) for p in a] == [tuple(int(v) for v in p) for p in b]
        assert got.eos(first_pad) is None and ref.on_eos(first_pad) is None
        a, b = got.eos(1 - first_pad), ref.on_eos(1 - first_pad)
        assert [tuple(int(v) for v in p) for p in a] == [tuple(int(v) for v in p) for p in b]
        wires[first_pad] = (split_wire(got.take_wire()), split_wire(bytes(ref.wire)))
        assert len(wires[first_pad][0]) == len(wires[first_pad][1])
    # sink_enc first, sink_mask second: flushed -> strictly more Frame records than in the other order
    assert len(wires[0][0]) > len(wires[1][0])
