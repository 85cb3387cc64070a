"""SORT tracker behind the path: the reference's own known-answer tests
(cova-rs/sort/src/lib.rs:230-408, cova-rs/bbox/src/bbox.rs:95-122) replayed on the oracle restatement AND on the
host C++ behind the C ABI, then oracle-vs-C++ on seeded box streams.  Host-only code: no GPU needed."""
import ctypes

import numpy as np
import pytest

from cova_b200 import _lib
from cova_b200.elements import SortTracker, deserialize_vec_full
from oracle import sort_ref
from oracle.bboxcc_ref import serialize_vec

f32 = np.float32


# ----------------------------------------------------------------------------------------------- ABI helpers
def abi_linear_assignment(cost):
    cost = np.ascontiguousarray(cost, dtype=f32)
    pairs = np.zeros((max(1, min(cost.shape)), 2), dtype=np.int32)
    n = ctypes.c_uint32()
    _lib.check(_lib.load().cova_sort_linear_assignment(cost.ctypes.data, cost.shape[0], cost.shape[1], pairs.ctypes.data, ctypes.byref(n)))
    return [tuple(int(v) for v in p) for p in pairs[: n.value]]


def abi_iou_matrix(preds, dets):
    p, d = np.ascontiguousarray(preds, dtype=f32), np.ascontiguousarray(dets, dtype=f32)
    out = np.zeros((len(p), len(d)), dtype=f32)
    _lib.check(_lib.load().cova_sort_iou_matrix(p.ctypes.data, len(p), d.ctypes.data, len(d), out.ctypes.data))
    return out


def abi_match_dets(preds, active, dets, thr):
    p, d = np.ascontiguousarray(preds, dtype=f32), np.ascontiguousarray(dets, dtype=f32)
    a = np.ascontiguousarray(active, dtype=np.uint8)
    pairs = np.zeros((max(1, min(len(p), len(d))), 2), dtype=np.int32)
    n = ctypes.c_uint32()
    _lib.check(_lib.load().cova_sort_match_dets(p.ctypes.data, a.ctypes.data, len(p), d.ctypes.data, len(d), thr, pairs.ctypes.data, ctypes.byref(n)))
    return [tuple(int(v) for v in q) for q in pairs[: n.value]]


def col_major(rows, cols, values, offset):
    """DMatrix::from_vec is column-major: the reference's literals read as the transpose."""
    return (np.array(values, dtype=f32).reshape(cols, rows).T + f32(offset)).astype(f32)


# the four Hungarian known-answer tests, sort/src/lib.rs:281-380
HUNGARIAN = [
    (5, 5, 2.0, [-1, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [(0, 0), (1, 1), (3, 2)]),
    (2, 3, 1.0, [-1, 0, 0, 0, 0, -1], [(0, 0), (1, 2)]),
    (3, 2, 1.0, [-1, 0, 0, 0, 0, -1], [(0, 0), (2, 1)]),
    (9, 8, 1.0, [-1, 0, 0, 0, 0, 0, 0, 0, 0,
                 0, -1, 0, 0, 0, 0, 0, 0, 0,
                 0, 0, -1, 0, 0, 0, 0, 0, 0,
                 0, 0, 0, 0, -1, 0, 0, 0, 0,
                 0, 0, 0, 0, 0, -1, 0, 0, 0,
                 0, 0, 0, 0, 0, 0, -1, 0, 0,
                 0, 0, 0, 0, 0, 0, 0, -1, 0,
                 0, 0, 0, 0, 0, 0, 0, 0, -1],
     [(0, 0), (1, 1), (2, 2), (4, 3), (5, 4), (6, 5), (7, 6), (8, 7)]),
]


@pytest.mark.parametrize("impl", ["oracle", "abi"])
@pytest.mark.parametrize("rows,cols,offset,values,expected", HUNGARIAN)
def test_reference_linear_assignment_cases(impl, rows, cols, offset, values, expected):
    cost = col_major(rows, cols, values, offset)
    got = sort_ref.linear_assignment(cost) if impl == "oracle" else abi_linear_assignment(cost)
    assert sorted(got) == sorted(expected)


@pytest.mark.parametrize("impl", ["oracle", "abi"])
def test_reference_iou_cases(impl):
    """bbox.rs:99-122 and lib.rs:270-278."""
    def one(a, b):
        if impl == "oracle":
            return sort_ref.iou(sort_ref.bbox(*a), sort_ref.bbox(*b))
        return -abi_iou_matrix([b], [a])[0, 0]
    assert one((0, 0, 2, 2), (0, 0, 2, 2)) == f32(1.0)
    assert one((0, 0, 2, 2), (1, 1, 2, 2)) == f32(1.0) / f32(7.0)
    assert one((0, 0, 2, 2), (2, 2, 2, 2)) == f32(0.0)
    dets, preds = [(0, 0, 2, 2), (1, 1, 1, 1)], [(1, 1, 1, 1)]
    m = (sort_ref.Sort.generate_iou_matrix([sort_ref.bbox(*p) for p in preds], [sort_ref.bbox(*d) for d in dets])
         if impl == "oracle" else abi_iou_matrix(preds, dets))
    assert m.shape == (1, 2) and m[0, 0] == f32(-0.25) and m[0, 1] == f32(-1.0)


@pytest.mark.parametrize("impl", ["oracle", "abi"])
def test_reference_match_dets(impl):
    """lib.rs:382-407: two fresh (inactive) trackers, zero velocity, three shifted detections -> [(1, 0)]."""
    first = [(0, 0, 4, 4), (1, 1, 4, 4)]
    second = [(1, 1, 4, 4), (2, 2, 4, 4), (3, 3, 4, 4)]
    if impl == "oracle":
        s = sort_ref.Sort()
        s.update([sort_ref.bbox(*b) for b in first], 0)
        assert len(s.trackers) == 2
        preds = [t.predict(0) for t in s.trackers]
        got = s.match_dets(preds, [sort_ref.bbox(*b) for b in second])
    else:
        got = abi_match_dets(first, [0, 0], second, 0.2)
    assert got == [(1, 0)]


def _run_abi(frames, **props):
    """frames: list of (pts, [(l,t,w,h), ...]); returns (per-frame outputs, eos output) as decoded tuples."""
    st = SortTracker(**props)
    st.set_caps(80, 45)
    outs = [deserialize_vec_full(st.transform(serialize_vec([sort_ref.bbox(*b)[:5] for b in boxes]), pts)) for pts, boxes in frames]
    return outs, deserialize_vec_full(st.eos()), st


def _run_oracle(frames, iou_threshold=0.1, maxage=30, minhits=30):
    st = sort_ref.SortTrackerRef(iou_threshold, maxage, minhits)
    outs = [deserialize_vec_full(st.transform(serialize_vec([sort_ref.bbox(*b)[:5] for b in boxes]), pts)) for pts, boxes in frames]
    return outs, deserialize_vec_full(st.eos()), st


def test_reference_new_sort_and_observation_model():
    """lib.rs:231-268: after one update two trackers exist whose state equals the detections; predict() with zero
    velocity returns the detection itself (square boxes: the `top = y - width/2` quirk is invisible)."""
    dets = [(0, 0, 2, 2), (1, 1, 2, 2)]
    s = sort_ref.Sort()
    s.update([sort_ref.bbox(*b) for b in dets], 0)
    assert s.frame_count == 1 and len(s.trackers) == 2
    for t, d in zip(s.trackers, dets):
        assert [float(v) for v in sort_ref.from_x(t.x)[:5]] == [float(v) for v in sort_ref.bbox(*d)[:5]]
        assert [float(v) for v in t.predict(0)[:5]] == [float(v) for v in sort_ref.bbox(*d)[:5]]
    outs, fin, st = _run_abi([(0, dets)], maxage=3, minhits=3, iou_threshold=0.2)
    assert outs == [[]] and fin == [] and st.n_tracks() == (2, 0)


def test_sorttracker_properties_and_errors():
    st = SortTracker()
    assert (st.get_property("iou-threshold"), st.get_property("maxage"), st.get_property("minhits")) == (pytest.approx(0.1), 30, 30)
    st.set_property("maxage", 7)
    assert st.get_property("maxage") == 7
    with pytest.raises(KeyError):
        st.set_property("timestep", 1)
    with pytest.raises(_lib.CovaError):
        st.set_property("iou-threshold", 1.5)
    with pytest.raises(_lib.CovaError):  # transform before caps: the reference unwraps a None Sort
        st.transform(serialize_vec([]), 0)
    st.set_caps(80, 45)
    with pytest.raises(_lib.CovaError):  # truncated bincode
        st.transform(serialize_vec([sort_ref.bbox(0, 0, 1, 1)[:5]])[:-1], 0)
    assert st.transform(serialize_vec([]), 0) == serialize_vec([])


def moving_boxes(seed, n_frames, n_obj, noise=0.3, drop=0.1, clutter=0.5):
    """Objects moving at constant velocity over an 80x45 grid, jittered, sometimes missed, plus clutter."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform([5, 5], [60, 30], size=(n_obj, 2))
    vel = rng.uniform(-0.6, 0.6, size=(n_obj, 2))
    size = rng.uniform(3, 9, size=(n_obj, 2))
    born = rng.integers(0, n_frames // 2, n_obj)
    dies = born + rng.integers(8, n_frames, n_obj)
    frames = []
    for f in range(n_frames):
        boxes = []
        for k in range(n_obj):
            if born[k] <= f < dies[k] and rng.random() > drop:
                p = pos[k] + vel[k] * (f - born[k]) + rng.normal(0, noise, 2)
                s = size[k] + rng.normal(0, noise, 2)
                boxes.append((round(float(p[0]), 2), round(float(p[1]), 2), max(1.0, round(float(s[0]), 2)), max(1.0, round(float(s[1]), 2))))
        for _ in range(rng.poisson(clutter)):
            boxes.append((float(rng.integers(0, 78)), float(rng.integers(0, 43)), float(rng.integers(1, 3)), float(rng.integers(1, 3))))
        frames.append((f * 33_333_333, boxes))
    return frames


def assert_same_tracks(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x[5:7] == y[5:7], "track id / timestamp differ"
        assert x[7:] == y[7:]
        np.testing.assert_allclose(np.array(x[:5], dtype=np.float64), np.array(y[:5], dtype=np.float64), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("seed,params", [(0, dict(iou_threshold=0.1, maxage=30, minhits=30)),
                                         (1, dict(iou_threshold=0.2, maxage=6, minhits=3)),
                                         (2, dict(iou_threshold=0.3, maxage=10, minhits=5)),
                                         (3, dict(iou_threshold=0.1, maxage=5, minhits=0))])
def test_cpp_tracker_matches_oracle_on_box_streams(seed, params):
    """Discrete outcomes (which tracks die on which frame, ids, timestamps, history lengths) identical; box
    coordinates within f32 round-off of the numpy restatement (tolerance 2e-3 abs/rel: the two implementations
    order the 7x7 products differently)."""
    frames = moving_boxes(seed, 120, 6)
    got, got_fin, st = _run_abi(frames, **params)
    ref, ref_fin, rst = _run_oracle(frames, **params)
    for g, r in zip(got, ref):
        assert_same_tracks(g, r)
    assert_same_tracks(got_fin, ref_fin)
    assert sum(len(g) for g in got) + len(got_fin) > 50, "the scenario should produce dead or final tracks"
    assert st.n_tracks() == (len(rst.sort.trackers), sum(t.active for t in rst.sort.trackers))


def test_dead_track_history_is_trimmed_and_ordered():
    """A single object seen for 12 frames then gone: with maxage 5 it dies on frame 12+5 and its history holds only
    the entries up to the last matched frame (trim_dead_history, tracker/mod.rs:146-153).  time_since_update is
    reset only from the 5th consecutive hit on (tracker/mod.rs:77-80), so with maxage < 5 a track dies before it
    can ever be refreshed: with the crate's Default (maxage 3) the same object dies on frame 4 with an empty history."""
    frames = [(i, [(10 + i, 10, 4, 4)] if i < 12 else []) for i in range(20)]
    outs, fin, _ = _run_abi(frames, maxage=5, minhits=3, iou_threshold=0.1)
    died = [i for i, o in enumerate(outs) if o]
    assert died == [17] and fin == []
    hist = outs[17]
    assert [b[6] for b in hist] == list(range(1, 12))  # first predict is at frame 1; frames 12..17 trimmed
    assert all(b[5] == 0 for b in hist)
    ref, _, _ = _run_oracle(frames, maxage=5, minhits=3, iou_threshold=0.1)
    assert_same_tracks(hist, ref[17])
    outs3, _, st3 = _run_abi(frames[:5], maxage=3, minhits=3, iou_threshold=0.1)
    ref3, _, _ = _run_oracle(frames[:5], maxage=3, minhits=3, iou_threshold=0.1)
    assert outs3 == ref3 == [[], [], [], [], []] and st3.n_tracks() == (0, 0)  # died at frame 4; its box was matched, so no new track either


def test_large_assignment_problem_is_optimal():
    """200 x 180 random costs: the C++ solver's total cost equals scipy's optimum."""
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(5)
    cost = rng.uniform(0.0, 1.9, size=(200, 180)).astype(f32)
    got = abi_linear_assignment(cost)
    r, c = linear_sum_assignment(cost.astype(np.float64))
    assert len(got) == 180
    assert abs(sum(float(cost[i, j]) for i, j in got) - float(cost[r, c].sum())) < 1e-6
