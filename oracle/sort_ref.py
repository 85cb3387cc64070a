"""CPU oracle of the SORT tracker that follows the blob-detection path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy float32 restatement of (paths relative to the
reference tree)

* ``Sort``               cova-rs/sort/src/lib.rs:25-213
* ``KalmanBoxTracker``   cova-rs/sort/src/tracker/mod.rs:33-153, motion_model.rs:37-66,
                         linear_observation_model.rs:29-56
* ``into_z`` / ``from_x``  cova-rs/sort/src/state.rs:10-27
* ``iou``                cova-rs/bbox/src/bbox.rs:39-56
* ``SortTrackerRef``     cova-rs/gst-plugins/src/sorttracker/imp.rs:214-287

Third-party arithmetic absent from the tree, restated from the published algorithms:
adskalman 0.13.0 (``TransitionModelLinearNoControl::predict``: x' = F x, P' = F P F^T + Q;
``ObservationModel::update`` with ``CovarianceUpdateMethod::JosephForm``: S = H P H^T + R inverted through its
Cholesky factor, K = P H^T S^-1, x = x' + K (z - H x'), P = (I-KH) P (I-KH)^T + K R K^T) and
linear_assignment 0.0.2 @ a992de6 (Kuhn-Munkres on a square matrix), written here as the classic six-step
Munkres so that it is an implementation independent of the shortest-augmenting-path solver in
``cova_b200/csrc/sort_tracker.hpp``.

Pinned by the reference's own known-answer tests (tests/test_sort.py replays them): Hungarian 5x5 / 2x3 / 3x2 /
9x8, IoU matrix, ``match_dets``, ``test_new_sort``, ``test_obeservation_model``, IoU same / quarter / none.
Not pinned: the Kalman arithmetic beyond those cases (no golden track in the tree) and the resolution of ties
between equally good assignments.
"""
from __future__ import annotations

import numpy as np

from .bboxcc_ref import deserialize_vec, serialize_vec

f32 = np.float32


def bbox(left, top, width, height, track_id=None, timestamp=None, class_id=None, confidence=None):
    left, top, width, height = f32(left), f32(top), f32(width), f32(height)
    return [left, top, width, height, f32(width * height), track_id, timestamp, class_id, confidence]


def iou(a, b) -> np.float32:
    """a.iou(b), bbox.rs:39-56."""
    ax2, ay2 = f32(a[0] + a[2]), f32(a[1] + a[3])
    bx2, by2 = f32(b[0] + b[2]), f32(b[1] + b[3])
    xl, yt = max(a[0], b[0]), max(a[1], b[1])
    xr, yb = min(ax2, bx2), min(ay2, by2)
    if xr <= xl or yb <= yt:
        return f32(0.0)
    inter = f32(f32(xr - xl) * f32(yb - yt))
    union = f32(f32(a[4] + b[4]) - inter)
    return f32(inter / union)


def into_z(b):
    """state.rs:11-17."""
    return np.array([b[0] + b[2] / f32(2), b[1] + b[3] / f32(2), b[4], b[2] / b[3]], dtype=f32)


def from_x(x):
    """state.rs:19-27: `top = y - width/2` is the reference's behaviour."""
    w = np.sqrt(f32(x[2] * x[3]), dtype=f32)
    with np.errstate(all="ignore"):
        h = f32(x[2] / w)
    return bbox(x[0] - w / f32(2), x[1] - w / f32(2), w, h)


# ------------------------------------------------------------------------------------------------ Munkres
def munkres_square(cost: np.ndarray):
    """Classic Kuhn-Munkres (row reduction, starring, covering, priming, augmenting, adjusting) on an
    n x n matrix; returns col_of_row."""
    c = np.array(cost, dtype=np.float64)
    n = c.shape[0]
    c -= c.min(axis=1, keepdims=True)
    star = np.zeros((n, n), bool)
    prime = np.zeros((n, n), bool)
    row_cov = np.zeros(n, bool)
    col_cov = np.zeros(n, bool)
    for i in range(n):
        for j in range(n):
            if c[i, j] == 0 and not star[i].any() and not star[:, j].any():
                star[i, j] = True
    while True:
        col_cov = star.any(axis=0)
        if col_cov.sum() >= n:
            break
        row_cov[:] = False
        prime[:] = False
        while True:
            z = np.argwhere((c == 0) & ~row_cov[:, None] & ~col_cov[None, :])
            if len(z) == 0:
                m = c[~row_cov][:, ~col_cov].min()
                c[row_cov] += m
                c[:, ~col_cov] -= m
                continue
            i, j = z[0]
            prime[i, j] = True
            if star[i].any():
                js = int(np.argmax(star[i]))
                row_cov[i] = True
                col_cov[js] = False
                continue
            path = [(i, j)]
            while star[:, path[-1][1]].any():
                r = int(np.argmax(star[:, path[-1][1]]))
                path.append((r, path[-1][1]))
                cc = int(np.argmax(prime[r]))
                path.append((r, cc))
            for k, (r, cc) in enumerate(path):
                star[r, cc] = k % 2 == 0
            break
    return [int(np.argmax(star[i])) for i in range(n)]


def linear_assignment(cost: np.ndarray):
    """lib.rs:25-56.  cost: float32 [n_trk, n_det]; returns sorted (tracker, detection) pairs."""
    n_trk, n_det = cost.shape
    if n_trk == 0 or n_det == 0:
        return []
    n = max(n_trk, n_det)
    sq = np.zeros((n, n), dtype=f32)
    sq[:n_trk, :n_det] = cost
    cols = munkres_square(sq)
    return [(i, j) for i, j in enumerate(cols) if i < n_trk and j < n_det and cost[i, j] != f32(2.0)]


# ------------------------------------------------------------------------------------------------ Kalman
F = np.eye(7, dtype=f32)
F[0, 4] = F[1, 5] = F[2, 6] = 1
Q = np.diag(np.array([1, 1, 1, 1, 0.01, 0.01, 0.0001], dtype=f32))
H = np.zeros((4, 7), dtype=f32)
H[:4, :4] = np.eye(4, dtype=f32)
R = np.diag(np.array([1, 1, 10, 10], dtype=f32))


class KalmanBoxTracker:
    def __init__(self, tid: int, b, start: int):
        self.id, self.start, self.last_match = tid, start, start
        self.seen_ts: list[int] = []
        self.active = False
        self.history: list = []
        self.hits = self.time_since_update = self.hit_streaks = self.age = 0
        self.x = np.concatenate([into_z(b), np.zeros(3, dtype=f32)]).astype(f32)
        self.P = np.diag(np.array([10, 10, 10, 10, 1e4, 1e4, 1e4], dtype=f32))
        self.prior = None

    def predict(self, ts: int):
        if self.x[6] + self.x[2] <= 0:
            self.x[6] = f32(0)
        xp = (F @ self.x).astype(f32)
        Pp = ((F @ self.P) @ F.T + Q).astype(f32)
        self.prior = (xp, Pp)
        b = from_x(xp)
        b[5], b[6] = self.id, ts
        self.age += 1
        self.time_since_update += 1
        self.history.append(b)
        return b

    def update(self, det):
        if det is None:
            self.hit_streaks = 0
            return
        self.hits += 1
        self.hit_streaks += 1
        if self.hit_streaks >= 5:
            self.time_since_update = 0
            self.last_match = det[6]
        xp, Pp = self.prior
        S = (H @ Pp @ H.T + R).astype(f32)
        L = np.linalg.cholesky(S)  # raises LinAlgError where adskalman returns CovarianceNotPositiveSemiDefinite
        Li = np.linalg.inv(L).astype(f32)
        Sinv = (Li.T @ Li).astype(f32)
        K = (Pp @ H.T @ Sinv).astype(f32)
        self.x = (xp + K @ (into_z(det) - H @ xp)).astype(f32)
        A = (np.eye(7, dtype=f32) - K @ H).astype(f32)
        self.P = (A @ Pp @ A.T + K @ R @ K.T).astype(f32)
        self.history[-1][7], self.history[-1][8] = det[7], det[8]

    def is_seen(self) -> bool:
        return any(self.start <= ts <= self.last_match for ts in self.seen_ts)


class Sort:
    def __init__(self, max_age=3, min_hits=3, iou_threshold=0.2):
        """Defaults = `impl Default for Sort` (lib.rs:216-225)."""
        self.max_age, self.min_hits, self.iou_threshold = max_age, min_hits, f32(iou_threshold)
        self.trackers: list[KalmanBoxTracker] = []
        self.frame_count = self.id_counter = 0

    @staticmethod
    def generate_iou_matrix(preds, dets) -> np.ndarray:
        m = np.zeros((len(preds), len(dets)), dtype=f32)
        for i, p in enumerate(preds):
            for j, d in enumerate(dets):
                m[i, j] = -iou(d, p)
        return m

    def match_dets(self, preds, dets):
        if not preds or not dets:
            return []
        cost = self.generate_iou_matrix(preds, dets)
        for i, t in enumerate(self.trackers):
            cost[i] += f32(1.0 if t.active else 2.0)
        out = []
        for i, j in linear_assignment(cost):
            thr = f32((1.0 if self.trackers[i].active else 2.0)) - self.iou_threshold
            if cost[i, j] <= f32(thr):
                out.append((i, j))
        return out

    def update(self, dets, pts: int):
        self.frame_count += 1
        dets = [list(d) for d in dets]
        preds = [t.predict(pts) for t in self.trackers]
        matches = self.match_dets(preds, dets)
        matched_dets = {j for _, j in matches}
        by_trk = dict(matches)
        for i, t in enumerate(self.trackers):
            det = None
            if i in by_trk:
                dets[by_trk[i]][6] = pts
                det = dets[by_trk[i]]
            t.update(det)
        for t in self.trackers:
            if not t.active and t.hit_streaks >= self.min_hits:
                t.active = True
        dead, alive = [], []
        for t in self.trackers:
            if t.time_since_update <= self.max_age:
                alive.append(t)
            elif t.active:
                t.history = t.history[: len(t.history) - t.time_since_update]
                dead.append(t)
        self.trackers = alive
        for j, d in enumerate(dets):
            if j not in matched_dets:
                self.trackers.append(KalmanBoxTracker(self.id_counter, d, pts))
                self.id_counter += 1
        return dead

    def mark_seen(self, ts: int):
        for t in self.trackers:
            t.seen_ts.append(ts)

    def oldest_start(self) -> int:
        return min([t.start for t in self.trackers], default=2**64 - 1)

    def finalize(self):
        out = [t for t in self.trackers if t.active and len(t.history) > self.min_hits]
        self.trackers = [t for t in self.trackers if not t.active]
        return out


class SortTrackerRef:
    """The `sorttracker` element (sorttracker/imp.rs): bincode in, bincode of dead tracks out."""

    def __init__(self, iou_threshold=0.1, maxage=30, minhits=30):
        self.sort = Sort(maxage, minhits, iou_threshold)

    @staticmethod
    def _ser(tracks) -> bytes:
        return serialize_vec([tuple(b) for t in tracks for b in t.history])

    def transform(self, buf: bytes, pts: int) -> bytes:
        dets = [list(b) for b in deserialize_vec(buf)]
        for d in dets:
            d[:5] = [f32(v) for v in d[:5]]
        return self._ser(self.sort.update(dets, pts))

    def eos(self) -> bytes:
        return self._ser(self.sort.finalize())
