// Host side of the step AFTER the blob-detection path: the SORT tracker that
// consumes the per-frame bincode(Vec<Bbox>) blobs the GPU path emits.  Plain C++17, no CUDA: the work is a
// few 7x7 matrix products and one assignment problem per frame, fed straight from the pinned box buffers.
//
// Written from the behaviour of (paths relative to the reference tree)
//   cova-rs/sort/src/lib.rs:25-187            Sort::update / match_dets / linear_assignment / finalize
//   cova-rs/sort/src/tracker/mod.rs:33-141    KalmanBoxTracker (predict / update / trim_dead_history / is_seen)
//   cova-rs/sort/src/tracker/motion_model.rs, linear_observation_model.rs   F, Q, H, R
//   cova-rs/sort/src/state.rs:10-27           Bbox <-> (x, y, s, r), including the `top = y - width/2` quirk
//   cova-rs/bbox/src/bbox.rs:3-56             Bbox, iou, bincode layout (Option tags)
// Third-party arithmetic that is not in the tree: adskalman 0.13.0 (predict: F x, F P F^T + Q; update: Cholesky
// inverse of S, Joseph-form covariance), linear_assignment 0.0.2 @a992de6 (Munkres on a zero-padded square
// matrix).  Both are restated from their published algorithms; an assignment problem with several optimal
// solutions may be resolved differently from the crate ("f2").
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <utility>
#include <vector>

namespace cova {
namespace host {

struct Bbox {
    float left = 0, top = 0, width = 0, height = 0, area = 0;
    bool has_track = false, has_ts = false, has_class = false, has_conf = false;
    uint64_t track_id = 0, timestamp = 0;
    uint32_t class_id = 0;
    float confidence = 0;

    static Bbox make(float l, float t, float w, float h) {  // Bbox::new, bbox.rs:17-29
        Bbox b;
        b.left = l, b.top = t, b.width = w, b.height = h, b.area = w * h;
        return b;
    }
    // bbox.rs:39-56; `this` is the detection, `o` the prediction at the call site (lib.rs:88-89)
    float iou(const Bbox &o) const {
        float sx2 = left + width, sy2 = top + height, tx2 = o.left + o.width, ty2 = o.top + o.height;
        float xl = fmaxf(left, o.left), yt = fmaxf(top, o.top), xr = fminf(sx2, tx2), yb = fminf(sy2, ty2);
        if (xr <= xl || yb <= yt) return 0.f;
        float inter = (xr - xl) * (yb - yt);
        float uni = area + o.area - inter;
        return inter / uni;
    }
    size_t wire_size() const { return 20 + 4 + (has_track ? 8 : 0) + (has_ts ? 8 : 0) + (has_class ? 4 : 0) + (has_conf ? 4 : 0); }
};

// ---- bincode 1.3.3, default options: fixed-width little-endian integers, u64 length prefix, Option = u8 tag ----
inline size_t boxes_wire_size(const std::vector<Bbox> &v) {
    size_t n = 8;
    for (const Bbox &b : v) n += b.wire_size();
    return n;
}
template <typename T> inline void put(uint8_t *&p, T v) { memcpy(p, &v, sizeof(T)), p += sizeof(T); }
inline uint8_t *encode_boxes_into(const std::vector<Bbox> &v, uint8_t *p) {
    put<uint64_t>(p, v.size());
    for (const Bbox &b : v) {
        put(p, b.left), put(p, b.top), put(p, b.width), put(p, b.height), put(p, b.area);
        put<uint8_t>(p, b.has_track); if (b.has_track) put(p, b.track_id);
        put<uint8_t>(p, b.has_ts);    if (b.has_ts) put(p, b.timestamp);
        put<uint8_t>(p, b.has_class); if (b.has_class) put(p, b.class_id);
        put<uint8_t>(p, b.has_conf);  if (b.has_conf) put(p, b.confidence);
    }
    return p;
}
// false on truncated input, an Option tag other than 0/1 or trailing garbage in the length (bincode errors)
inline bool decode_boxes(const uint8_t *p, size_t len, std::vector<Bbox> &out) {
    const uint8_t *end = p + len;
    auto get = [&](void *dst, size_t n) { if ((size_t)(end - p) < n) return false; memcpy(dst, p, n); p += n; return true; };
    uint64_t n;
    if (!get(&n, 8)) return false;
    if (n > len / 24) return false;  // every box takes at least 24 bytes
    out.clear();
    out.reserve((size_t)n);
    for (uint64_t i = 0; i < n; i++) {
        Bbox b;
        uint8_t tag;
        if (!get(&b.left, 4) || !get(&b.top, 4) || !get(&b.width, 4) || !get(&b.height, 4) || !get(&b.area, 4)) return false;
        if (!get(&tag, 1) || tag > 1) return false; b.has_track = tag; if (tag && !get(&b.track_id, 8)) return false;
        if (!get(&tag, 1) || tag > 1) return false; b.has_ts = tag;    if (tag && !get(&b.timestamp, 8)) return false;
        if (!get(&tag, 1) || tag > 1) return false; b.has_class = tag; if (tag && !get(&b.class_id, 4)) return false;
        if (!get(&tag, 1) || tag > 1) return false; b.has_conf = tag;  if (tag && !get(&b.confidence, 4)) return false;
        out.push_back(b);
    }
    return true;
}

// ---- assignment problem ------------------------------------------------------------------------------------
// Minimum-cost perfect matching of an n x n matrix (row-major), shortest augmenting paths with potentials,
// O(n^3).  Costs are f32 values widened to double so that the optimum is not disturbed by rounding.
inline void solve_square_assignment(const std::vector<double> &c, int n, std::vector<int> &col_of_row) {
    const double INF = std::numeric_limits<double>::infinity();
    std::vector<double> u(n + 1, 0.0), v(n + 1, 0.0), minv(n + 1);
    std::vector<int> p(n + 1, 0), way(n + 1, 0);
    std::vector<char> used(n + 1);
    for (int i = 1; i <= n; i++) {
        p[0] = i;
        int j0 = 0;
        std::fill(minv.begin(), minv.end(), INF);
        std::fill(used.begin(), used.end(), 0);
        do {
            used[j0] = 1;
            int i0 = p[j0], j1 = 0;
            double delta = INF;
            const double *row = &c[(size_t)(i0 - 1) * n];
            for (int j = 1; j <= n; j++) {
                if (used[j]) continue;
                double cur = row[j - 1] - u[i0] - v[j];
                if (cur < minv[j]) minv[j] = cur, way[j] = j0;
                if (minv[j] < delta) delta = minv[j], j1 = j;
            }
            for (int j = 0; j <= n; j++) {
                if (used[j]) u[p[j]] += delta, v[j] -= delta;
                else minv[j] -= delta;
            }
            j0 = j1;
        } while (p[j0] != 0);
        do {
            int j1 = way[j0];
            p[j0] = p[j1];
            j0 = j1;
        } while (j0);
    }
    col_of_row.assign(n, -1);
    for (int j = 1; j <= n; j++) col_of_row[p[j] - 1] = j - 1;
}

// lib.rs:25-56: pad to a square with zeros, solve, drop padded rows/columns and pairs whose cost == 2.0
// cost: row-major [n_trk][n_det].  Result sorted by tracker index (the crate returns a HashSet: no order).
inline std::vector<std::pair<int, int>> linear_assignment(const std::vector<float> &cost, int n_trk, int n_det) {
    std::vector<std::pair<int, int>> out;
    if (n_trk <= 0 || n_det <= 0) return out;
    int n = std::max(n_trk, n_det);
    std::vector<double> sq((size_t)n * n, 0.0);
    for (int i = 0; i < n_trk; i++)
        for (int j = 0; j < n_det; j++) sq[(size_t)i * n + j] = cost[(size_t)i * n_det + j];
    std::vector<int> col;
    solve_square_assignment(sq, n, col);
    for (int i = 0; i < n_trk; i++) {
        int j = col[i];
        if (j < n_det && cost[(size_t)i * n_det + j] != 2.0f) out.emplace_back(i, j);
    }
    return out;
}

// ---- Kalman box tracker ---------------------------------------------------------------------------------------
struct Mat7 { float m[7][7]; };

struct KalmanBoxTracker {
    uint64_t id, start, last_match;
    std::vector<uint64_t> seen_ts;
    bool active = false;
    std::vector<Bbox> history;
    uint64_t hits = 0, time_since_update = 0, hit_streaks = 0, age = 0;
    float x[7];      // previous_estimate: u, v, s, r, du, dv, ds
    Mat7 P;
    float xp[7];     // prior (valid after predict)
    Mat7 Pp;
    bool has_prior = false;

    static void to_z(const Bbox &b, float z[4]) {  // state.rs:11-17
        z[0] = b.left + b.width / 2.f, z[1] = b.top + b.height / 2.f, z[2] = b.area, z[3] = b.width / b.height;
    }
    static Bbox from_x(const float *s) {  // state.rs:19-27 (top uses width: reproduced on purpose)
        float w = sqrtf(s[2] * s[3]);
        float h = s[2] / w;
        return Bbox::make(s[0] - w / 2.f, s[1] - w / 2.f, w, h);
    }

    KalmanBoxTracker(uint64_t id_, const Bbox &b, uint64_t start_) : id(id_), start(start_), last_match(start_) {  // mod.rs:33-71
        float z[4];
        to_z(b, z);
        for (int i = 0; i < 7; i++) x[i] = i < 4 ? z[i] : 0.f;
        memset(&P, 0, sizeof(P));
        for (int i = 0; i < 7; i++) P.m[i][i] = i < 4 ? 10.f : 10000.f;
    }

    // mod.rs:110-126 + adskalman predict: x' = F x, P' = (F P) F^T + Q with F = I + e0 e4^T + e1 e5^T + e2 e6^T
    const Bbox &predict(uint64_t ts) {
        if (x[6] + x[2] <= 0.f) x[6] = 0.f;
        for (int i = 0; i < 7; i++) xp[i] = x[i];
        xp[0] += x[4], xp[1] += x[5], xp[2] += x[6];
        Mat7 FP = P;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 7; c++) FP.m[r][c] = P.m[r][c] + P.m[r + 4][c];
        Pp = FP;
        for (int r = 0; r < 7; r++)
            for (int c = 0; c < 3; c++) Pp.m[r][c] = FP.m[r][c] + FP.m[r][c + 4];
        static const float q[7] = {1.f, 1.f, 1.f, 1.f, 0.01f, 0.01f, 0.0001f};
        for (int i = 0; i < 7; i++) Pp.m[i][i] += q[i];
        has_prior = true;
        Bbox b = from_x(xp);
        b.has_track = true, b.track_id = id, b.has_ts = true, b.timestamp = ts;
        age++, time_since_update++;
        history.push_back(b);
        return history.back();
    }

    // mod.rs:73-108 + adskalman ObservationModel::update (H = [I4 | 0], R = diag(1,1,10,10), Joseph form).
    // false: S is not positive definite (adskalman returns an error, the element panics on it)
    bool update(const Bbox *det) {
        if (!det) { hit_streaks = 0; return true; }
        hits++, hit_streaks++;
        if (hit_streaks >= 5) time_since_update = 0, last_match = det->timestamp;
        if (!has_prior) return false;
        static const float Rd[4] = {1.f, 1.f, 10.f, 10.f};
        float z[4];
        to_z(*det, z);
        float S[4][4], L[4][4] = {}, Li[4][4] = {}, Si[4][4];
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) S[i][j] = Pp.m[i][j] + (i == j ? Rd[i] : 0.f);
        for (int j = 0; j < 4; j++) {  // Cholesky S = L L^T
            float d = S[j][j];
            for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
            if (!(d > 0.f)) return false;
            L[j][j] = sqrtf(d);
            for (int i = j + 1; i < 4; i++) {
                float s = S[i][j];
                for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
                L[i][j] = s / L[j][j];
            }
        }
        for (int c = 0; c < 4; c++)  // Li = L^-1 (forward substitution)
            for (int i = 0; i < 4; i++) {
                float s = i == c ? 1.f : 0.f;
                for (int k = 0; k < i; k++) s -= L[i][k] * Li[k][c];
                Li[i][c] = s / L[i][i];
            }
        for (int i = 0; i < 4; i++)  // S^-1 = Li^T Li
            for (int j = 0; j < 4; j++) {
                float s = 0.f;
                for (int k = 0; k < 4; k++) s += Li[k][i] * Li[k][j];
                Si[i][j] = s;
            }
        float K[7][4];
        for (int i = 0; i < 7; i++)
            for (int j = 0; j < 4; j++) {
                float s = 0.f;
                for (int k = 0; k < 4; k++) s += Pp.m[i][k] * Si[k][j];
                K[i][j] = s;
            }
        float innov[4];
        for (int j = 0; j < 4; j++) innov[j] = z[j] - xp[j];
        for (int i = 0; i < 7; i++) {
            float s = 0.f;
            for (int j = 0; j < 4; j++) s += K[i][j] * innov[j];
            x[i] = xp[i] + s;
        }
        float A[7][7], AP[7][7];  // A = I - K H
        for (int i = 0; i < 7; i++)
            for (int j = 0; j < 7; j++) A[i][j] = (i == j ? 1.f : 0.f) - (j < 4 ? K[i][j] : 0.f);
        for (int i = 0; i < 7; i++)
            for (int j = 0; j < 7; j++) {
                float s = 0.f;
                for (int k = 0; k < 7; k++) s += A[i][k] * Pp.m[k][j];
                AP[i][j] = s;
            }
        for (int i = 0; i < 7; i++)
            for (int j = 0; j < 7; j++) {
                float s = 0.f;
                for (int k = 0; k < 7; k++) s += AP[i][k] * A[j][k];
                float kr = 0.f;
                for (int k = 0; k < 4; k++) kr += K[i][k] * Rd[k] * K[j][k];
                P.m[i][j] = s + kr;
            }
        Bbox &last = history.back();
        last.has_class = det->has_class, last.class_id = det->class_id;
        last.has_conf = det->has_conf, last.confidence = det->confidence;
        return true;
    }

    bool should_live(uint64_t max_age) const { return time_since_update <= max_age; }
    void check_activate(uint64_t min_hits) { if (!active && hit_streaks >= min_hits) active = true; }
    bool is_seen() const {  // mod.rs:140-144
        for (uint64_t ts : seen_ts) if (start <= ts && last_match >= ts) return true;
        return false;
    }
    void trim_dead_history() { history.resize(history.size() - (size_t)time_since_update); }  // mod.rs:146-153
};

// ---- Sort ---------------------------------------------------------------------------------------------------
struct Sort {
    uint64_t max_age, min_hits;
    float iou_threshold;
    std::vector<KalmanBoxTracker> trackers;
    uint64_t frame_count = 0, id_counter = 0;

    Sort(uint64_t max_age_, uint64_t min_hits_, float iou) : max_age(max_age_), min_hits(min_hits_), iou_threshold(iou) {}

    // lib.rs:81-93: cost[i][j] = -iou(det j, pred i)
    static std::vector<float> iou_cost(const std::vector<Bbox> &preds, const std::vector<Bbox> &dets) {
        std::vector<float> c(preds.size() * dets.size());
        for (size_t i = 0; i < preds.size(); i++)
            for (size_t j = 0; j < dets.size(); j++) c[i * dets.size() + j] = -dets[j].iou(preds[i]);
        return c;
    }

    // lib.rs:98-134 (active[i] = trackers[i].active)
    static std::vector<std::pair<int, int>> match_dets(const std::vector<Bbox> &preds, const std::vector<char> &active,
                                                       const std::vector<Bbox> &dets, float iou_threshold) {
        std::vector<std::pair<int, int>> out;
        const int np = (int)preds.size(), nd = (int)dets.size();
        if (!np || !nd) return out;
        std::vector<float> cost = iou_cost(preds, dets);
        for (int i = 0; i < np; i++) {
            const float w = active[i] ? 1.f : 2.f;
            for (int j = 0; j < nd; j++) cost[(size_t)i * nd + j] += w;
        }
        for (auto &m : linear_assignment(cost, np, nd)) {
            const float thr = active[m.first] ? 1.f - iou_threshold : 2.f - iou_threshold;
            if (cost[(size_t)m.first * nd + m.second] <= thr) out.push_back(m);
        }
        return out;
    }

    // lib.rs:137-187.  false when a Kalman update fails (the reference propagates the error and panics)
    bool update(std::vector<Bbox> dets, uint64_t pts, std::vector<KalmanBoxTracker> &dead) {
        frame_count++;
        std::vector<Bbox> preds;
        preds.reserve(trackers.size());
        for (auto &t : trackers) preds.push_back(t.predict(pts));
        std::vector<char> active(trackers.size());
        for (size_t i = 0; i < trackers.size(); i++) active[i] = trackers[i].active;
        auto matches = match_dets(preds, active, dets, iou_threshold);
        std::vector<int> det_of_trk(trackers.size(), -1);
        std::vector<char> det_used(dets.size(), 0);
        for (auto &m : matches) det_of_trk[m.first] = m.second, det_used[m.second] = 1;
        for (size_t i = 0; i < trackers.size(); i++) {
            const Bbox *d = nullptr;
            if (det_of_trk[i] >= 0) {
                dets[det_of_trk[i]].has_ts = true, dets[det_of_trk[i]].timestamp = pts;
                d = &dets[det_of_trk[i]];
            }
            if (!trackers[i].update(d)) return false;
        }
        for (auto &t : trackers) t.check_activate(min_hits);
        dead.clear();
        std::vector<KalmanBoxTracker> alive;
        alive.reserve(trackers.size() + dets.size());
        for (auto &t : trackers) {
            if (t.should_live(max_age)) alive.push_back(std::move(t));
            else if (t.active) { t.trim_dead_history(); dead.push_back(std::move(t)); }
        }
        trackers.swap(alive);
        for (size_t j = 0; j < dets.size(); j++)
            if (!det_used[j]) trackers.emplace_back(id_counter++, dets[j], pts);
        return true;
    }

    void mark_seen(uint64_t ts) { for (auto &t : trackers) t.seen_ts.push_back(ts); }  // lib.rs:189-193
    void mark_active_seen(uint64_t ts) {                                              // lib.rs:195-201
        for (auto &t : trackers) if (t.active && t.start <= ts) t.seen_ts.push_back(ts);
    }
    bool any_valid() const { for (auto &t : trackers) if (t.active) return true; return false; }
    uint64_t oldest_start() const {  // cova/tracker.rs:85-90
        uint64_t m = UINT64_MAX;
        for (auto &t : trackers) m = std::min(m, t.start);
        return m;
    }
    // lib.rs:207-213: active trackers leave the set; only those with more than min_hits history entries are returned
    std::vector<KalmanBoxTracker> finalize() {
        std::vector<KalmanBoxTracker> out, rest;
        for (auto &t : trackers) {
            if (!t.active) rest.push_back(std::move(t));
            else if (t.history.size() > (size_t)min_hits) out.push_back(std::move(t));
        }
        trackers.swap(rest);
        return out;
    }
};

}  // namespace host
}  // namespace cova
