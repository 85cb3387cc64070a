// C ABI of the host-side components that sit directly behind the GPU path (include/cova_b200.h, "sorttracker").
// No CUDA in this translation unit.
#include <stdio.h>

#include <new>

#include "../../include/cova_b200.h"
#include "sort_tracker.hpp"

namespace cova {
extern thread_local char g_err[512];
}
static int fail(int code, const char *msg) {
    snprintf(cova::g_err, sizeof(cova::g_err), "%s", msg);
    return code;
}

using cova::host::Bbox;
using cova::host::KalmanBoxTracker;
using cova::host::Sort;

struct cova_sorttracker {
    uint32_t maxage, minhits;
    float iou;
    Sort *sort;  // created by set_caps (sorttracker/imp.rs:214-236), like the element
    std::vector<Bbox> scratch;
};

static int emit(const std::vector<KalmanBoxTracker> &tracks, uint8_t *out, size_t cap, size_t *out_len, std::vector<Bbox> &flat) {
    flat.clear();
    for (const auto &t : tracks) flat.insert(flat.end(), t.history.begin(), t.history.end());
    size_t need = cova::host::boxes_wire_size(flat);
    if (out_len) *out_len = need;
    if (need > cap || !out) return fail(COVA_E_TOOSMALL, "output buffer too small for the serialized tracks");
    cova::host::encode_boxes_into(flat, out);
    return COVA_OK;
}

extern "C" int cova_sorttracker_new(cova_sorttracker **out) {
    if (!out) return fail(COVA_E_INVAL, "null out");
    cova_sorttracker *s = new (std::nothrow) cova_sorttracker{30, 30, 0.1f, nullptr, {}};
    if (!s) return fail(COVA_E_NOMEM, "host allocation failed");
    *out = s;
    return COVA_OK;
}
extern "C" void cova_sorttracker_free(cova_sorttracker *s) {
    if (!s) return;
    delete s->sort;
    delete s;
}
extern "C" int cova_sorttracker_set_property(cova_sorttracker *s, const char *name, double value) {
    if (!s || !name) return fail(COVA_E_INVAL, "null argument");
    if (!strcmp(name, "iou-threshold")) {
        if (!(value >= 0.0 && value <= 1.0)) return fail(COVA_E_INVAL, "iou-threshold is a float in [0, 1]");
        s->iou = (float)value;
    } else if (!strcmp(name, "maxage")) {
        if (!(value >= 0.0 && value <= 4294967295.0)) return fail(COVA_E_INVAL, "maxage is a u32");
        s->maxage = (uint32_t)value;
    } else if (!strcmp(name, "minhits")) {
        if (!(value >= 0.0 && value <= 4294967295.0)) return fail(COVA_E_INVAL, "minhits is a u32");
        s->minhits = (uint32_t)value;
    } else {
        return fail(COVA_E_INVAL, "unknown property (iou-threshold, maxage, minhits)");
    }
    return COVA_OK;
}
extern "C" int cova_sorttracker_get_property(const cova_sorttracker *s, const char *name, double *value) {
    if (!s || !name || !value) return fail(COVA_E_INVAL, "null argument");
    if (!strcmp(name, "iou-threshold")) *value = s->iou;
    else if (!strcmp(name, "maxage")) *value = s->maxage;
    else if (!strcmp(name, "minhits")) *value = s->minhits;
    else return fail(COVA_E_INVAL, "unknown property (iou-threshold, maxage, minhits)");
    return COVA_OK;
}
extern "C" int cova_sorttracker_set_caps(cova_sorttracker *s, int32_t width, int32_t height) {
    if (!s) return fail(COVA_E_INVAL, "null handle");
    if (width < 0 || height < 0) return fail(COVA_E_INVAL, "caps width/height are in [0, i32::MAX]");
    delete s->sort;
    s->sort = new (std::nothrow) Sort(s->maxage, s->minhits, s->iou);
    return s->sort ? COVA_OK : fail(COVA_E_NOMEM, "host allocation failed");
}
extern "C" int cova_sorttracker_transform(cova_sorttracker *s, const uint8_t *boxes, size_t boxes_len, uint64_t pts_ns,
                                          uint8_t *out, size_t out_cap, size_t *out_len) {
    if (!s || !boxes) return fail(COVA_E_INVAL, "null argument");
    if (!s->sort) return fail(COVA_E_INVAL, "transform before set_caps");
    std::vector<Bbox> dets;
    if (!cova::host::decode_boxes(boxes, boxes_len, dets)) return fail(COVA_E_INVAL, "input is not bincode(Vec<Bbox>)");
    std::vector<KalmanBoxTracker> dead;
    if (!s->sort->update(std::move(dets), pts_ns, dead)) return fail(COVA_E_NUMERIC, "Kalman update: innovation covariance not positive definite");
    return emit(dead, out, out_cap, out_len, s->scratch);
}
extern "C" int cova_sorttracker_eos(cova_sorttracker *s, uint8_t *out, size_t out_cap, size_t *out_len) {
    if (!s) return fail(COVA_E_INVAL, "null handle");
    if (!s->sort) return fail(COVA_E_INVAL, "eos before set_caps");
    // the tracks leave the tracker only when the caller's buffer can take them
    std::vector<KalmanBoxTracker> keep = s->sort->trackers;
    std::vector<KalmanBoxTracker> fin = s->sort->finalize();
    int rc = emit(fin, out, out_cap, out_len, s->scratch);
    if (rc != COVA_OK) s->sort->trackers.swap(keep);
    return rc;
}
extern "C" int cova_sorttracker_n_tracks(const cova_sorttracker *s, uint32_t *n_total, uint32_t *n_active) {
    if (!s || !s->sort) return fail(COVA_E_INVAL, "no tracker state");
    uint32_t a = 0;
    for (const auto &t : s->sort->trackers) a += t.active;
    if (n_total) *n_total = (uint32_t)s->sort->trackers.size();
    if (n_active) *n_active = a;
    return COVA_OK;
}

// building blocks, exported so that the reference's own unit tests can be replayed through the ABI
extern "C" int cova_sort_linear_assignment(const float *cost, uint32_t n_trk, uint32_t n_det, int32_t *pairs, uint32_t *n_pairs) {
    if (!cost || !pairs || !n_pairs) return fail(COVA_E_INVAL, "null argument");
    std::vector<float> c(cost, cost + (size_t)n_trk * n_det);
    auto m = cova::host::linear_assignment(c, (int)n_trk, (int)n_det);
    for (size_t i = 0; i < m.size(); i++) pairs[2 * i] = m[i].first, pairs[2 * i + 1] = m[i].second;
    *n_pairs = (uint32_t)m.size();
    return COVA_OK;
}
extern "C" int cova_sort_iou_matrix(const float *preds, uint32_t n_preds, const float *dets, uint32_t n_dets, float *out) {
    if ((!preds && n_preds) || (!dets && n_dets) || !out) return fail(COVA_E_INVAL, "null argument");
    std::vector<Bbox> p, d;
    for (uint32_t i = 0; i < n_preds; i++) p.push_back(Bbox::make(preds[4 * i], preds[4 * i + 1], preds[4 * i + 2], preds[4 * i + 3]));
    for (uint32_t i = 0; i < n_dets; i++) d.push_back(Bbox::make(dets[4 * i], dets[4 * i + 1], dets[4 * i + 2], dets[4 * i + 3]));
    auto c = Sort::iou_cost(p, d);
    memcpy(out, c.data(), c.size() * sizeof(float));
    return COVA_OK;
}
extern "C" int cova_sort_match_dets(const float *preds, const uint8_t *active, uint32_t n_preds, const float *dets, uint32_t n_dets,
                                    float iou_threshold, int32_t *pairs, uint32_t *n_pairs) {
    if ((!preds && n_preds) || (!active && n_preds) || (!dets && n_dets) || !pairs || !n_pairs) return fail(COVA_E_INVAL, "null argument");
    std::vector<Bbox> p, d;
    std::vector<char> a(active, active + n_preds);
    for (uint32_t i = 0; i < n_preds; i++) p.push_back(Bbox::make(preds[4 * i], preds[4 * i + 1], preds[4 * i + 2], preds[4 * i + 3]));
    for (uint32_t i = 0; i < n_dets; i++) d.push_back(Bbox::make(dets[4 * i], dets[4 * i + 1], dets[4 * i + 2], dets[4 * i + 3]));
    auto m = Sort::match_dets(p, a, d, iou_threshold);
    for (size_t i = 0; i < m.size(); i++) pairs[2 * i] = m[i].first, pairs[2 * i + 1] = m[i].second;
    *n_pairs = (uint32_t)m.size();
    return COVA_OK;
}
