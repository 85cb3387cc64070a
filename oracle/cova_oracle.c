/* CPU oracle, plain C restatement of the byte/integer stages of the CoVA blob-detection path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs.  The product library (cova_b200/csrc) never links or calls this file.
 *
 * Follows (reference tree):
 *   oracle_metapreprocess_*   cova-rs/gst-plugins/src/metapreprocess/imp.rs:204-236,288-332
 *   oracle_ccl                cova-rs/gst-plugins/src/bboxcc/process.rs:14-30 -> third-party
 *                             cv::connectedComponentsWithStats(8-connectivity, CV_32S); label order =
 *                             raster order of the first 2x2-aligned block of each component
 *   oracle_regionprops        cova-rs/gst-plugins/src/bboxcc/process.rs:37-48 + cova-rs/bbox/src/bbox.rs:17-29,84-86
 *
 * Pinned against cv2 golden vectors through tests/test_oracle.py (tests/golden/ccl_golden.npz).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ metapreprocess */
typedef struct {
    uint32_t timestep, gamma, gamma_idx, n_prev;
    size_t size_per_buf;
    uint8_t *prev; /* (timestep-1) slots, slot 0 = newest (LinkedList push_front / pop_back) */
} oracle_mp;

oracle_mp *oracle_metapreprocess_new(uint32_t width_px, uint32_t height_px, uint32_t timestep, uint32_t gamma) {
    if (!timestep || !gamma) return NULL;
    oracle_mp *m = (oracle_mp *)calloc(1, sizeof(*m));
    m->timestep = timestep;
    m->gamma = gamma;
    /* imp.rs:262-268: width/16, height/16*timestep, RGBA; :233 size_per_buf = size / timestep */
    size_t out_size = (size_t)(width_px / 16) * (height_px / 16 * timestep) * 4;
    m->size_per_buf = out_size / timestep;
    m->prev = (uint8_t *)malloc(m->size_per_buf * (timestep > 1 ? timestep - 1 : 1));
    return m;
}

void oracle_metapreprocess_free(oracle_mp *m) {
    if (m) { free(m->prev); free(m); }
}

static void mp_push_front(oracle_mp *m, const uint8_t *in, int pop_back) {
    size_t S = m->size_per_buf;
    uint32_t cap = m->timestep - 1;
    if (!cap) return;
    uint32_t keep = pop_back ? cap - 1 : m->n_prev;
    memmove(m->prev + S, m->prev, (size_t)keep * S);
    memcpy(m->prev, in, S);
    if (!pop_back) m->n_prev++;
}

/* returns 0 = Ok (out written: timestep*size_per_buf bytes), 1 = DROPPED */
int oracle_metapreprocess_transform(oracle_mp *m, const uint8_t *in, uint8_t *out) {
    size_t S = m->size_per_buf;
    if (m->n_prev < m->timestep - 1) {           /* imp.rs:302-305 */
        mp_push_front(m, in, 0);
        return 1;
    }
    if (m->gamma_idx == 0) {                     /* imp.rs:306-324 */
        memcpy(out, in, S);
        for (uint32_t k = 0; k < m->n_prev; k++) memcpy(out + (size_t)(k + 1) * S, m->prev + (size_t)k * S, S);
        mp_push_front(m, in, 1);
        m->gamma_idx = m->gamma - 1;
        return 0;
    }
    mp_push_front(m, in, 1);                     /* imp.rs:325-330 */
    m->gamma_idx -= 1;
    return 1;
}

/* ------------------------------------------------------------------ CCL */
static int32_t uf_find(int32_t *p, int32_t i) {
    int32_t r = i;
    while (p[r] != r) r = p[r];
    while (p[i] != r) { int32_t n = p[i]; p[i] = r; i = n; }
    return r;
}
static void uf_union(int32_t *p, int32_t a, int32_t b) {
    a = uf_find(p, a); b = uf_find(p, b);
    if (a < b) p[b] = a; else if (b < a) p[a] = b;
}

/* labels: i32 [h*w] out; stats: i32 [(max_labels)*5] out (row 0 = background, left zero);
 * scratch: caller-provided i32 [3*h*w + 16].  returns number of labels including background. */
int oracle_ccl(const uint8_t *mask, int h, int w, int32_t *labels, int32_t *stats, int32_t *scratch) {
    int n = h * w;
    int32_t *parent = scratch, *key = scratch + n, *order = scratch + 2 * n;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int p = y * w + x;
            if (!mask[p]) { parent[p] = -1; continue; }
            parent[p] = p;
            if (x > 0 && mask[p - 1]) uf_union(parent, p, p - 1);
            if (y > 0) {
                if (x > 0 && mask[p - w - 1]) uf_union(parent, p, p - w - 1);
                if (mask[p - w]) uf_union(parent, p, p - w);
                if (x + 1 < w && mask[p - w + 1]) uf_union(parent, p, p - w + 1);
            }
        }
    int wb = (w + 1) / 2, nb = wb * ((h + 1) / 2);
    /* key[root] = min block index over the component; order[block] = root owning that first block */
    for (int p = 0; p < n; p++) key[p] = 0x7fffffff;
    for (int b = 0; b < nb; b++) order[b] = -1;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int p = y * w + x;
            if (parent[p] < 0) continue;
            int r = uf_find(parent, p);
            int k = (y >> 1) * wb + (x >> 1);
            if (k < key[r]) key[r] = k;
        }
    for (int p = 0; p < n; p++)
        if (parent[p] == p) order[key[p]] = p;   /* keys are unique per component */
    int nl = 1;
    for (int b = 0; b < nb; b++)
        if (order[b] >= 0) { key[order[b]] = nl; nl++; }   /* reuse key[] as root -> label */
    memset(stats, 0, sizeof(int32_t) * 5 * (size_t)nl);
    for (int k = 1; k < nl; k++) { stats[5 * k] = w; stats[5 * k + 1] = h; stats[5 * k + 2] = -1; stats[5 * k + 3] = -1; }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int p = y * w + x;
            if (parent[p] < 0) { labels[p] = 0; continue; }
            int l = key[uf_find(parent, p)];
            labels[p] = l;
            int32_t *s = stats + 5 * l;
            if (x < s[0]) s[0] = x;
            if (y < s[1]) s[1] = y;
            if (x > s[2]) s[2] = x;
            if (y > s[3]) s[3] = y;
            s[4]++;
        }
    for (int k = 1; k < nl; k++) { stats[5 * k + 2] = stats[5 * k + 2] - stats[5 * k] + 1; stats[5 * k + 3] = stats[5 * k + 3] - stats[5 * k + 1] + 1; }
    return nl;
}

/* mask -> bincode(Vec<Bbox>) ; returns bytes written, or -(required) if cap too small.
 * scratch: i32 [3*h*w+16 + h*w + 5*(ceil(h/2)*ceil(w/2)+1)] */
long oracle_bboxcc(const uint8_t *mask, int h, int w, int32_t area_thresh, uint8_t *out, size_t cap, int32_t *scratch) {
    int n = h * w;
    int32_t *labels = scratch + 3 * n + 16;
    int32_t *stats = labels + n;
    int nl = oracle_ccl(mask, h, w, labels, stats, scratch);
    uint64_t cnt = 0;
    for (int k = 1; k < nl; k++) if (stats[5 * k + 4] >= area_thresh) cnt++;
    size_t need = 8 + 24 * (size_t)cnt;
    if (need > cap) return -(long)need;
    memcpy(out, &cnt, 8);
    uint8_t *o = out + 8;
    for (int k = 1; k < nl; k++) {
        if (stats[5 * k + 4] < area_thresh) continue;
        float f[5] = {(float)stats[5 * k], (float)stats[5 * k + 1], (float)stats[5 * k + 2], (float)stats[5 * k + 3], 0.f};
        f[4] = f[2] * f[3];                      /* bbox.rs:23 area = width*height */
        memcpy(o, f, 20);
        memset(o + 20, 0, 4);                    /* four Option::None tags */
        o += 24;
    }
    return (long)need;
}

/* batch helpers so the CPU baseline is timed without per-call FFI overhead */
long oracle_bboxcc_batch(const uint8_t *masks, int n_masks, int h, int w, int32_t area_thresh, uint8_t *out, size_t cap_each, int64_t *lens, int32_t *scratch) {
    long total = 0;
    for (int i = 0; i < n_masks; i++) {
        long r = oracle_bboxcc(masks + (size_t)i * h * w, h, w, area_thresh, out + (size_t)i * cap_each, cap_each, scratch);
        lens[i] = r;
        if (r > 0) total += r;
    }
    return total;
}

long oracle_metapreprocess_stream(const uint8_t *frames, int n_frames, uint32_t width_px, uint32_t height_px, uint32_t timestep, uint32_t gamma, uint8_t *out) {
    oracle_mp *m = oracle_metapreprocess_new(width_px, height_px, timestep, gamma);
    size_t S = m->size_per_buf;
    long n_out = 0;
    for (int f = 0; f < n_frames; f++)
        if (oracle_metapreprocess_transform(m, frames + (size_t)f * S, out + (size_t)n_out * S * timestep) == 0) n_out++;
    oracle_metapreprocess_free(m);
    return n_out;
}
