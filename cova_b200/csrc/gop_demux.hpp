// Host side, the step BEFORE the entropy decoder: find the frames and key frames
// of an H.264 stream and cut it into per-GPU shards of whole GoPs the way `gopsplit` does, so that the
// multi-GPU driver can consume a real .mp4 / .h264 instead of pre-cut arrays.  Plain C++17, no CUDA.
//
// Written from the behaviour of (paths relative to the reference tree)
//   gst-plugins/gst-gopsplit/gstgopsplit.cpp:700-729   chain: a buffer without DELTA_UNIT starts a GoP
//   gst-plugins/gst-gopsplit/gstgopsplit.cpp:500-640   split_and_push: floor(G/P) GoPs per pad, remainder to the
//                                                      last pad; G < P: pad i gets GoP i, the other pads nothing
//   pipeline/cova/pipeline.py:60-92                    filesrc ! qtdemux ! h264parse ! gopsplit ! avdec_h264
// Not in the tree (GStreamer): qtdemux marks every sample that is not in the `stss` sync-sample table as
// DELTA_UNIT; h264parse on a byte stream marks access units without an IDR slice.  Both are restated here from
// ISO/IEC 14496-12 (box layout) and ITU-T H.264 section 7.3 / Annex B (NAL unit syntax).
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

namespace cova {
namespace host {

struct Sample {
    uint64_t offset;   // byte offset in the file / stream
    uint32_t size;
    uint32_t flags;    // 1 = DELTA_UNIT (not a key frame)
    uint64_t dts, pts; // in track timescale units (MP4) or frame index (Annex B)
};

struct Mp4Info {
    uint32_t timescale = 0, width = 0, height = 0, nal_length_size = 0;
};

// ---- ISO base media file: sample table of the first video ('vide') track --------------------------------------------
struct BoxReader {
    const uint8_t *d;
    size_t n;
    static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
    static uint64_t be64(const uint8_t *p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }
    // finds the first child box of `type` in [beg, end); body = [*b, *e)
    bool find(size_t beg, size_t end, const char *type, size_t *b, size_t *e, size_t from = 0) const {
        size_t off = from ? from : beg;
        while (off + 8 <= end) {
            uint64_t sz = be32(d + off);
            size_t hdr = 8;
            if (sz == 1) { if (off + 16 > end) return false; sz = be64(d + off + 8); hdr = 16; }
            else if (sz == 0) sz = end - off;
            if (sz < hdr || off + sz > end) return false;
            if (!memcmp(d + off + 4, type, 4)) { *b = off + hdr; *e = off + (size_t)sz; return true; }
            off += (size_t)sz;
        }
        return false;
    }
};

// 0 ok; -1 malformed / truncated; -2 no video track with an avc1 sample entry
inline int mp4_video_samples(const uint8_t *data, size_t len, std::vector<Sample> &out, Mp4Info &info) {
    BoxReader r{data, len};
    size_t mb, me;
    if (!r.find(0, len, "moov", &mb, &me)) return -1;
    size_t from = 0;
    for (;;) {
        size_t tb, te;
        if (!r.find(mb, me, "trak", &tb, &te, from)) return -2;
        from = te;
        size_t db, de, hb, he, ib, ie, sb, se;
        if (!r.find(tb, te, "mdia", &db, &de)) continue;
        if (!r.find(db, de, "hdlr", &hb, &he) || he - hb < 12 || memcmp(data + hb + 8, "vide", 4)) continue;
        size_t mhb, mhe;
        if (!r.find(db, de, "mdhd", &mhb, &mhe) || mhe - mhb < 24) return -1;
        info.timescale = data[mhb] == 1 ? BoxReader::be32(data + mhb + 20) : BoxReader::be32(data + mhb + 12);
        if (!r.find(db, de, "minf", &ib, &ie) || !r.find(ib, ie, "stbl", &sb, &se)) return -1;
        size_t b, e;
        // stsd -> avc1 -> avcC (lengthSizeMinusOne)
        if (!r.find(sb, se, "stsd", &b, &e) || e - b < 16) return -1;
        {
            size_t eb = b + 8;   // first sample entry: size, type, 6 reserved, data_ref, 16 video bytes, width, height ...
            if (eb + 86 > e) return -1;
            if (memcmp(data + eb + 4, "avc1", 4) && memcmp(data + eb + 4, "avc3", 4)) return -2;
            info.width = (data[eb + 32] << 8) | data[eb + 33];
            info.height = (data[eb + 34] << 8) | data[eb + 35];
            // the sample entry's own size field bounds the search for avcC; a malformed (oversized or undersized) entry
            // must not send the reader past the stsd body
            const size_t entry_size = BoxReader::be32(data + eb);
            if (entry_size < 86) return -1;
            const size_t entry_end = entry_size > e - eb ? e : eb + entry_size;
            size_t cb, ce;
            if (r.find(eb + 86, entry_end, "avcC", &cb, &ce) && ce - cb >= 5) info.nal_length_size = (data[cb + 4] & 3) + 1;
        }
        std::vector<uint32_t> sizes;
        if (!r.find(sb, se, "stsz", &b, &e) || e - b < 12) return -1;
        {
            const uint32_t fixed = BoxReader::be32(data + b + 4), cnt = BoxReader::be32(data + b + 8);
            if (!fixed && (e - b - 12) / 4 < cnt) return -1;
            if (cnt > len) return -1;               // every sample occupies at least one byte of the input: bounds the allocation
            sizes.resize(cnt);
            for (uint32_t i = 0; i < cnt; i++) sizes[i] = fixed ? fixed : BoxReader::be32(data + b + 12 + 4 * (size_t)i);
        }
        std::vector<uint64_t> chunk_off;
        if (r.find(sb, se, "stco", &b, &e)) {
            if (e - b < 8) return -1;
            const uint32_t cnt = BoxReader::be32(data + b + 4);
            if ((e - b - 8) / 4 < cnt) return -1;
            for (uint32_t i = 0; i < cnt; i++) chunk_off.push_back(BoxReader::be32(data + b + 8 + 4 * (size_t)i));
        } else if (r.find(sb, se, "co64", &b, &e)) {
            if (e - b < 8) return -1;
            const uint32_t cnt = BoxReader::be32(data + b + 4);
            if ((e - b - 8) / 8 < cnt) return -1;
            for (uint32_t i = 0; i < cnt; i++) chunk_off.push_back(BoxReader::be64(data + b + 8 + 8 * (size_t)i));
        } else return -1;
        if (!r.find(sb, se, "stsc", &b, &e) || e - b < 8) return -1;
        const uint32_t n_stsc = BoxReader::be32(data + b + 4);
        if ((e - b - 8) / 12 < n_stsc) return -1;
        out.assign(sizes.size(), Sample{0, 0, 1, 0, 0});
        {
            size_t s = 0;
            for (uint32_t k = 0; k < n_stsc && s < sizes.size(); k++) {
                const uint8_t *ent = data + b + 8 + 12 * (size_t)k;
                const uint32_t first = BoxReader::be32(ent), per = BoxReader::be32(ent + 4);
                const uint32_t next = k + 1 < n_stsc ? BoxReader::be32(ent + 12) : (uint32_t)chunk_off.size() + 1;
                if (first < 1 || next < first) return -1;
                for (uint32_t c = first; c < next && c <= chunk_off.size() && s < sizes.size(); c++) {
                    uint64_t off = chunk_off[c - 1];
                    for (uint32_t j = 0; j < per && s < sizes.size(); j++, s++) {
                        out[s].offset = off, out[s].size = sizes[s];
                        off += sizes[s];
                    }
                }
            }
            if (s != sizes.size()) return -1;
        }
        if (r.find(sb, se, "stts", &b, &e) && e - b >= 8) {
            const uint32_t cnt = BoxReader::be32(data + b + 4);
            if ((e - b - 8) / 8 < cnt) return -1;
            uint64_t t = 0;
            size_t s = 0;
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t c = BoxReader::be32(data + b + 8 + 8 * (size_t)k), dlt = BoxReader::be32(data + b + 12 + 8 * (size_t)k);
                for (uint32_t j = 0; j < c && s < out.size(); j++, s++) out[s].dts = out[s].pts = t, t += dlt;
            }
        }
        if (r.find(sb, se, "ctts", &b, &e) && e - b >= 8) {
            const uint32_t cnt = BoxReader::be32(data + b + 4);
            if ((e - b - 8) / 8 < cnt) return -1;
            size_t s = 0;
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t c = BoxReader::be32(data + b + 8 + 8 * (size_t)k);
                const int32_t o = (int32_t)BoxReader::be32(data + b + 12 + 8 * (size_t)k);
                for (uint32_t j = 0; j < c && s < out.size(); j++, s++) out[s].pts = (uint64_t)((int64_t)out[s].dts + o);
            }
        }
        if (r.find(sb, se, "stss", &b, &e) && e - b >= 8) {   // sync samples are the key frames
            const uint32_t cnt = BoxReader::be32(data + b + 4);
            if ((e - b - 8) / 4 < cnt) return -1;
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t s = BoxReader::be32(data + b + 8 + 4 * (size_t)k);
                if (s >= 1 && s <= out.size()) out[s - 1].flags = 0;
            }
        } else {
            for (auto &s : out) s.flags = 0;   // no stss: every sample is a sync sample (14496-12, 8.6.2)
        }
        return 0;
    }
}

// ---- Annex B byte stream: access units and key frames -----------------------------------------------------------
// An access unit starts at an AUD (type 9), at an SPS/PPS/SEI (7, 8, 6) that follows a slice, or at a slice
// (types 1, 5) whose first_mb_in_slice is 0 when the unit already holds a slice.  Key frame = holds an IDR slice.
inline void annexb_frames(const uint8_t *d, size_t n, std::vector<Sample> &out) {
    out.clear();
    struct Nal { size_t sc, payload, end; };
    std::vector<Nal> nals;
    for (size_t i = 0; i + 3 <= n;) {
        if (d[i] == 0 && d[i + 1] == 0 && d[i + 2] == 1) {
            const size_t sc = (i > 0 && d[i - 1] == 0) ? i - 1 : i;
            if (!nals.empty()) nals.back().end = sc;
            nals.push_back(Nal{sc, i + 3, n});
            i += 3;
        } else {
            i++;
        }
    }
    bool have_slice = false, open = false, key = false;
    size_t start = 0;
    auto close = [&](size_t end) {
        if (open && have_slice) {
            out.push_back(Sample{start, (uint32_t)(end - start), key ? 0u : 1u, (uint64_t)out.size(), (uint64_t)out.size()});
        }
    };
    for (const Nal &nal : nals) {
        if (nal.payload >= nal.end) continue;
        const int type = d[nal.payload] & 0x1f;
        bool boundary = false;
        if (type == 9 || type == 6 || type == 7 || type == 8) boundary = have_slice;
        else if (type == 1 || type == 5) {
            // first_mb_in_slice = ue(v) right after the NAL header: zero <=> its first bit is 1
            const bool first_mb_zero = nal.payload + 1 < nal.end && (d[nal.payload + 1] & 0x80);
            boundary = have_slice && first_mb_zero;
        }
        if (boundary) { close(nal.sc); open = false; }
        if (!open) { open = true, start = nal.sc, have_slice = false, key = false; }
        if (type == 1 || type == 5) have_slice = true;
        if (type == 5) key = true;
    }
    close(n);
}

// ---- gopsplit: contiguous [first_frame, end_frame) per pad -------------------------------------------------------
// Delta frames that precede the first key frame form a GoP of their own (gstgopsplit.cpp:712-726: they are appended
// to `bufs`, which the first IDR then closes).  false: no pads.
inline bool gopsplit_ranges(const uint32_t *flags, size_t n_frames, uint32_t n_pads, uint64_t *first, uint64_t *end) {
    if (!n_pads) return false;
    std::vector<size_t> starts;
    for (size_t i = 0; i < n_frames; i++)
        if (i == 0 || !(flags[i] & 1u)) starts.push_back(i);
    const size_t G = starts.size();
    auto frame_at = [&](size_t g) { return g < G ? starts[g] : n_frames; };
    for (uint32_t p = 0; p < n_pads; p++) first[p] = end[p] = 0;
    if (G == 0) return true;
    if (G < n_pads) {   // "Too many pads": pad i pushes GoP i
        for (size_t g = 0; g < G; g++) first[g] = frame_at(g), end[g] = frame_at(g + 1);
        return true;
    }
    const size_t per = G / n_pads;
    for (uint32_t p = 0; p < n_pads; p++) {
        first[p] = frame_at((size_t)p * per);
        end[p] = p + 1 == n_pads ? n_frames : frame_at((size_t)(p + 1) * per);
    }
    return true;
}

}  // namespace host
}  // namespace cova
