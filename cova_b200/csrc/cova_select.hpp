// Host side, two steps after the blob-detection path: the frame-selection logic of
// the `cova` element.  It keeps the ENCODED frames of every GoP, feeds the per-frame boxes to a SORT tracker and
// decides which frames a pixel decoder still has to decode: the first frame at or after the start of every track
// that died unseen, plus the frames it depends on (flagged DROPPABLE), optionally `alpha`/`beta` extra frames.
//
// Written from the behaviour of (paths relative to the reference tree)
//   cova-rs/gst-plugins/src/cova/imp.rs:90-289     sink_mask_chain (selection, droppable GoPs, counters)
//   cova-rs/gst-plugins/src/cova/imp.rs:292-331    sink_enc_chain (GoP bookkeeping)
//   cova-rs/gst-plugins/src/cova/imp.rs:332-431    EOS on both pads: drain + Tracker::flush
//   cova-rs/gst-plugins/src/cova/tracker.rs:43-125 Tracker::update / seen / flush, Frame wire format
//   cova-rs/bbox/src/lib.rs:7-22                   Frame { range_start, oldest, bboxes } (bincode)
// Third party not in the tree: tokio_util LengthDelimitedCodec (default: 4-byte big-endian length prefix).
// Reproduced on purpose: the BytesMut that accumulates encoded Frames is not cleared between dead tracks
// (tracker.rs:62-81), so track k of one update is preceded on the wire by tracks 0..k-1 again; a buffer popped from a
// GoP after the track was already inferenced elsewhere is lost without being counted (imp.rs:172-177).
#pragma once
#include <deque>
#include <list>
#include <memory>

#include "sort_tracker.hpp"

namespace cova {
namespace host {

constexpr uint32_t BUF_DELTA_UNIT = 1u, BUF_DISCONT = 2u, BUF_DROPPABLE = 4u;   // the gst::BufferFlags this logic touches

struct EncBuf {
    uint64_t id, pts;
    uint32_t flags;
};
struct Pushed {            // one buffer of one pushed gst::BufferList
    uint64_t id, pts;
    uint32_t flags, list;  // list: running index of the BufferList within this call (an EMPTY list pushed at EOS
};                         // is reported as one entry with id = UINT64_MAX)

struct Gop {
    uint64_t min, max;
    std::deque<EncBuf> in, out;
    bool finalized;
};

struct CovaSelect {
    // settings (imp.rs:22-56)
    float sort_iou = 0.1f;
    uint32_t sort_maxage = 30, sort_minhits = 30, port = 0, alpha = 0, beta = 0;
    bool infer_i = false, debug = false;
    // counters (imp.rs:72-77)
    uint64_t decoded_dependency = 0, decoded_inference = 0, dropped = 0;
    // state
    std::list<Gop> bufs;
    std::unique_ptr<Sort> sort;
    bool have_range = false;
    uint64_t range_start = 0;
    bool eos[2] = {false, false};
    std::vector<uint8_t> wire;   // what Tracker would have written to its TcpStream (only when port != 0)

    static constexpr uint64_t kSecond = 1000000000ull;

    void write_frames(const std::vector<KalmanBoxTracker> &tracks, uint64_t oldest) {
        if (!port) return;
        std::vector<uint8_t> acc;   // the BytesMut of tracker.rs:62 / :100 - never cleared inside one call
        for (const auto &t : tracks) {
            const size_t body = 16 + boxes_wire_size(t.history);
            const size_t at = acc.size();
            acc.resize(at + 4 + body);
            uint8_t *p = acc.data() + at;
            p[0] = (uint8_t)(body >> 24), p[1] = (uint8_t)(body >> 16), p[2] = (uint8_t)(body >> 8), p[3] = (uint8_t)body;
            p += 4;
            put<uint64_t>(p, range_start), put<uint64_t>(p, oldest);
            encode_boxes_into(t.history, p);
            wire.insert(wire.end(), acc.begin(), acc.end());
        }
    }

    // tracker.rs:43-83.  has = false <=> None
    bool tracker_update(std::vector<Bbox> boxes, uint64_t pts, bool &has, uint64_t &min_required) {
        if (!sort) sort.reset(new Sort(sort_maxage, sort_minhits, (float)(double)sort_iou));
        if (!have_range) have_range = true, range_start = pts;
        std::vector<KalmanBoxTracker> dead;
        if (!sort->update(std::move(boxes), pts, dead)) return false;
        has = !dead.empty();
        min_required = 0;
        for (const auto &t : dead)
            if (!t.is_seen()) min_required = std::max(min_required, t.start);
        write_frames(dead, sort->oldest_start());
        return true;
    }

    // imp.rs:292-331.  false: a delta unit arrived before any key frame (the reference unwraps None)
    bool push_enc(uint64_t id, uint64_t pts, bool delta_unit) {
        if (!delta_unit) {
            if (!bufs.empty()) bufs.back().finalized = true;
            Gop g{pts, pts, {}, {}, false};
            g.in.push_back(EncBuf{id, pts, BUF_DISCONT});
            bufs.push_back(std::move(g));
        } else {
            if (bufs.empty()) return false;
            Gop &b = bufs.back();
            if (pts < b.min) b.min = pts;
            else if (pts > b.max) b.max = pts;
            b.in.push_back(EncBuf{id, pts, BUF_DELTA_UNIT});
        }
        return true;
    }

    static void push_list(std::deque<EncBuf> &out, std::vector<Pushed> &pushed, uint32_t &n_lists) {
        if (out.empty()) pushed.push_back(Pushed{UINT64_MAX, 0, 0, n_lists});
        for (const EncBuf &b : out) pushed.push_back(Pushed{b.id, b.pts, b.flags, n_lists});
        out.clear();
        n_lists++;
    }

    // imp.rs:90-289.  rc: 0 ok, -1 Kalman failure, -2 the reference's assert!(track_inferenced > 0) would fire
    int push_boxes(std::vector<Bbox> boxes, uint64_t pts, std::vector<Pushed> &pushed) {
        bool has = false;
        uint64_t min_track_pts = 0;
        if (!tracker_update(std::move(boxes), pts, has, min_track_pts)) return -1;
        const uint64_t maxage_pts = (kSecond / 30) * ((uint64_t)sort_maxage + 10);   // SAFETY_BUFFER = 10 frames
        const uint64_t max_track_pts = pts >= maxage_pts ? pts - maxage_pts : 0;
        if (has) {
            size_t track_inferenced = 0;
            uint64_t dep = 0, inf = 0;
            auto in_range = [&](const Gop &g) { return min_track_pts <= g.max && g.min <= max_track_pts; };
            for (auto it = bufs.rbegin(); it != bufs.rend(); ++it) {   // newest GoP first
                Gop &g = *it;
                if (!in_range(g)) continue;
                bool already = false;
                for (const EncBuf &b : g.out)
                    if (min_track_pts < b.pts) { track_inferenced++; already = true; break; }
                if (already) continue;
                while (!g.in.empty()) {
                    EncBuf b = g.in.front();
                    g.in.pop_front();
                    if (track_inferenced > 0) break;   // the popped buffer is gone (imp.rs:172-177)
                    if (min_track_pts <= b.pts) {
                        sort->mark_seen(b.pts);
                        inf++, g.out.push_back(b), track_inferenced++;
                        break;
                    }
                    b.flags |= BUF_DROPPABLE;
                    dep++, g.out.push_back(b);
                }
            }
            if (track_inferenced < (size_t)beta) {
                for (auto it = bufs.rbegin(); it != bufs.rend(); ++it) {
                    Gop &g = *it;
                    if (!in_range(g) || g.out.empty()) continue;
                    const size_t extra_decode = std::min(g.in.size(), (size_t)alpha);
                    const size_t extra_infer = std::min(extra_decode, (size_t)beta - track_inferenced);
                    if (!extra_decode || !extra_infer) continue;
                    const size_t step = extra_decode / extra_infer, rem = extra_decode % extra_infer;
                    auto pop_dep = [&]() {
                        EncBuf b = g.in.front();
                        g.in.pop_front();
                        b.flags |= BUF_DROPPABLE;
                        dep++, g.out.push_back(b);
                    };
                    for (size_t i = 0; i < rem; i++) pop_dep();
                    for (size_t i = 0; i < extra_infer; i++) {
                        for (size_t k = 0; k + 1 < step; k++) pop_dep();
                        EncBuf b = g.in.front();
                        g.in.pop_front();
                        sort->mark_seen(b.pts);
                        inf++, g.out.push_back(b), track_inferenced++;
                    }
                }
            }
            decoded_inference += inf, decoded_dependency += dep;
            if (!track_inferenced) return -2;
        }
        uint64_t n_dropped = 0, inf = 0;
        const uint64_t gop_pts = kSecond / 30 * 250;
        const uint64_t droppable_pts = pts >= gop_pts ? pts - gop_pts : 0;
        uint32_t n_lists = 0;
        for (auto it = bufs.begin(); it != bufs.end();) {
            Gop &g = *it;
            if (!(g.finalized && g.max <= droppable_pts)) { ++it; continue; }
            if (infer_i && !g.in.empty()) {
                EncBuf b = g.in.front();
                g.in.pop_front();
                if (!(b.flags & BUF_DELTA_UNIT)) inf++, g.out.push_back(b);
                else n_dropped++;
            }
            if (!g.out.empty()) push_list(g.out, pushed, n_lists);
            n_dropped += g.in.size();
            it = bufs.erase(it);
        }
        decoded_inference += inf, dropped += n_dropped;
        return 0;
    }

    // imp.rs:332-431: which = 0 (sink_enc) or 1 (sink_mask); acts once both have seen EOS.  Returns true when drained.
    bool on_eos(int which, std::vector<Pushed> &pushed) {
        eos[which ? 1 : 0] = true;
        if (!(eos[0] && eos[1])) return false;
        uint32_t n_lists = 0;
        uint64_t n_dropped = 0;
        for (Gop &g : bufs) {
            n_dropped += g.in.size();
            push_list(g.out, pushed, n_lists);   // the reference pushes the list even when it is empty
        }
        bufs.clear();
        dropped += n_dropped;
        // Only sink_mask_event takes and flushes the tracker (imp.rs:387-390); when the sink_enc EOS completes the pair
        // (imp.rs:399-424) the GoP lists are drained but the still-active tracks are never written to the aggregator.
        if (sort && which == 1) {   // Tracker::flush, tracker.rs:96-125
            const uint64_t oldest = sort->oldest_start();
            write_frames(sort->finalize(), oldest);
            sort.reset();
        }
        return true;
    }
};

}  // namespace host
}  // namespace cova
