"""CPU oracle of the `cova` element's frame selection.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain-Python restatement of (paths relative to the
reference tree)

* ``sink_enc_chain``    cova-rs/gst-plugins/src/cova/imp.rs:292-331
* ``sink_mask_chain``   cova-rs/gst-plugins/src/cova/imp.rs:90-289
* EOS handling          cova-rs/gst-plugins/src/cova/imp.rs:332-431
* ``Tracker``           cova-rs/gst-plugins/src/cova/tracker.rs:43-125
* ``Frame``             cova-rs/bbox/src/lib.rs:7-22 (bincode), framed by tokio_util's LengthDelimitedCodec
                        (default configuration: 4-byte big-endian length; crate not in the tree)

PARITY UNPINNED: the reference has no test for this element; this file follows the source line by line, with
its quirks (the BytesMut that is never cleared between dead tracks; the buffer popped and lost when the track was
already inferenced in a newer GoP).
"""
from __future__ import annotations

import struct

from . import sort_ref
from .bboxcc_ref import serialize_vec

DELTA_UNIT, DISCONT, DROPPABLE = 1, 2, 4
SECOND = 1_000_000_000
EMPTY_LIST = 2**64 - 1


class CovaSelectRef:
    def __init__(self, sort_iou=0.1, sort_maxage=30, sort_minhits=30, port=0, infer_i=False, alpha=0, beta=0):
        self.sort_iou, self.sort_maxage, self.sort_minhits = sort_iou, sort_maxage, sort_minhits
        self.port, self.infer_i, self.alpha, self.beta = port, infer_i, alpha, beta
        self.decoded_dependency = self.decoded_inference = self.dropped = 0
        self.bufs: list[list] = []          # [min, max, in, out, finalized]; buffers are [id, pts, flags]
        self.sort = None
        self.range_start = None
        self.eos = [False, False]
        self.wire = bytearray()

    # ---- Tracker (cova/tracker.rs)
    def _write_frames(self, tracks, oldest):
        if not self.port:
            return
        acc = bytearray()
        for t in tracks:
            body = struct.pack("<QQ", self.range_start, oldest) + serialize_vec([tuple(b) for b in t.history])
            acc += struct.pack(">I", len(body)) + body
            self.wire += acc

    def _tracker_update(self, boxes, pts):
        if self.sort is None:
            self.sort = sort_ref.Sort(self.sort_maxage, self.sort_minhits, self.sort_iou)
        if self.range_start is None:
            self.range_start = pts
        dead = self.sort.update(boxes, pts)
        ret = max([t.start for t in dead if not t.is_seen()], default=0) if dead else None
        self._write_frames(dead, self.sort.oldest_start())
        return ret

    # ---- pads
    def sink_enc(self, buf_id, pts, flags):
        if not flags & DELTA_UNIT:
            if self.bufs:
                self.bufs[-1][4] = True
            self.bufs.append([pts, pts, [[buf_id, pts, DISCONT]], [], False])
        else:
            back = self.bufs[-1]
            if pts < back[0]:
                back[0] = pts
            elif pts > back[1]:
                back[1] = pts
            back[2].append([buf_id, pts, DELTA_UNIT])

    def sink_mask(self, boxes, pts):
        pushed = []
        min_track = self._tracker_update(boxes, pts)
        maxage_pts = (SECOND // 30) * (self.sort_maxage + 10)
        max_track = pts - maxage_pts if pts >= maxage_pts else 0
        if min_track is not None:
            ti = dep = inf = 0
            sel = [g for g in reversed(self.bufs) if min_track <= g[1] and g[0] <= max_track]
            for g in sel:
                if any(min_track < b[1] for b in g[3]):
                    ti += 1
                    continue
                while g[2]:
                    b = g[2].pop(0)
                    if ti > 0:
                        break
                    if min_track <= b[1]:
                        self.sort.mark_seen(b[1])
                        inf += 1
                        g[3].append(b)
                        ti += 1
                        break
                    b[2] |= DROPPABLE
                    dep += 1
                    g[3].append(b)
            if ti < self.beta:
                for g in sel:
                    if not g[3]:
                        continue
                    extra_decode = min(len(g[2]), self.alpha)
                    extra_infer = min(extra_decode, self.beta - ti)
                    if extra_decode == 0 or extra_infer == 0:
                        continue
                    step, rem = divmod(extra_decode, extra_infer)

                    def pop_dep():
                        nonlocal dep
                        b = g[2].pop(0)
                        b[2] |= DROPPABLE
                        dep += 1
                        g[3].append(b)
                    for _ in range(rem):
                        pop_dep()
                    for _ in range(extra_infer):
                        for _ in range(max(step - 1, 0)):
                            pop_dep()
                        b = g[2].pop(0)
                        self.sort.mark_seen(b[1])
                        inf += 1
                        g[3].append(b)
                        ti += 1
            self.decoded_inference += inf
            self.decoded_dependency += dep
            assert ti > 0
        gop_pts = SECOND // 30 * 250
        droppable = pts - gop_pts if pts >= gop_pts else 0
        keep, n_lists = [], 0
        for g in self.bufs:
            if not (g[4] and g[1] <= droppable):
                keep.append(g)
                continue
            if self.infer_i and g[2]:
                b = g[2].pop(0)
                if not b[2] & DELTA_UNIT:
                    self.decoded_inference += 1
                    g[3].append(b)
                else:
                    self.dropped += 1
            if g[3]:
                pushed += [(b[0], b[1], b[2], n_lists) for b in g[3]]
                n_lists += 1
            self.dropped += len(g[2])
        self.bufs = keep
        return pushed

    def on_eos(self, pad):
        self.eos[pad] = True
        if not all(self.eos):
            return None
        pushed = []
        for n_lists, g in enumerate(self.bufs):
            self.dropped += len(g[2])
            pushed += [(b[0], b[1], b[2], n_lists) for b in g[3]] or [(EMPTY_LIST, 0, 0, n_lists)]
        self.bufs = []
        # imp.rs:387-390 vs 399-424: only the sink_mask event handler takes and flushes the tracker; a sink_enc EOS that
        # arrives second drains the GoP lists and leaves the active tracks unwritten
        if self.sort is not None and pad == 1:
            oldest = self.sort.oldest_start()
            self._write_frames(self.sort.finalize(), oldest)
            self.sort = None
        return pushed


def split_wire(wire: bytes):
    """LengthDelimitedCodec frames -> list of (range_start, oldest, bincode(Vec<Bbox>) bytes)."""
    out, off = [], 0
    while off < len(wire):
        (n,) = struct.unpack_from(">I", wire, off)
        body = wire[off + 4: off + 4 + n]
        out.append(struct.unpack_from("<QQ", body, 0) + (bytes(body[16:]),))
        off += 4 + n
    assert off == len(wire)
    return out
