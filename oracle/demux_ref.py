"""CPU oracle for the demux + gopsplit step in front of the path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Independent Python restatements of

* ``gopsplit_ranges``   gst-plugins/gst-gopsplit/gstgopsplit.cpp:500-640,700-729 (which GoPs each src pad pushes)
* ``mp4_video_samples`` what qtdemux hands downstream for the first H.264 track: one buffer per `stsz` entry at the
                        offset given by `stco`/`stsc`, DELTA_UNIT unless listed in `stss` (ISO/IEC 14496-12; qtdemux
                        itself is GStreamer code outside the tree)
* ``annexb_frames``     what h264parse does on a byte stream: access-unit boundaries and IDR detection
                        (ITU-T H.264 7.3.1, 7.4.1.2.3, Annex B)

Pinned by the real file: tools/make_moov_fixture.py stores the moov box of demo/1m.mp4 (tests/golden/demo_1m_moov.bin);
its sync samples must be exactly the frames the patched decoder reports as all-intra in
tests/golden/demo_1m_meta.npz (frames 0, 250, ..., 1750), 1802 samples in total.
"""
from __future__ import annotations

import struct


def gopsplit_ranges(is_key, n_pads):
    """Simulates the element: chain() collects GoPs, split_and_push assigns them to pads; returns the
    [first_frame, end_frame) each pad receives."""
    if n_pads < 1:
        raise ValueError("there are no pads")
    gops, bufs = [], []
    for i, k in enumerate(is_key):          # chain(), :712-726
        if k:
            if bufs:
                gops.append(bufs)
            bufs = []
        bufs.append(i)
    if bufs:                                 # EOS, :664-672
        gops.append(bufs)
    per_pad = [[] for _ in range(n_pads)]
    n = len(gops)
    if n == 0:
        return [(0, 0)] * n_pads
    if n < n_pads:                           # :531-553
        for g in range(n):
            per_pad[g] += gops[g]
    else:
        per = n // n_pads
        for p in range(n_pads):              # :574-600
            for g in range(p * per, (p + 1) * per):
                per_pad[p] += gops[g]
        for g in range(per * n_pads, n):     # :603-625 remaining gops go to the last pad
            per_pad[-1] += gops[g]
    out = []
    for frames in per_pad:
        if not frames:
            out.append((0, 0))
        else:
            assert frames == list(range(frames[0], frames[-1] + 1)), "a pad receives a contiguous run"
            out.append((frames[0], frames[-1] + 1))
    return out


def _boxes(d, beg, end):
    off = beg
    while off + 8 <= end:
        size, typ = struct.unpack_from(">I4s", d, off)
        hdr = 8
        if size == 1:
            (size,) = struct.unpack_from(">Q", d, off + 8)
            hdr = 16
        elif size == 0:
            size = end - off
        yield typ, off + hdr, off + size
        off += size


def _child(d, beg, end, typ):
    for t, b, e in _boxes(d, beg, end):
        if t == typ:
            return b, e
    return None


def mp4_video_samples(d: bytes):
    """-> (list of (offset, size, is_key, dts, pts), info dict) for the first 'vide' track."""
    moov = _child(d, 0, len(d), b"moov")
    for t, tb, te in _boxes(d, *moov):
        if t != b"trak":
            continue
        mdia = _child(d, tb, te, b"mdia")
        hdlr = _child(d, *mdia, b"hdlr")
        if d[hdlr[0] + 8: hdlr[0] + 12] != b"vide":
            continue
        mdhd = _child(d, *mdia, b"mdhd")
        timescale = struct.unpack_from(">I", d, mdhd[0] + (20 if d[mdhd[0]] == 1 else 12))[0]
        stbl = _child(d, *_child(d, *mdia, b"minf"), b"stbl")
        stsd = _child(d, *stbl, b"stsd")
        entry = stsd[0] + 8
        width, height = struct.unpack_from(">HH", d, entry + 32)
        avcc = _child(d, entry + 86, entry + struct.unpack_from(">I", d, entry)[0], b"avcC")
        nls = (d[avcc[0] + 4] & 3) + 1
        b, _ = _child(d, *stbl, b"stsz")
        fixed, cnt = struct.unpack_from(">II", d, b + 4)
        sizes = [fixed] * cnt if fixed else list(struct.unpack_from(f">{cnt}I", d, b + 12))
        co = _child(d, *stbl, b"stco")
        if co:
            (n,) = struct.unpack_from(">I", d, co[0] + 4)
            chunks = list(struct.unpack_from(f">{n}I", d, co[0] + 8))
        else:
            co = _child(d, *stbl, b"co64")
            (n,) = struct.unpack_from(">I", d, co[0] + 4)
            chunks = list(struct.unpack_from(f">{n}Q", d, co[0] + 8))
        b, _ = _child(d, *stbl, b"stsc")
        (n,) = struct.unpack_from(">I", d, b + 4)
        stsc = [struct.unpack_from(">III", d, b + 8 + 12 * i) for i in range(n)]
        per_chunk = {}
        for i, (first, per, _desc) in enumerate(stsc):
            last = stsc[i + 1][0] - 1 if i + 1 < n else len(chunks)
            for c in range(first, last + 1):
                per_chunk[c] = per
        offsets, s = [], 0
        for c, off in enumerate(chunks, start=1):
            for _ in range(per_chunk[c]):
                if s >= len(sizes):
                    break
                offsets.append(off)
                off += sizes[s]
                s += 1
        b, _ = _child(d, *stbl, b"stts")
        (n,) = struct.unpack_from(">I", d, b + 4)
        dts, t = [], 0
        for i in range(n):
            c, delta = struct.unpack_from(">II", d, b + 8 + 8 * i)
            for _ in range(c):
                dts.append(t)
                t += delta
        pts = list(dts)
        ctts = _child(d, *stbl, b"ctts")
        if ctts:
            (n,) = struct.unpack_from(">I", d, ctts[0] + 4)
            s = 0
            for i in range(n):
                c, o = struct.unpack_from(">Ii", d, ctts[0] + 8 + 8 * i)
                for _ in range(c):
                    pts[s] = dts[s] + o
                    s += 1
        stss = _child(d, *stbl, b"stss")
        if stss:
            (n,) = struct.unpack_from(">I", d, stss[0] + 4)
            keys = {k - 1 for k in struct.unpack_from(f">{n}I", d, stss[0] + 8)}
        else:
            keys = set(range(len(sizes)))
        samples = [(offsets[i], sizes[i], i in keys, dts[i], pts[i]) for i in range(len(sizes))]
        return samples, dict(timescale=timescale, width=width, height=height, nal_length_size=nls)
    raise ValueError("no video track")


def annexb_frames(d: bytes):
    """-> list of (offset, size, is_key)."""
    nals, i = [], 0
    while True:
        j = d.find(b"\x00\x00\x01", i)
        if j < 0:
            break
        sc = j - 1 if j > 0 and d[j - 1] == 0 else j
        if nals:
            nals[-1][2] = sc
        nals.append([sc, j + 3, len(d)])
        i = j + 3
    frames, start, have_slice, key, is_open = [], 0, False, False, False
    for sc, payload, end in nals:
        if payload >= end:
            continue
        typ = d[payload] & 0x1F
        boundary = False
        if typ in (6, 7, 8, 9):
            boundary = have_slice
        elif typ in (1, 5):
            boundary = have_slice and payload + 1 < end and bool(d[payload + 1] & 0x80)
        if boundary:
            frames.append((start, sc - start, key))
            is_open = False
        if not is_open:
            is_open, start, have_slice, key = True, sc, False, False
        if typ in (1, 5):
            have_slice = True
        if typ == 5:
            key = True
    if is_open and have_slice:
        frames.append((start, len(d) - start, key))
    return frames
