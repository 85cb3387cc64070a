#!/usr/bin/env python
"""Writes tests/golden/demo_1m_moov.bin: the ftyp + moov boxes of the reference's demo/1m.mp4 (22 KB of sample
tables, no media data), used by tests/test_demux.py to pin the MP4 sample-table reader on the real file.
Run in the build container (needs /root/reference); the GPU box only sees the committed fixture."""
import os
import struct
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/demo/1m.mp4"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "demo_1m_moov.bin")
d = open(SRC, "rb").read()
out, off = b"", 0
while off < len(d):
    size, typ = struct.unpack_from(">I4s", d, off)
    if size == 1:
        (size,) = struct.unpack_from(">Q", d, off + 8)
    if typ in (b"ftyp", b"moov"):
        out += d[off: off + size]
    off += size
open(DST, "wb").write(out)
print(DST, len(out), "bytes")
