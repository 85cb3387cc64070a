"""Work partitioning over the GPUs of one box.  The path shards into independent units (a window depends
only on `timestep` consecutive frames of one chain), so there is NO data-path collective; the optional
gather of the (tiny, variable-length) box blobs is host-side.

Reference behaviour mirrored here:
  * gopsplit (gst-plugins/gst-gopsplit/gstgopsplit.cpp:557-630): pad i of P gets the contiguous GoP range
    [i*floor(G/P), (i+1)*floor(G/P)), the remainder goes to the LAST pad; every chain then starts with an
    empty window, so its first timestep-1 frames emit nothing (metapreprocess/imp.rs:302-305).
  * many-stream configurations: stream s -> GPU s mod P.
"""
from __future__ import annotations


def gop_ranges(n_gops: int, n_pads: int) -> list[tuple[int, int]]:
    """[start, end) GoP range of every pad, exactly as gopsplit assigns them."""
    if n_pads < 1:
        raise ValueError("there are no pads")
    if n_gops < n_pads:                     # "Too many pads": pad i pushes GoP i (gstgopsplit.cpp:531-553)
        return [(i, i + 1) if i < n_gops else (0, 0) for i in range(n_pads)]
    per = n_gops // n_pads
    out = []
    for i in range(n_pads):
        s, e = i * per, (i + 1) * per
        if i == n_pads - 1:
            e = n_gops                      # remaining GoPs are pushed to the last pad
        out.append((s, e))
    return out


def bind_rank_to_gpu(device: int) -> dict:
    """One process per GPU: keep this process (and the pinned frame/box buffers it allocates from now on) on the CPUs of
    the NUMA node `device` hangs off (cova_bind_host_to_device).  Returns {"numa_node", "n_cpus"}; n_cpus == 0 means the
    affinity was left alone (no topology information, or nothing left inside the launcher's own CPU set)."""
    import ctypes

    from . import _lib
    node, ncpu = ctypes.c_int(-1), ctypes.c_int(0)
    _lib.check(_lib.load().cova_bind_host_to_device(int(device), ctypes.byref(node), ctypes.byref(ncpu)))
    return {"numa_node": node.value, "n_cpus": ncpu.value}


def gop_starts(is_keyframe: list[bool]) -> list[int]:
    """Frame index at which every GoP starts (a buffer without DELTA_UNIT; gstgopsplit.cpp:712-723).  Delta
    frames ahead of the first key frame form a GoP of their own."""
    return [i for i, k in enumerate(is_keyframe) if k or i == 0]


def demux_mp4(data: bytes):
    """(samples, info) of the first H.264 track: samples = list of (offset, size, is_key, dts, pts)."""
    import ctypes

    import numpy as np

    from . import _lib
    lib = _lib.load()
    a = np.frombuffer(data, dtype=np.uint8)
    n, info = ctypes.c_size_t(), _lib.Mp4Info()
    rc = lib.cova_demux_mp4_samples(a.ctypes.data, a.size, None, 0, ctypes.byref(n), ctypes.byref(info))
    if rc not in (_lib.OK, _lib.E_TOOSMALL):
        _lib.check(rc)
    out = (_lib.Sample * max(1, n.value))()
    _lib.check(lib.cova_demux_mp4_samples(a.ctypes.data, a.size, out, n.value, ctypes.byref(n), ctypes.byref(info)))
    return ([(s.offset, s.size, not s.flags & 1, s.dts, s.pts) for s in out[: n.value]],
            dict(timescale=info.timescale, width=info.width, height=info.height, nal_length_size=info.nal_length_size))


def demux_annexb(data: bytes):
    """Access units of an Annex-B stream: list of (offset, size, is_key)."""
    import ctypes

    import numpy as np

    from . import _lib
    lib = _lib.load()
    a = np.frombuffer(data, dtype=np.uint8)
    n = ctypes.c_size_t()
    rc = lib.cova_demux_annexb_frames(a.ctypes.data, a.size, None, 0, ctypes.byref(n))
    if rc not in (_lib.OK, _lib.E_TOOSMALL):
        _lib.check(rc)
    out = (_lib.Sample * max(1, n.value))()
    _lib.check(lib.cova_demux_annexb_frames(a.ctypes.data, a.size, out, n.value, ctypes.byref(n)))
    return [(s.offset, s.size, not s.flags & 1) for s in out[: n.value]]


def gopsplit_ranges(is_keyframe, n_pads: int) -> list[tuple[int, int]]:
    """[first_frame, end_frame) per pad, computed by the native library (cova_gopsplit_ranges)."""
    import ctypes

    import numpy as np

    from . import _lib
    flags = np.array([0 if k else 1 for k in is_keyframe], dtype=np.uint32)
    first, end = np.zeros(n_pads, dtype=np.uint64), np.zeros(n_pads, dtype=np.uint64)
    _lib.check(_lib.load().cova_gopsplit_ranges(flags.ctypes.data, flags.size, n_pads, first.ctypes.data, end.ctypes.data))
    return [(int(a), int(b)) for a, b in zip(first, end)]


def frames_of_shard(is_keyframe: list[bool], n_pads: int, pad: int) -> tuple[int, int]:
    """[first_frame, end_frame) handed to `pad` when a stream with these key frames is split."""
    starts = gop_starts(is_keyframe)
    if not starts:
        return (0, 0)
    s, e = gop_ranges(len(starts), n_pads)[pad]
    if s == e:
        return (0, 0)
    first = starts[s]
    end = starts[e] if e < len(starts) else len(is_keyframe)
    return (first, end)


def streams_of_rank(n_streams: int, world: int, rank: int) -> list[int]:
    return list(range(rank, n_streams, world))


def windows_of_chain(n_frames: int, timestep: int, gamma: int = 1) -> int:
    return 0 if n_frames < timestep else (n_frames - timestep) // gamma + 1


def gather_blobs(local: dict[int, list[bytes]], world: int):
    """Optional final gather of per-stream box blobs to rank 0 (host-side; negligible bytes)."""
    if world == 1:
        return dict(local)
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, local)
    merged = {}
    for d in out:
        merged.update(d)
    return merged
