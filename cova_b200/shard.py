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
    per = n_gops // n_pads
    out = []
    for i in range(n_pads):
        s, e = i * per, (i + 1) * per
        if i == n_pads - 1:
            e = n_gops                      # remaining GoPs are pushed to the last pad
        out.append((s, e))
    return out


def gop_starts(is_keyframe: list[bool]) -> list[int]:
    """Frame index at which every GoP starts (a buffer without DELTA_UNIT; gstgopsplit.cpp:712-723)."""
    return [i for i, k in enumerate(is_keyframe) if k]


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
