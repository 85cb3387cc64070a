"""CPU tests of the host logic and the C-ABI surface (no GPU compute)."""
import ctypes
import os
import re

import numpy as np
import pytest

from cova_b200 import _lib, shard, synth, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cova_b200.h")).read()
    declared = set(re.findall(r"\b(cova_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cova_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.cova_version()


def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    n = ctypes.c_int(-1)
    assert lib.cova_device_count(ctypes.byref(n)) == 0
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    for rc in (lib.cova_bboxcc_new(ctypes.byref(h), 0, 80, 45, 1),
               lib.cova_metapreprocess_new(ctypes.byref(h), 0, 1280, 720, 4, 1)):
        assert rc == _lib.E_NODEVICE and not h.value
    blob = weights.to_blob(weights.random_weights(0))
    buf = ctypes.create_string_buffer(blob, len(blob))
    rc = lib.cova_pipeline_new(ctypes.byref(h), 0, 80, 45, 4, 1, 1, 8, ctypes.cast(buf, ctypes.c_void_p), len(blob), 1, 0)
    assert rc == _lib.E_NODEVICE
    assert b"no CPU fallback" in lib.cova_last_error()
    node, ncpu = ctypes.c_int(7), ctypes.c_int(7)
    assert lib.cova_bind_host_to_device(0, ctypes.byref(node), ctypes.byref(ncpu)) == _lib.E_NODEVICE
    assert (node.value, ncpu.value) == (-1, 0)


def test_argument_validation_comes_before_device_use():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.cova_metapreprocess_new(ctypes.byref(h), 0, 1280, 720, 0, 1) == _lib.E_INVAL      # timestep >= 1
    assert lib.cova_bboxcc_new(ctypes.byref(h), 0, 0, 45, 1) == _lib.E_INVAL
    assert lib.cova_pipeline_new(ctypes.byref(h), 0, 80, 45, 3, 1, 1, 8, None, 0, 1, 0) == _lib.E_UNSUPPORTED
    assert lib.cova_strerror(_lib.E_TOOSMALL) == b"output buffer too small"


def test_weight_container_roundtrip_and_size():
    w = weights.random_weights(3)
    blob = weights.to_blob(w)
    assert len(blob) == 16 + 4 * weights.n_params()
    assert 300_000 < weights.n_params() < 340_000          # "~320 K parameters"
    w2 = weights.from_blob(blob)
    for k in w:
        assert (w[k] == w2[k]).all()
    from oracle import blobnet_ref
    w3 = blobnet_ref.parse_blob(blob)
    assert all((w[k] == w3[k]).all() for k in w)
    with pytest.raises(ValueError):
        weights.from_blob(blob[:-4])


def test_flops_per_window_match_survey():
    from oracle import blobnet_ref
    assert blobnet_ref.flops_per_window(45, 80) == 153_786_880
    assert blobnet_ref.flops_per_window(68, 120) == 340_823_040
    assert blobnet_ref.flops_per_window(135, 240) == 1_333_946_880


def test_gopsplit_ranges_remainder_to_last_pad():
    assert shard.gop_ranges(8, 3) == [(0, 2), (2, 4), (4, 8)]
    # fewer GoPs than pads: pad i pushes GoP i, the other pads nothing (gstgopsplit.cpp:531-553)
    assert shard.gop_ranges(7, 8) == [(i, i + 1) for i in range(7)] + [(0, 0)]
    key = [i % 250 == 0 for i in range(1802)]              # demo/1m.mp4: 1802 frames, 8 key frames
    spans = [shard.frames_of_shard(key, 4, p) for p in range(4)]
    assert spans == [(0, 500), (500, 1000), (1000, 1500), (1500, 1802)]
    # every chain restarts its window: 3 frames per shard emit nothing
    assert sum(shard.windows_of_chain(e - s, 4) for s, e in spans) == 1802 - 4 * 3


def test_stream_sharding_is_a_partition():
    for world in (1, 2, 4, 8):
        got = sorted(s for r in range(world) for s in shard.streams_of_rank(256, world, r))
        assert got == list(range(256))


def test_synthetic_stream_statistics():
    fr = synth.synth_streams(2, 67, 45, 80, config_idx=1)
    assert fr.shape == (2, 67, 45, 80, 4) and fr.dtype == np.uint8
    assert (fr[:, 0, :, :, 0] == 6).all()                  # frame 0 is the IDR: all intra
    assert fr[..., 3].max() <= 4                           # stale byte
    assert (fr[:, 1:, :, :, 0] == 1).mean() > 0.8          # mostly skip MBs


def test_validation_kernels_are_not_in_the_shipped_library():
    """libcova_b200.so holds the hot path only: it refuses COVA_IMPL_SIMT; libcova_b200_val.so (same sources,
    -DCOVA_VALIDATION) is the build the layer-by-layer parity tests load.  Both export the whole ABI."""
    blob = weights.to_blob(weights.random_weights(0))
    buf = ctypes.create_string_buffer(blob, len(blob))
    h = ctypes.c_void_p()
    prod, val = _lib.load(), _lib.load_validation()
    assert prod is not val
    rc = prod.cova_pipeline_new(ctypes.byref(h), 0, 80, 45, 4, 1, 1, 8, ctypes.cast(buf, ctypes.c_void_p), len(blob), 1, _lib.IMPL_SIMT)
    assert rc == _lib.E_UNSUPPORTED and b"libcova_b200_val.so" in prod.cova_last_error()
    rc = val.cova_pipeline_new(ctypes.byref(h), 0, 80, 45, 4, 1, 1, 8, ctypes.cast(buf, ctypes.c_void_p), len(blob), 1, _lib.IMPL_SIMT)
    assert rc in (_lib.OK, _lib.E_NODEVICE)                # accepted; fails later only for want of a device
    if rc == _lib.OK:
        val.cova_pipeline_free(h)
    for name in _lib.SIGNATURES:
        assert hasattr(val, name)


def test_stream_batch_argument_validation():
    lib = _lib.load()
    assert lib.cova_pipeline_submit_host2(None, None, 1, 1, None, None, 0) == _lib.E_INVAL
    n = ctypes.c_size_t()
    assert lib.cova_pipeline_collect_host2(None, None, 0, ctypes.byref(n), None, None, None, None, None) == _lib.E_INVAL
    assert lib.cova_pipeline_reset_streams(None, None, 0) == _lib.E_INVAL


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_frame_packer_is_exact(threads):
    """cova_packer_pack (AVX2 or scalar, any thread count) == min(b, 6) of bytes 0..2 in 3 bits each; byte 3 ignored;
    odd lengths and lengths below the threading threshold included."""
    from cova_b200.elements import FramePacker
    rng = np.random.default_rng(threads)
    pk = FramePacker(threads)
    for n in (0, 1, 15, 16, 17, 3600, 65535, 65536, 128 * 3600 + 7):
        q = rng.integers(0, 256, (n, 4), dtype=np.uint8)
        q[: n // 2, :3] = rng.integers(0, 8, (n // 2, 3), dtype=np.uint8)        # the realistic range, with 7 in it
        want = (np.minimum(q[:, 0], 6).astype(np.uint16) | (np.minimum(q[:, 1], 6).astype(np.uint16) << 3)
                | (np.minimum(q[:, 2], 6).astype(np.uint16) << 6))
        got = pk.pack(q.reshape(n, 1, 4)).reshape(-1) if n else pk.pack(q.reshape(0, 1, 4)).reshape(-1)
        assert (got == want).all(), n
    fr = synth.synth_streams(2, 5, 45, 80, config_idx=1)
    assert pk.pack(fr).shape == fr.shape[:-1]


def test_python_constants_match_the_header():
    hdr = open(os.path.join(ROOT, "include", "cova_b200.h")).read()

    def define(name):
        m = re.search(r"#define\s+" + name + r"\s+\(?(-?0x[0-9a-fA-F]+|-?\d+)u?\)?", hdr)
        assert m, name
        return int(m.group(1), 0)

    assert define("COVA_FLAG_KEEP_LOGITS") == _lib.FLAG_KEEP_LOGITS and define("COVA_FLAG_KEEP_STACKED") == _lib.FLAG_KEEP_STACKED
    assert define("COVA_FLAG_INPUT_PACKED16") == _lib.FLAG_INPUT_PACKED16 and define("COVA_SUBMIT_CONTINUE") == _lib.SUBMIT_CONTINUE
    assert define("COVA_IMPL_TCGEN05") == _lib.IMPL_TCGEN05 and define("COVA_IMPL_SIMT") == _lib.IMPL_SIMT
    for name, val in (("OK", _lib.OK), ("DROPPED", _lib.DROPPED), ("E_INVAL", _lib.E_INVAL), ("E_CUDA", _lib.E_CUDA), ("E_NOMEM", _lib.E_NOMEM),
                      ("E_TOOSMALL", _lib.E_TOOSMALL), ("E_WEIGHTS", _lib.E_WEIGHTS), ("E_UNSUPPORTED", _lib.E_UNSUPPORTED),
                      ("E_NODEVICE", _lib.E_NODEVICE), ("E_NUMERIC", _lib.E_NUMERIC), ("E_STATE", _lib.E_STATE)):
        assert define("COVA_" + name) == val, name
    # packed input is refused for odd widths and together with KEEP_STACKED before any device is touched
    lib = _lib.load()
    blob = weights.to_blob(weights.random_weights(0))
    buf = ctypes.create_string_buffer(blob, len(blob))
    h = ctypes.c_void_p()
    args = (ctypes.byref(h), 0, 37, 21, 4, 1, 1, 8, ctypes.cast(buf, ctypes.c_void_p), len(blob), 1)
    assert lib.cova_pipeline_new(*args, _lib.FLAG_INPUT_PACKED16) == _lib.E_UNSUPPORTED
    args = (ctypes.byref(h), 0, 80, 45, 4, 1, 1, 8, ctypes.cast(buf, ctypes.c_void_p), len(blob), 1)
    assert lib.cova_pipeline_new(*args, _lib.FLAG_INPUT_PACKED16 | _lib.FLAG_KEEP_STACKED) == _lib.E_INVAL
