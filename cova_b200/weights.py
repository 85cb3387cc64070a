"""BlobNet weight container ("CVBN" v1) and the seeded synthetic initialiser.

No trained weights ship with the reference (model/tf_model/.placeholder, README.md:198 points at a
Google-Drive link), so benchmarks and tests use random-init weights of the reference architecture
(utils/train-blobnet.py:57-69): He-normal conv kernels (encoder.py:41, decoder.py:20), BatchNorm
statistics drawn so that gamma can be exercised on both sides of 1, final bias tuned by the caller.

Container layout (little endian), parsed by csrc/weights_pack.cuh and by the oracle:
    u32 magic 'CVBN' (0x4E425643), u32 version = 1, u32 timestep = 4, u32 reserved = 0, then fp32
    tensors in torch layouts, in this order:
      for e in 0..3 (Cin,Cout = 3->16->32->64->128):
          conv_w[Cout][Cin][3][3], conv_b[Cout], bn_gamma, bn_beta, bn_mean, bn_var [Cout],
          tn_w1[4][4], tn_w2[4][4]            (Conv1D over T as [T_in][T_out])
      for d in 0..3 (Cin,Cout = 128->64, 128->32, 64->16, 32->16):
          convt_w[Cin][Cout][4][4], convt_b[Cout], and for d < 3: bn_gamma, bn_beta, bn_mean, bn_var
      head_w[16], head_b[1]
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = 0x4E425643
VERSION = 1
T = 4
ENC_CH = [(3, 16), (16, 32), (32, 64), (64, 128)]
DEC_CH = [(128, 64), (128, 32), (64, 16), (32, 16)]


def schema():
    s = []
    for i, (ci, co) in enumerate(ENC_CH):
        s += [(f"enc{i}.conv_w", (co, ci, 3, 3)), (f"enc{i}.conv_b", (co,)),
              (f"enc{i}.bn_gamma", (co,)), (f"enc{i}.bn_beta", (co,)),
              (f"enc{i}.bn_mean", (co,)), (f"enc{i}.bn_var", (co,)),
              (f"enc{i}.tn_w1", (T, T)), (f"enc{i}.tn_w2", (T, T))]
    for i, (ci, co) in enumerate(DEC_CH):
        s += [(f"dec{i}.convt_w", (ci, co, 4, 4)), (f"dec{i}.convt_b", (co,))]
        if i < 3:
            s += [(f"dec{i}.bn_gamma", (co,)), (f"dec{i}.bn_beta", (co,)),
                  (f"dec{i}.bn_mean", (co,)), (f"dec{i}.bn_var", (co,))]
    s += [("head_w", (DEC_CH[-1][1],)), ("head_b", (1,))]
    return s


def n_params() -> int:
    return sum(int(np.prod(shape)) for _, shape in schema())


def random_weights(seed: int = 0, head_bias: float = 0.0) -> dict[str, np.ndarray]:
    """He-normal kernels, BN gamma~U(0.5,1.5), small beta/mean, var~U(0.5,1.5)."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in schema():
        kind = name.split(".")[-1]
        if kind == "conv_w":
            fan_in = shape[1] * shape[2] * shape[3]
            v = rng.normal(0.0, np.sqrt(2.0 / fan_in), shape)
        elif kind == "convt_w":
            # Keras he_normal on a Conv3DTranspose kernel (kd,kh,kw,Cout,Cin): fan_in = kh*kw*Cout
            fan_in = shape[1] * shape[2] * shape[3]
            v = rng.normal(0.0, np.sqrt(2.0 / fan_in), shape)
        elif kind in ("conv_b", "convt_b", "bn_beta", "bn_mean"):
            v = rng.normal(0.0, 0.05, shape)
        elif kind in ("bn_gamma", "bn_var"):
            v = rng.uniform(0.5, 1.5, shape)
        elif kind in ("tn_w1", "tn_w2"):
            v = rng.normal(0.0, 0.5, shape)          # glorot-ish for a 4x4 matrix, both signs
        elif name == "head_w":
            v = rng.normal(0.0, np.sqrt(2.0 / shape[0]), shape)
        elif name == "head_b":
            v = np.full(shape, head_bias)
        else:
            raise AssertionError(name)
        w[name] = np.ascontiguousarray(v, dtype=np.float32)
    return w


def to_blob(w: dict[str, np.ndarray]) -> bytes:
    out = bytearray(struct.pack("<IIII", MAGIC, VERSION, T, 0))
    for name, shape in schema():
        a = np.ascontiguousarray(w[name], dtype="<f4")
        if a.shape != tuple(shape):
            raise ValueError(f"{name}: shape {a.shape} != {shape}")
        out += a.tobytes()
    return bytes(out)


def from_blob(blob: bytes) -> dict[str, np.ndarray]:
    magic, version, t, _ = struct.unpack_from("<IIII", blob, 0)
    if magic != MAGIC or version != VERSION or t != T:
        raise ValueError("not a CVBN v1 weight blob")
    off = 16
    w = {}
    for name, shape in schema():
        n = int(np.prod(shape))
        w[name] = np.frombuffer(blob, dtype="<f4", count=n, offset=off).reshape(shape).copy()
        off += 4 * n
    if off != len(blob):
        raise ValueError("blob length mismatch")
    return w
