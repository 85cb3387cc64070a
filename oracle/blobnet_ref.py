"""Oracle restatement of BlobNet inference in torch fp32 (TEST INFRASTRUCTURE).

Follows (reference tree):
  utils/model/preprocessing.py:5-8      clip(x, 0, 6) / 6
  utils/model/encoder.py:33-76          Conv3D(1,3,3 same, bias, ReLU) -> BatchNorm(axis=1, eps 1e-3)
                                        -> MaxPool3D(1,2,2 valid) -> zero-pad top/left when the
                                        pre-pool extent is odd -> PointWiseTN
  utils/model/pointwise.py:10-26        two Conv1D(4, k=1, ReLU, no bias) over T, residual, ReLU
  utils/model/blobnet.py:32             skips = x[:, :, :1] of every encoder output, reversed
  utils/model/decoder.py:5-64,106-134   [ReLU -> Conv3DTranspose(1,4,4 stride 1,2,2 valid, bias)
                                        -> crop (ceil, floor)] -> BatchNorm -> concat, x3;
                                        last block without BN/concat; Conv3D(1,1) head; sigmoid
  utils/train-blobnet.py:57-69          channel plan 16/32/64/128, decoder 64/32/16/16, T = 4
  utils/train-blobnet.py:113-116        deployment Reshape (3, T*H, W) -> (3, T, H, W)

PARITY UNPINNED: neither TensorFlow nor trained weights / golden tensors exist offline
(SURVEY.md section 8c).  Dropout layers are inference-inert and omitted.

The weight blob ("CVBN" v1) is the canonical fp32 container both this oracle and the
CUDA library parse; tensors are stored in torch layouts (Conv: [Cout,Cin,kh,kw],
ConvTranspose: [Cin,Cout,kh,kw], Conv1D over T as a [T_in,T_out] matrix).
"""
from __future__ import annotations

import struct

import numpy as np
import torch
import torch.nn.functional as F

MAGIC = 0x4E425643  # 'CVBN' little endian
VERSION = 1
T = 4
ENC_CH = [(3, 16), (16, 32), (32, 64), (64, 128)]
DEC_CH = [(128, 64), (128, 32), (64, 16), (32, 16)]
BN_EPS = 1e-3  # Keras BatchNormalization default


def blob_schema():
    """Ordered (name, shape) list of the CVBN v1 container."""
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


def parse_blob(blob: bytes) -> dict[str, np.ndarray]:
    magic, version, t, _ = struct.unpack_from("<IIII", blob, 0)
    if magic != MAGIC or version != VERSION or t != T:
        raise ValueError("not a CVBN v1 weight blob")
    off = 16
    out = {}
    for name, shape in blob_schema():
        n = int(np.prod(shape))
        out[name] = np.frombuffer(blob, dtype="<f4", count=n, offset=off).reshape(shape).copy()
        off += 4 * n
    if off != len(blob):
        raise ValueError(f"blob length {len(blob)} != expected {off}")
    return out


def _bn(x, w, p):
    g, b, m, v = (torch.from_numpy(w[f"{p}.bn_{k}"]) for k in ("gamma", "beta", "mean", "var"))
    scale = g / torch.sqrt(v + BN_EPS)
    return x * scale.view(1, -1, 1, 1) + (b - m * scale).view(1, -1, 1, 1)


def _pointwise_tn(x, w1, w2):
    """x: [N, C, T, H, W]; Conv1D acts on the last axis after the transpose (pointwise.py:18-26)."""
    xt = x.permute(0, 1, 3, 4, 2)                               # [N,C,H,W,T]
    h = torch.relu(xt @ torch.from_numpy(w1))                   # [.., T_in] @ [T_in, T_out]
    h = torch.relu(h @ torch.from_numpy(w2))
    return torch.relu(h.permute(0, 1, 4, 2, 3) + x)


def _crop_amounts(res: int, desired: int) -> tuple[int, int]:
    pad = res - desired                                          # decoder.py:42-59
    assert pad >= 0, "ZeroPadding branch never triggers for ceil-halved skips"
    return pad // 2 + pad % 2, pad // 2


@torch.no_grad()
def blobnet_forward(w: dict[str, np.ndarray], x_nchw: np.ndarray, return_intermediates: bool = False):
    """x_nchw: float32 [N, 3, T, H, W] raw byte values.  Returns logits float32 [N, H, W]
    (prob = sigmoid(logit); blobnet.py:44 squeezes the channel axis)."""
    x = torch.from_numpy(np.ascontiguousarray(x_nchw, dtype=np.float32))
    N, C, Tn, H, W = x.shape
    assert C == 3 and Tn == T
    inter = {}
    x = torch.clamp(x, 0.0, 6.0) / 6.0
    skips = []
    for i in range(4):
        n, c, t, h, wd = x.shape
        y = x.permute(0, 2, 1, 3, 4).reshape(n * t, c, h, wd)    # kernel depth 1 -> per-frame conv2d
        y = F.conv2d(y, torch.from_numpy(w[f"enc{i}.conv_w"]), torch.from_numpy(w[f"enc{i}.conv_b"]), padding=1)
        y = torch.relu(y)
        y = _bn(y, w, f"enc{i}")
        y = F.max_pool2d(y, 2)                                   # valid: floors
        if h % 2:
            y = F.pad(y, (0, 0, 1, 0))                           # one zero row on top   (encoder.py:68-71)
        if wd % 2:
            y = F.pad(y, (1, 0, 0, 0))                           # one zero column left  (encoder.py:72-75)
        y = y.reshape(n, t, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
        x = _pointwise_tn(y, w[f"enc{i}.tn_w1"], w[f"enc{i}.tn_w2"])
        inter[f"enc{i}"] = x
        skips.append(x[:, :, 0])                                 # T index 0 only (blobnet.py:32)
    x = skips[3]
    targets = [skips[2].shape[-2:], skips[1].shape[-2:], skips[0].shape[-2:], (H, W)]
    for i in range(4):
        x = torch.relu(x)
        x = F.conv_transpose2d(x, torch.from_numpy(w[f"dec{i}.convt_w"]), torch.from_numpy(w[f"dec{i}.convt_b"]), stride=2)
        ct, cb = _crop_amounts(x.shape[-2], targets[i][0])
        cl, cr = _crop_amounts(x.shape[-1], targets[i][1])
        x = x[:, :, ct: x.shape[-2] - cb, cl: x.shape[-1] - cr]
        if i < 3:
            x = _bn(x, w, f"dec{i}")
            x = torch.cat([x, skips[2 - i]], dim=1)
        inter[f"dec{i}"] = x
    logit = (x * torch.from_numpy(w["head_w"]).view(1, -1, 1, 1)).sum(1) + float(w["head_b"][0])
    if return_intermediates:
        return logit.numpy(), {k: v.numpy() for k, v in inter.items()}
    return logit.numpy()


def mask_from_logits(logit: np.ndarray) -> np.ndarray:
    """nvinfer segmentation-threshold 0.5, strict '>' (config/blobnet/amsterdam_b128.txt:26) gives
    class_map in {-1, 0}; maskcopy writes class_map + 1 (gstmaskcopy.cpp:226-230).
    sigmoid(z) > 0.5  <=>  z > 0."""
    return (logit > 0).astype(np.uint8)


def flops_per_window(H: int, W: int) -> int:
    """2*MAC of the network at an H x W macroblock grid."""
    mac = 0
    h, wd = H, W
    sizes = []
    for ci, co in ENC_CH:
        mac += T * h * wd * 9 * ci * co
        h, wd = (h + 1) // 2, (wd + 1) // 2
        mac += co * h * wd * 2 * T * T
        sizes.append((h, wd))
    for i, (ci, co) in enumerate(DEC_CH):
        mac += h * wd * 16 * ci * co
        h, wd = sizes[2 - i] if i < 3 else (H, W)
    mac += H * W * DEC_CH[-1][1]
    return 2 * mac
