"""Second, independent restatement of BlobNet inference (TEST INFRASTRUCTURE): plain NumPy loops over kernel taps,
float64, written layer by layer from the reference's Keras model and consuming the weights in KERAS variable layouts
and names - not from oracle/blobnet_ref.py (torch fp32, torch layouts, conv2d / conv_transpose2d library calls).

The two restatements share no code and no tensor layout; tests/test_oracle.py asserts they agree to float32 round-off.
That does not pin the oracle to the reference (no TensorFlow, no trained weights offline - still PARITY UNPINNED), but
it removes "the torch calls mean something else than the Keras layers" as a failure mode, and it is the checker of the
Keras -> CVBN weight converter (tools/keras_to_cvbn.py).

Keras semantics restated here (reference tree):
  utils/model/preprocessing.py:5-8   clip_by_value(x, 0, 6) / 6
  utils/model/encoder.py:33-44       Conv3D(filters, (1,3,3), padding="same", channels_first, bias, relu):
                                     out[n,o,t,y,x] = b[o] + sum_{ky,kx,c} in[n,c,t,y+ky-1,x+kx-1] * K[0,ky,kx,c,o]
                                     (cross-correlation, zeros outside)
  utils/model/encoder.py:45-48       BatchNormalization(axis=1), inference: (x - mean) / sqrt(var + 1e-3) * gamma + beta
  utils/model/encoder.py:50-52       MaxPool3D((1,2,2), valid): floor(H/2) x floor(W/2)
  utils/model/encoder.py:68-76       ZeroPadding3D top 1 row / left 1 column when the PRE-pool extent was odd
  utils/model/pointwise.py:10-26     transpose to [N,C,H,W,T]; Conv1D(4, 1, relu, no bias) twice: out[..,m] = relu(sum_t in[..,t] K[0,t,m]);
                                     transpose back; relu(out + x)
  utils/model/blobnet.py:32          x[:, :, :1] of every encoder output, reversed
  utils/model/decoder.py:9-24        [ReLU] -> Conv3DTranspose(filters, (1,4,4), strides (1,2,2), valid, bias):
                                     out[n,o,t,2i+ky,2j+kx] += in[n,c,t,i,j] * K[0,ky,kx,o,c]; extent 2*in + 2
  utils/model/decoder.py:42-59       Cropping3D((pad//2 + pad%2, pad//2)) per spatial axis
  utils/model/decoder.py:121-134     up -> BatchNorm -> concatenate([x, skip], axis=1) three times; last up without
                                     BN/concat; Conv3D(1, 1) head; sigmoid; blobnet.py:44 squeezes the channel axis
Variable names: Keras numbers layers per type in creation order (conv3d, conv3d_1, ...).  Decoder.__init__ creates a
throw-away Conv3DTranspose after every real one (decoder.py:27-40), so the real ones are conv3d_transpose, _2, _4, _6.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-3
ENC_CH = [(3, 16), (16, 32), (32, 64), (64, 128)]
DEC_CH = [(128, 64), (128, 32), (64, 16), (32, 16)]


def _sfx(i: int) -> str:
    return "" if i == 0 else f"_{i}"


def keras_names():
    """(name, shape) of every variable of the reference model, by Keras' default layer naming."""
    out = []
    for i, (ci, co) in enumerate(ENC_CH):
        out += [(f"conv3d{_sfx(i)}/kernel", (1, 3, 3, ci, co)), (f"conv3d{_sfx(i)}/bias", (co,))]
        out += [(f"batch_normalization{_sfx(i)}/{k}", (co,)) for k in ("gamma", "beta", "moving_mean", "moving_variance")]
        out += [(f"conv1d{_sfx(2 * i)}/kernel", (1, 4, 4)), (f"conv1d{_sfx(2 * i + 1)}/kernel", (1, 4, 4))]
    for i, (ci, co) in enumerate(DEC_CH):
        out += [(f"conv3d_transpose{_sfx(2 * i)}/kernel", (1, 4, 4, co, ci)), (f"conv3d_transpose{_sfx(2 * i)}/bias", (co,))]
        if i < 3:
            out += [(f"batch_normalization{_sfx(4 + i)}/{k}", (co,)) for k in ("gamma", "beta", "moving_mean", "moving_variance")]
    out += [("conv3d_4/kernel", (1, 1, 1, 16, 1)), ("conv3d_4/bias", (1,))]
    return out


def _conv3d_same_133(x, K, b):
    n, c, t, h, w = x.shape
    xp = np.zeros((n, c, t, h + 2, w + 2))
    xp[:, :, :, 1:-1, 1:-1] = x
    out = np.zeros((n, K.shape[-1], t, h, w))
    for ky in range(3):
        for kx in range(3):
            patch = xp[:, :, :, ky: ky + h, kx: kx + w]                    # in[.., y+ky-1, x+kx-1]
            out += np.einsum("nctyx,co->notyx", patch, K[0, ky, kx])
    return out + b.reshape(1, -1, 1, 1, 1)


def _bn(x, v, name):
    g, be, m, var = (v[f"{name}/{k}"].astype(np.float64).reshape(1, -1, 1, 1, 1) for k in ("gamma", "beta", "moving_mean", "moving_variance"))
    return (x - m) / np.sqrt(var + BN_EPS) * g + be


def _maxpool_122(x):
    n, c, t, h, w = x.shape
    out = np.full((n, c, t, h // 2, w // 2), -np.inf)
    for dy in range(2):
        for dx in range(2):
            out = np.maximum(out, x[:, :, :, dy: 2 * (h // 2): 2, dx: 2 * (w // 2): 2])
    return out


def _pointwise_tn(x, k1, k2):
    o = np.transpose(x, (0, 1, 3, 4, 2))                                     # [N,C,H,W,T]
    o = np.maximum(np.einsum("nchwt,tm->nchwm", o, k1[0]), 0.0)
    o = np.maximum(np.einsum("nchwt,tm->nchwm", o, k2[0]), 0.0)
    o = np.transpose(o, (0, 1, 4, 2, 3))
    return np.maximum(o + x, 0.0)


def _conv3d_transpose_144_s122(x, K, b):
    n, c, t, h, w = x.shape
    out = np.zeros((n, K.shape[3], t, 2 * h + 2, 2 * w + 2))
    for ky in range(4):
        for kx in range(4):
            contrib = np.einsum("nctij,oc->notij", x, K[0, ky, kx])          # K is (kd,kh,kw,Cout,Cin)
            out[:, :, :, ky: ky + 2 * h: 2, kx: kx + 2 * w: 2] += contrib    # scatter to (2i+ky, 2j+kx)
    return out + b.reshape(1, -1, 1, 1, 1)


def _crop(x, desired_h, desired_w):
    hp, wp = x.shape[-2] - desired_h, x.shape[-1] - desired_w
    assert hp >= 0 and wp >= 0
    top, bottom, left, right = hp // 2 + hp % 2, hp // 2, wp // 2 + wp % 2, wp // 2
    return x[:, :, :, top: x.shape[-2] - bottom, left: x.shape[-1] - right]


def blobnet_forward_np(v: dict, x: np.ndarray, return_intermediates: bool = False):
    """v: Keras-named variables (keras_names()); x: [N,3,T,H,W] raw byte values.  Returns logits [N,H,W] float64
    (the reference's output is sigmoid(logit))."""
    v = {k: np.asarray(a, dtype=np.float64) for k, a in v.items()}
    x = np.clip(np.asarray(x, dtype=np.float64), 0.0, 6.0) / 6.0
    H, W = x.shape[-2:]
    enc, inter = [], {}
    for i in range(4):
        s = x.shape
        x = np.maximum(_conv3d_same_133(x, v[f"conv3d{_sfx(i)}/kernel"], v[f"conv3d{_sfx(i)}/bias"]), 0.0)
        x = _bn(x, v, f"batch_normalization{_sfx(i)}")
        x = _maxpool_122(x)
        if s[-2] % 2:
            x = np.concatenate([np.zeros(x.shape[:3] + (1, x.shape[4])), x], axis=3)
        if s[-1] % 2:
            x = np.concatenate([np.zeros(x.shape[:4] + (1,)), x], axis=4)
        x = _pointwise_tn(x, v[f"conv1d{_sfx(2 * i)}/kernel"], v[f"conv1d{_sfx(2 * i + 1)}/kernel"])
        enc.append(x)
        inter[f"enc{i}"] = x
    rev = [e[:, :, :1] for e in reversed(enc)]
    shapes = [r.shape for r in rev] + [(None, 3, 4, H, W)]
    x = rev[0]
    for i in range(4):
        x = np.maximum(x, 0.0)
        x = _conv3d_transpose_144_s122(x, v[f"conv3d_transpose{_sfx(2 * i)}/kernel"], v[f"conv3d_transpose{_sfx(2 * i)}/bias"])
        x = _crop(x, shapes[i + 1][-2], shapes[i + 1][-1])
        if i < 3:
            x = _bn(x, v, f"batch_normalization{_sfx(4 + i)}")
            x = np.concatenate([x, rev[i + 1]], axis=1)
        inter[f"dec{i}"] = x[:, :, 0]
    logit = np.einsum("nctyx,co->notyx", x, v["conv3d_4/kernel"][0, 0, 0]) + v["conv3d_4/bias"].reshape(1, -1, 1, 1, 1)
    logit = logit[:, 0, 0]
    return (logit, inter) if return_intermediates else logit
