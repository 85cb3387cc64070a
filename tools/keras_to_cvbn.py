#!/usr/bin/env python
"""Trained BlobNet weights (Keras variables) -> CVBN v1 container (cova_b200/weights.py) for cova_pipeline_new().

The reference trains BlobNet in Keras and ships it as SavedModel -> ONNX -> TensorRT (utils/train-blobnet.py:101-119,
model/tasks.py:17-30); no weights are in the tree.  A maintainer who has the SavedModel dumps its variables with

    m = tf.keras.models.load_model("model/tf_model/...")         # or the un-wrapped BlobNet model
    np.savez("blobnet_vars.npz", **{v.name: v.numpy() for v in m.variables})

and runs   python tools/keras_to_cvbn.py blobnet_vars.npz blobnet.cvbn

Mapping (Keras layout -> CVBN / torch layout):
    Conv3D kernel          (kd=1, 3, 3, Cin, Cout)  -> conv_w [Cout][Cin][3][3]
    Conv3DTranspose kernel (kd=1, 4, 4, Cout, Cin)  -> convt_w[Cin][Cout][4][4]   (both are gradient-of-correlation: no flip)
    Conv1D kernel          (1, T_in, T_out)         -> tn_w   [T_in][T_out]
    BatchNormalization     gamma, beta, moving_mean, moving_variance (eps 1e-3 is applied by the consumer)
    final Conv3D(1, 1)     (1, 1, 1, 16, 1), bias   -> head_w[16], head_b[1]
Variables are matched by layer TYPE and by the rank of the layer's numeric suffix within its type (Keras numbers layers
per type in creation order; Decoder.__init__ burns every second conv3d_transpose index on a throw-away layer,
utils/model/decoder.py:27-40), then checked against the architecture's shapes - so prefixes such as
"model/encoder/sequential_3/" and the ":0" suffix do not matter.  Dropout has no variables.
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cova_b200 import weights  # noqa: E402

_PAT = re.compile(r"(conv3d_transpose|conv3d|conv1d|batch_normalization)(?:_(\d+))?/(kernel|bias|gamma|beta|moving_mean|moving_variance)(?::\d+)?$")


def keras_to_cvbn(variables: dict) -> dict:
    """variables: Keras variable name -> array.  Returns the CVBN tensor dict (weights.schema())."""
    by_type: dict = {}
    for name, arr in variables.items():
        m = _PAT.search(name)
        if not m:
            raise ValueError(f"unrecognised variable name {name!r}")
        by_type.setdefault(m.group(1), {}).setdefault(int(m.group(2) or 0), {})[m.group(3)] = np.asarray(arr, dtype=np.float32)

    def layers(kind, n):
        idx = sorted(by_type.get(kind, {}))
        if len(idx) != n:
            raise ValueError(f"expected {n} {kind} layers with variables, found {len(idx)}")
        return [by_type[kind][i] for i in idx]

    conv3d, bn, conv1d, convt = layers("conv3d", 5), layers("batch_normalization", 7), layers("conv1d", 8), layers("conv3d_transpose", 4)
    w = {}

    def take(src, key, shape, what):
        a = src[key]
        if a.shape != tuple(shape):
            raise ValueError(f"{what}: shape {a.shape}, expected {tuple(shape)}")
        return a

    for i, (ci, co) in enumerate(weights.ENC_CH):
        w[f"enc{i}.conv_w"] = take(conv3d[i], "kernel", (1, 3, 3, ci, co), f"conv3d #{i} kernel")[0].transpose(3, 2, 0, 1)
        w[f"enc{i}.conv_b"] = take(conv3d[i], "bias", (co,), f"conv3d #{i} bias")
        for k, kk in (("gamma", "gamma"), ("beta", "beta"), ("mean", "moving_mean"), ("var", "moving_variance")):
            w[f"enc{i}.bn_{k}"] = take(bn[i], kk, (co,), f"batch_normalization #{i} {kk}")
        w[f"enc{i}.tn_w1"] = take(conv1d[2 * i], "kernel", (1, 4, 4), f"conv1d #{2 * i} kernel")[0]
        w[f"enc{i}.tn_w2"] = take(conv1d[2 * i + 1], "kernel", (1, 4, 4), f"conv1d #{2 * i + 1} kernel")[0]
    for i, (ci, co) in enumerate(weights.DEC_CH):
        w[f"dec{i}.convt_w"] = take(convt[i], "kernel", (1, 4, 4, co, ci), f"conv3d_transpose #{i} kernel")[0].transpose(3, 2, 0, 1)
        w[f"dec{i}.convt_b"] = take(convt[i], "bias", (co,), f"conv3d_transpose #{i} bias")
        if i < 3:
            for k, kk in (("gamma", "gamma"), ("beta", "beta"), ("mean", "moving_mean"), ("var", "moving_variance")):
                w[f"dec{i}.bn_{k}"] = take(bn[4 + i], kk, (co,), f"batch_normalization #{4 + i} {kk}")
    w["head_w"] = take(conv3d[4], "kernel", (1, 1, 1, 16, 1), "final conv3d kernel")[0, 0, 0, :, 0]
    w["head_b"] = take(conv3d[4], "bias", (1,), "final conv3d bias")
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in w.items()}


def main(argv):
    if len(argv) != 3:
        sys.exit(__doc__)
    with np.load(argv[1]) as z:
        blob = weights.to_blob(keras_to_cvbn({k: z[k] for k in z.files}))
    with open(argv[2], "wb") as f:
        f.write(blob)
    print(f"{argv[2]}: {len(blob)} bytes, {weights.n_params()} parameters")


if __name__ == "__main__":
    main(sys.argv)
