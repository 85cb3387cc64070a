"""CPU tests: the two independent BlobNet restatements of the oracle against each other, and the Keras -> CVBN weight
converter (tools/keras_to_cvbn.py).

oracle/blobnet_ref.py  torch fp32, torch layouts, library conv2d / conv_transpose2d
oracle/blobnet_np.py   NumPy float64 loops over kernel taps, Keras layouts and variable names, scatter-form transposed
                       convolution, BatchNorm in its (x - mean) / sqrt(var + eps) * gamma + beta form
Neither is pinned to the reference (no TensorFlow / trained weights offline); agreeing to float32 round-off rules out
that one of them mis-states a Keras layer (kernel flip, pad side, crop side, Conv1D axis, BN epsilon ...)."""
import importlib.util
import os

import numpy as np
import pytest

from cova_b200 import synth, weights
from oracle import blobnet_np, blobnet_ref, metapreprocess_ref as mpr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("keras_to_cvbn", os.path.join(ROOT, "tools", "keras_to_cvbn.py"))
k2c = importlib.util.module_from_spec(spec)
spec.loader.exec_module(k2c)


def keras_variables(seed: int, head_bias: float = -0.3) -> dict:
    """Random variables in KERAS layouts under the names a SavedModel of the reference would carry (nested prefixes, ':0'),
    inserted in a scrambled order.  Asymmetric on purpose: every axis of every kernel is distinguishable."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in blobnet_np.keras_names():
        kind = name.split("/")[-1]
        if kind == "kernel":
            fan = int(np.prod(shape[:-1])) if "transpose" not in name else int(np.prod(shape[:3])) * shape[3]
            a = rng.normal(0.0, np.sqrt(2.0 / max(fan, 1)), shape) if len(shape) > 3 else rng.normal(0.0, 0.5, shape)
        elif kind in ("gamma", "moving_variance"):
            a = rng.uniform(0.5, 1.5, shape)
        else:
            a = rng.normal(0.0, 0.05, shape)
        if name == "conv3d_4/bias":
            a = np.full(shape, head_bias)
        out[name] = a.astype(np.float32)
    out["batch_normalization_5/gamma"][::4] *= -1.0                      # negative BN scales must survive too
    names = list(out)
    rng.shuffle(names)
    prefix = {"conv3d": "model/encoder/sequential", "conv1d": "model/encoder/point_wise_tn/sequential", "batch_normalization": "model",
              "conv3d_transpose": "model/decoder/sequential"}
    return {f"{prefix[n.split('/')[0].rstrip('_0123456789')]}/{n}:0": out[n] for n in names}


@pytest.mark.parametrize("h,w,n,seed", [(45, 80, 1, 0), (21, 37, 2, 1), (16, 16, 2, 2), (34, 23, 1, 3)])
def test_numpy_keras_restatement_equals_torch_oracle_through_the_converter(h, w, n, seed):
    kv = keras_variables(seed)
    cv = k2c.keras_to_cvbn(kv)                                             # Keras names/layouts -> CVBN (torch layouts)
    assert set(cv) == {name for name, _ in weights.schema()}
    blob = weights.to_blob(cv)
    assert len(blob) == 16 + 4 * weights.n_params()
    frames = synth.synth_streams(n, 5, h, w, config_idx=seed)
    stacked = np.concatenate([mpr.tensorise_stream(frames[s], 4, 1) for s in range(n)])
    x = mpr.stacked_to_nchw(stacked, 4)
    ref, ref_i = blobnet_ref.blobnet_forward(blobnet_ref.parse_blob(blob), x, return_intermediates=True)
    got, got_i = blobnet_np.blobnet_forward_np({k.split("/", k.count("/") - 1)[-1].split(":")[0]: v for k, v in kv.items()}, x, return_intermediates=True)
    for k in ref_i:
        a, b = got_i[k], ref_i[k]
        assert a.shape == b.shape, k
        assert np.abs(a - b).max() <= 2e-6 * max(1.0, np.abs(b).max()), (k, np.abs(a - b).max())
    assert got.shape == ref.shape == (2 * n, h, w)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= 3e-6, err
    assert ((got > 0) != (ref > 0)).mean() <= 1e-4                          # only exact-zero-crossing round-off may differ


def test_converter_rejects_wrong_shapes_and_missing_layers():
    kv = keras_variables(0)
    bad = dict(kv)
    k = next(n for n in bad if "conv3d_transpose_2/kernel" in n)
    bad[k] = bad[k].transpose(0, 1, 2, 4, 3)                               # (Cin, Cout) swapped: torch order in a Keras file
    with pytest.raises(ValueError):
        k2c.keras_to_cvbn(bad)
    missing = {n: v for n, v in kv.items() if "conv1d_7" not in n}
    with pytest.raises(ValueError):
        k2c.keras_to_cvbn(missing)
    with pytest.raises(ValueError):
        k2c.keras_to_cvbn({**kv, "model/dense/kernel:0": np.zeros((4, 4), np.float32)})


def test_converter_command_line_roundtrip(tmp_path):
    kv = keras_variables(5)
    np.savez(tmp_path / "vars.npz", **kv)
    k2c.main(["keras_to_cvbn.py", str(tmp_path / "vars.npz"), str(tmp_path / "w.cvbn")])
    w = weights.from_blob(open(tmp_path / "w.cvbn", "rb").read())
    want = k2c.keras_to_cvbn(kv)
    assert all((w[k] == want[k]).all() for k in want)
