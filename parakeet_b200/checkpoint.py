"""Read the reference's on-disk formats without PaddlePaddle.

What the reference writes (none of it can be produced here - Paddle is not installable - so the layouts below are restated
from Paddle 2.1's `paddle.save` / `paddle.load` (python/paddle/framework/io.py) and recorded as decisions to re-verify):

* `snapshot_iter_*.pdz` (parakeet/training/extensions/snapshot.py:86-91 -> updater.save, e.g.
  fastspeech2_updater / standard_updater `state_dict()`): `paddle.save` of a nested dict
  `{"main_params": model.state_dict(), "main_optimizer": ..., "epoch": int, "iteration": int}` (PWG:
  `"generator_params"`, `"discriminator_params"`, ...).  For a nested object Paddle pickles with a reducer that turns every
  Tensor into the tuple `(tensor.name, ndarray)`.
* `step-N.pdparams` (parakeet/utils/checkpoint.py:125-136): `paddle.save(model.state_dict())`; a flat state dict is
  stored as `{structured_name: ndarray, ..., "StructuredToParameterName@@": {...}}`; arrays above 2^30 bytes are split
  into flat slices listed under `"UnpackBigParamInfor@@"`.
* `*_stats.npy` (utils/compute_statistics.py:106-107): `np.stack([mean, scale])`, consumed as `mu, std = np.load(path)`
  (examples/fastspeech2/synthesize.py:78-85).

`load()` returns plain nested dicts of numpy arrays, which `Layer.set_state_dict` accepts directly:

    model.set_state_dict(checkpoint.load(path)["main_params"])          # examples/fastspeech2/synthesize.py:68-69
"""
import pickle
from collections import OrderedDict

import numpy as np

_ALLOWED = {
    ("builtins", "tuple"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "set"), ("builtins", "frozenset"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "complex"), ("builtins", "bool"), ("builtins", "str"),
    ("builtins", "bytes"), ("builtins", "bytearray"), ("builtins", "slice"), ("__builtin__", "tuple"),
    ("collections", "OrderedDict"), ("copy_reg", "_reconstructor"), ("copyreg", "_reconstructor"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "scalar"), ("numpy.core.numeric", "_frombuffer"),
    ("numpy._core.numeric", "_frombuffer"),
    ("_codecs", "encode"),      # how protocol-2 pickles carry the raw bytes of an ndarray
}


class _RestrictedUnpickler(pickle.Unpickler):
    """A checkpoint is data: only containers and numpy arrays may be rebuilt (a pickle can otherwise run arbitrary code)."""

    def find_class(self, module, name):
        if (module, name) in _ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"checkpoint refers to {module}.{name}, which is not a container or numpy type")


def _unpack(obj):
    if isinstance(obj, dict):
        info = obj.get("UnpackBigParamInfor@@")
        if isinstance(info, dict):                                   # re-assemble arrays Paddle split into < 2^30-byte slices
            obj = dict(obj)
            for key, value in info.items():
                parts = [np.asarray(obj.pop(part)).reshape(-1) for part in value["slices"]]
                obj[key] = np.concatenate(parts).reshape(value["OriginShape"])
            obj.pop("UnpackBigParamInfor@@")
        out = OrderedDict()
        for k, v in obj.items():
            if k == "StructuredToParameterName@@":
                continue
            out[k] = _unpack(v)
        return out
    if isinstance(obj, tuple) and len(obj) == 2 and isinstance(obj[0], str) and isinstance(obj[1], np.ndarray):
        return obj[1]                                                # (tensor.name, ndarray) written by the Tensor reducer
    if isinstance(obj, (list, tuple)):
        return type(obj)(_unpack(v) for v in obj)
    return obj


def load(path):
    """`paddle.load(path)` for .pdz / .pdparams / .pdopt files: nested dicts with every tensor as a numpy array."""
    with open(path, "rb") as f:
        obj = _RestrictedUnpickler(f, encoding="latin1").load()
    return _unpack(obj)


def load_stats(path):
    """`mu, std = np.load(stats.npy)` (examples/fastspeech2/synthesize.py:78-85) -> two float32 vectors for ZScore."""
    stat = np.load(path, allow_pickle=False)
    if stat.ndim != 2 or stat.shape[0] != 2:
        raise ValueError(f"{path}: expected np.stack([mean, scale]) of shape (2, n_mels), got {stat.shape}")
    return stat[0].astype(np.float32), stat[1].astype(np.float32)


def save(obj, path):
    """`paddle.save(obj, path)` for nested dicts of arrays / tensors / scalars (snapshot_iter_*.pdz: updater.state_dict(),
    step-N.pdparams: model.state_dict()): every tensor is stored as a plain numpy array, pickle protocol 2 like Paddle 2.1.
    `load()` above (and `paddle.load`, which passes ndarrays through) reads it back."""
    import torch

    def conv(o):
        if torch.is_tensor(o):
            return o.detach().cpu().numpy()
        if isinstance(o, dict):
            return OrderedDict((k, conv(v)) for k, v in o.items())
        if isinstance(o, (list, tuple)):
            return type(o)(conv(v) for v in o)
        return o
    with open(path, "wb") as f:
        pickle.dump(conv(obj), f, protocol=2)
