"""Minimal stand-in for paddle.nn.Layer: named parameters with the reference's state-dict key names.

The reference's checkpoint boundary is `layer.state_dict()` / `layer.set_state_dict()` keyed by Paddle parameter
names; this class keeps those names and Paddle's layouts (Linear [in, out], Conv [out, in, k]).
Parameters are torch tensors (fp32, on the layer's device); packed / split-bf16 copies for the kernels are derived
lazily and invalidated whenever a parameter changes.
"""
from collections import OrderedDict

import numpy as np
import torch

from .graph import GraphRunner


class Layer:
    def __init__(self, device=None):
        self._params = OrderedDict()
        self._graphs = GraphRunner() # CUDA graphs of the inference paths; they bake the packed weights' addresses
        self._packed = None          # derived kernel-ready weights (built on first forward)
        self.training = True
        self.device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))

    @property
    def _packed(self):
        return self.__dict__.get("_packed_value")

    @_packed.setter
    def _packed(self, value):
        self.__dict__["_packed_value"] = value
        if value is None and "_graphs" in self.__dict__:
            self._graphs.clear()     # captured graphs point at the old packed weights

    # -- parameter registry ----------------------------------------------------------------------------------
    def _register(self, name, tensor):
        self._params[name] = tensor.detach().to(device=self.device, dtype=torch.float32).contiguous()

    def state_dict(self):
        return OrderedDict((k, v) for k, v in self._params.items())

    def set_state_dict(self, state):
        missing = [k for k in self._params if k not in state]
        if missing:
            raise KeyError(f"missing keys in state dict: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        for k in self._params:
            v = state[k]
            if isinstance(v, np.ndarray):
                v = torch.from_numpy(v)
            v = torch.as_tensor(v)
            if tuple(v.shape) != tuple(self._params[k].shape):
                raise ValueError(f"shape mismatch for {k}: got {tuple(v.shape)}, expected {tuple(self._params[k].shape)}")
            cur = self._params[k]
            # in place: a training step may have turned the parameters into views of its flat buffer (training/flat.py);
            # re-registering fresh tensors would silently detach the model from the optimiser's storage
            cur.copy_(v.detach().to(device=cur.device, dtype=cur.dtype))
        self._packed = None

    load_dict = set_state_dict  # paddle alias

    def parameters(self):
        return list(self._params.values())

    def named_parameters(self):
        return list(self._params.items())

    def eval(self):
        self.training = False
        return self

    def train(self):
        self.training = True
        return self

    def to(self, device):
        self.device = torch.device(device)
        for k, v in list(self._params.items()):
            self._params[k] = v.to(self.device)
        self._packed = None
        return self

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)
