"""ZScore (reference parakeet/modules/normalizer.py:18-33): feature-last z-score with buffers `mu`, `sigma`."""
import numpy as np
import torch

from .. import ops
from ..layer import Layer


class ZScore(Layer):
    def __init__(self, mu, sigma, device=None):
        super().__init__(device)
        to_t = lambda v: torch.from_numpy(np.asarray(v, dtype=np.float32)) if not torch.is_tensor(v) else v  # noqa: E731
        self._register("mu", to_t(mu).reshape(-1))
        self._register("sigma", to_t(sigma).reshape(-1))

    @property
    def mu(self):
        return self._params["mu"]

    @property
    def sigma(self):
        return self._params["sigma"]

    def forward(self, x):
        return ops.zscore(x, self.mu, self.sigma, inverse=False)

    def inverse(self, x):
        return ops.zscore(x, self.mu, self.sigma, inverse=True)
