"""Spectrogram magnitude normalisers (reference parakeet/audio/spec_normalizer.py:31-74): invertible element-wise maps used by
the WaveFlow / Tacotron2 data pipelines.  numpy in -> numpy out as in the reference; a torch tensor (CPU or CUDA) stays a
tensor on its device, so the maps can sit behind the GPU mel extraction without a host round trip."""
import math

import numpy as np
import torch

__all__ = ["NormalizerBase", "LogMagnitude", "UnitMagnitude"]


class NormalizerBase(object):
    def transform(self, spec):
        raise NotImplementedError("transform must be implemented")

    def inverse(self, normalized):
        raise NotImplementedError("inverse must be implemented")


class LogMagnitude(NormalizerBase):
    """log(max(x, min)) / exp (spec_normalizer.py:39-53)."""

    def __init__(self, min=1e-5):
        self.min = min

    def transform(self, x):
        if torch.is_tensor(x):
            return torch.log(torch.clamp(x, min=self.min))
        return np.log(np.maximum(x, self.min))

    def inverse(self, x):
        return torch.exp(x) if torch.is_tensor(x) else np.exp(x)


class UnitMagnitude(NormalizerBase):
    """dB scale mapped to [0, 1] (spec_normalizer.py:56-74): clip((20 log10(max(min, x)) - 20 + 100) / 100, 0, 1)."""

    def __init__(self, min=1e-5):
        self.min = min

    def transform(self, x):
        if torch.is_tensor(x):
            return torch.clamp((20 * torch.log10(torch.clamp(x, min=self.min)) - 20 + 100) / 100, 0, 1)
        return np.clip((20 * np.log10(np.maximum(self.min, x)) - 20 + 100) / 100, 0, 1)

    def inverse(self, x):
        if torch.is_tensor(x):
            return torch.exp((torch.clamp(x, 0, 1) * 100 - 100 + 20) / 20 * math.log(10))
        return np.exp((np.clip(x, 0, 1) * 100 - 100 + 20) / 20 * np.log(10))
