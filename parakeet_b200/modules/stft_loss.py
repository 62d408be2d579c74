"""Multi-resolution STFT loss, forward value (reference parakeet/modules/stft_loss.py:20-219)."""
import torch

from .. import _lib
from ..layer import Layer
from ..ops import _ptr, _stream
from .audio import STFT


def stft(x, fft_size, hop_length=None, win_length=None, window="hann", center=True, pad_mode="reflect", _cache={}):
    """(B, T) -> (B, frames, bins) magnitude with the reference's 1e-7 power clip (stft_loss.py:20-67)."""
    key = (fft_size, hop_length, win_length, window, center, str(x.device))
    if key not in _cache:
        _cache[key] = STFT(fft_size, hop_length, win_length, window, center=center, pad_mode=pad_mode, device=x.device)
    return _cache[key]._run(x, mag=True, mag_layout=1, power_clip=1e-7)["mag"]


class STFTLoss(Layer):
    def __init__(self, fft_size=1024, shift_size=120, win_length=600, window="hann", device=None):
        super().__init__(device)
        self.fft_size, self.shift_size, self.win_length, self.window = fft_size, shift_size, win_length, window

    def forward(self, x, y):
        """-> (spectral convergence loss, log STFT magnitude loss) as 0-d CUDA tensors (stft_loss.py:70-160)."""
        xm = stft(x, self.fft_size, self.shift_size, self.win_length, self.window)
        ym = stft(y, self.fft_size, self.shift_size, self.win_length, self.window)
        sums = torch.empty(3, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().pk_spectral_loss_sums(_ptr(xm), _ptr(ym), xm.numel(), 1e-7, _ptr(sums), _stream()),
                   "pk_spectral_loss_sums")
        sc = torch.sqrt(sums[0]) / torch.clamp(torch.sqrt(sums[1]), min=1e-10)
        return sc, sums[2] / xm.numel()


class MultiResolutionSTFTLoss(Layer):
    def __init__(self, fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50], win_lengths=[600, 1200, 240], window="hann",
                 device=None):
        super().__init__(device)
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        self.stft_losses = [STFTLoss(fs, ss, wl, window, device=device) for fs, ss, wl in zip(fft_sizes, hop_sizes, win_lengths)]

    def forward(self, x, y):
        sc_loss, mag_loss = 0.0, 0.0
        for f in self.stft_losses:
            sc, mag = f(x, y)
            sc_loss, mag_loss = sc_loss + sc, mag_loss + mag
        return sc_loss / len(self.stft_losses), mag_loss / len(self.stft_losses)
