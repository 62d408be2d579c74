"""AudioProcessor (reference parakeet/audio/audio.py:21-102) with the transforms on the GPU.

Same constructor and methods as the reference (`read_wav`, `write_wav`, `stft`, `istft`, `spectrogram`, `mel_spectrogram`,
attributes `mel_filter` / `inv_mel_filter`).  The reference delegates to librosa / soundfile, which are not part of this image:
* stft / spectrogram / mel_spectrogram run in pk_stft (csrc/stft.cu: windowed radix-2 FFT + fused magnitude / mel epilogue)
  and return what librosa returns: complex64 (bins, frames), float32 (bins, frames), float32 (n_mels, frames);
* `mel_filter` is the Slaney filterbank librosa.filters.mel builds by default (modules/audio.py: mel_filterbank);
* wav IO is scipy.io.wavfile (+ polyphase resampling when the file's rate differs), volume normalisation as the reference;
* `istft` (used only by Griffin-Lim style callers, not on the hot path) is torch.istft with the same window conventions.
"""
import numpy as np
import torch

from ..modules.audio import STFT, mel_filterbank

__all__ = ["AudioProcessor"]


class AudioProcessor(object):
    def __init__(self, sample_rate: int, n_fft: int, win_length: int, hop_length: int, n_mels: int = 80, fmin: int = 0,
                 fmax: int = None, window="hann", center=True, pad_mode="reflect", normalize=True, device=None):
        self.sample_rate, self.normalize = sample_rate, normalize
        self.n_fft, self.win_length, self.hop_length = n_fft, win_length, hop_length
        self.window, self.center, self.pad_mode = window, center, pad_mode
        self.n_mels, self.fmin, self.fmax = n_mels, fmin, fmax
        self._stft = STFT(n_fft, hop_length, win_length, window, center=center, pad_mode=pad_mode, device=device)
        self.mel_filter = self._create_mel_filter()
        self.inv_mel_filter = np.linalg.pinv(self.mel_filter)
        self._mel_w = torch.from_numpy(self.mel_filter).to(self._stft.device)

    def _create_mel_filter(self):
        return mel_filterbank(self.sample_rate, self.n_fft, n_mels=self.n_mels, fmin=self.fmin or 0.0, fmax=self.fmax)

    # -- IO (host) ---------------------------------------------------------------------------------------------------
    def read_wav(self, filename):
        from scipy.io import wavfile
        sr, wav = wavfile.read(filename)
        if wav.dtype.kind == "i":
            wav = wav.astype(np.float32) / float(np.iinfo(wav.dtype).max + 1)
        elif wav.dtype.kind == "u":                                   # 8-bit PCM is unsigned
            wav = (wav.astype(np.float32) - 128.0) / 128.0
        wav = wav.astype(np.float32)
        if wav.ndim == 2:                                             # librosa.load(mono=True)
            wav = wav.mean(axis=1)
        if sr != self.sample_rate:                                    # "resampling may occur" (audio.py:62-63)
            from math import gcd
            from scipy.signal import resample_poly
            g = gcd(int(sr), int(self.sample_rate))
            wav = resample_poly(wav, self.sample_rate // g, sr // g).astype(np.float32)
        if self.normalize:
            wav = wav / np.max(np.abs(wav)) * 0.999
        return wav

    def write_wav(self, path, wav):
        from scipy.io import wavfile
        wavfile.write(path, self.sample_rate, np.asarray(wav, dtype=np.float32))

    # -- transforms (GPU) ----------------------------------------------------------------------------------------------
    def _dev(self, wav):
        return torch.as_tensor(np.asarray(wav), dtype=torch.float32, device=self._stft.device).reshape(1, -1)

    def stft(self, wav):
        re, im = self._stft(self._dev(wav))
        return torch.complex(re[0], im[0]).cpu().numpy()

    def istft(self, D):
        D = torch.as_tensor(np.asarray(D), device=self._stft.device)
        win = self._stft._win if self.win_length == self.n_fft else None
        if win is None:
            from scipy import signal
            win = torch.from_numpy(signal.get_window(self.window, self.win_length, fftbins=True).astype(np.float32)).to(D.device)
        return torch.istft(D, self.n_fft, self.hop_length, self.win_length, window=win, center=self.center).cpu().numpy()

    def spectrogram(self, wav):
        return self._stft.magnitude(self._dev(wav))[0].cpu().numpy()

    def mel_spectrogram(self, wav):
        o = self._stft._run(self._dev(wav), mel_w=self._mel_w, mel_log10=False, mel_clip=0.0)
        return o["mel"][0].transpose(0, 1).contiguous().cpu().numpy()
