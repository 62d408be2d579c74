"""Oracle: STFT / mel front-end (numpy + torch-CPU restatement).  TEST INFRASTRUCTURE ONLY.

Follows, in the reference (/root/reference/parakeet):
  modules/audio.py      STFT.__init__ :111-159 (window = scipy get_window(fftbins=True), pad_center to n_fft; weight rows =
                        Re/Im(fft(eye(n_fft)))[:bins] * window), forward :161-185 (reflect pad n_fft//2, conv1d stride hop),
                        power :187-201, magnitude :203-215, MelScale :218-229
  modules/stft_loss.py  stft :20-67 (paddle.signal.stft -> sqrt(clip(re^2+im^2, 1e-7)), transposed), SpectralConvergenceLoss
                        :70-91, LogSTFTMagnitudeLoss :94-118, STFTLoss :121-160, MultiResolutionSTFTLoss :163-219
  data/get_feats.py     LogMelFBank :20-88, Energy._calculate_energy :196-203
librosa.filters.mel (absent here) is restated below (Slaney mel scale, Slaney area normalisation = librosa defaults) and
cross-checked against torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney") in the tests.
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy import signal


def make_window(window, win_length, n_fft):
    w = signal.get_window(window, win_length, fftbins=True)
    if n_fft != win_length:                       # librosa.util.pad_center(mode="constant")
        lpad = (n_fft - win_length) // 2
        w = np.pad(w, (lpad, n_fft - win_length - lpad))
    return w


def stft_dft_conv(x, n_fft, hop_length=None, win_length=None, window="hann", center=True):
    """STFT.forward (modules/audio.py:161-185): the O(N^2) DFT-matrix conv1d.  x (B, T) -> real, imag (B, bins, frames)."""
    win_length = win_length or n_fft
    hop_length = hop_length or win_length // 4
    n_bin = 1 + n_fft // 2
    w = make_window(window, win_length, n_fft)
    weight = np.fft.fft(np.eye(n_fft))[:n_bin]
    wk = np.concatenate([weight.real, weight.imag], axis=0) * w
    wk = torch.tensor(wk[:, None, :], dtype=torch.float32)
    x = x.unsqueeze(1)
    if center:
        x = F.pad(x, [n_fft // 2, n_fft // 2], mode="reflect")
    out = F.conv1d(x, wk, stride=hop_length)
    real, imag = torch.chunk(out, 2, dim=1)
    return real, imag


def stft_magnitude(x, fft_size, hop_length, win_length, window="hann", eps=1e-7):
    """stft() of modules/stft_loss.py:20-67: (B, T) -> (B, frames, bins) magnitude with the 1e-7 power clip."""
    w = torch.tensor(signal.get_window(window, win_length, fftbins=True), dtype=x.dtype)
    X = torch.stft(x, fft_size, hop_length, win_length, window=w, center=True, pad_mode="reflect", return_complex=True)
    return torch.sqrt(torch.clip(X.real ** 2 + X.imag ** 2, min=eps)).transpose(1, 2)


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels=80, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with htk=False, norm='slaney' -> (n_mels, 1 + n_fft//2) float32."""
    fmax = sr / 2 if fmax is None else fmax
    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (weights * enorm[:, None]).astype(np.float32)


def log_mel_fbank(wav, sr=24000, n_fft=2048, hop_length=300, win_length=None, window="hann", n_mels=80, fmin=80, fmax=7600):
    """LogMelFBank.get_log_mel_fbank (data/get_feats.py:47-88): log10(clip(mel_filter @ |stft|, 1e-10)).T -> (frames, n_mels)."""
    win_length = win_length or n_fft
    x = torch.as_tensor(wav, dtype=torch.float32)[None]
    w = torch.tensor(signal.get_window(window, win_length, fftbins=True), dtype=torch.float32)
    X = torch.stft(x, n_fft, hop_length, win_length, window=w, center=True, pad_mode="reflect", return_complex=True)[0]
    S = X.abs().numpy()
    mel = np.dot(mel_filterbank(sr, n_fft, n_mels, fmin, fmax), S)
    return np.log10(np.clip(mel, 1e-10, None)).T


def energy(wav, n_fft=2048, hop_length=300, win_length=None, window="hann"):
    """Energy._calculate_energy (data/get_feats.py:196-203): sqrt(clip(sum_k |X|^2, 1e-10)) -> (frames,)."""
    win_length = win_length or n_fft
    x = torch.as_tensor(wav, dtype=torch.float32)[None]
    w = torch.tensor(signal.get_window(window, win_length, fftbins=True), dtype=torch.float32)
    X = torch.stft(x, n_fft, hop_length, win_length, window=w, center=True, pad_mode="reflect", return_complex=True)[0]
    return np.sqrt(np.clip((X.abs().numpy() ** 2).sum(0), 1e-10, None))


def multi_resolution_stft_loss(x, y, fft_sizes=(1024, 2048, 512), hop_sizes=(120, 240, 50), win_lengths=(600, 1200, 240),
                               window="hann"):
    """MultiResolutionSTFTLoss.forward (stft_loss.py:190-219): (sc_loss, mag_loss), each the mean over resolutions."""
    sc, mag = 0.0, 0.0
    for fs, ss, wl in zip(fft_sizes, hop_sizes, win_lengths):
        xm, ym = stft_magnitude(x, fs, ss, wl, window), stft_magnitude(y, fs, ss, wl, window)
        sc = sc + torch.norm(ym - xm, p="fro") / torch.clip(torch.norm(ym, p="fro"), min=1e-10)
        mag = mag + F.l1_loss(torch.log(torch.clip(xm, min=1e-7)), torch.log(torch.clip(ym, min=1e-7)))
    return sc / len(fft_sizes), mag / len(fft_sizes)
