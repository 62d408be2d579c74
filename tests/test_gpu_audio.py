"""GPU parity tests of the STFT / mel front-end against the oracle (numpy / torch.stft / DFT-matrix conv)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def test_stft_three_way(cuda):
    """The reference's own comparison (tests/unit/test_stft.py:25-42): STFT(1024,256,1024) vs torch.stft vs the DFT conv."""
    from oracle import stft as ost
    from parakeet_b200.modules.audio import STFT
    x = torch.randn(4, 46080)
    re_ref, im_ref = ost.stft_dft_conv(x[:, :8192], 1024, 256, 1024, "hann")          # O(N^2) restatement, short signal
    st = STFT(1024, 256, 1024, "hanning", device=cuda)
    re, im = st(x[:, :8192].to(cuda))
    assert list(re.shape) == list(re_ref.shape) == [4, 513, 33]
    scale = max(re_ref.abs().max().item(), im_ref.abs().max().item())
    assert (re.cpu() - re_ref).abs().max().item() < 1e-4 * scale and (im.cpu() - im_ref).abs().max().item() < 1e-4 * scale
    X = torch.stft(x, 1024, 256, 1024, window=torch.hann_window(1024, periodic=True), center=True, pad_mode="reflect",
                   return_complex=True)
    mag = st.magnitude(x.to(cuda))
    assert rel_err(mag, X.abs()) < 1e-4
    assert rel_err(st.power(x.to(cuda)), X.abs() ** 2) < 1e-4


@pytest.mark.parametrize("n_fft,hop,win", [(1024, 120, 600), (2048, 240, 1200), (512, 50, 240)])
def test_stft_loss_magnitudes(cuda, n_fft, hop, win):
    """stft() of stft_loss.py at the three MR-STFT resolutions on 25 500-sample clips: frames 213 / 107 / 511."""
    from oracle import stft as ost
    from parakeet_b200.modules.stft_loss import stft
    x = torch.randn(3, 25500)
    ref = ost.stft_magnitude(x, n_fft, hop, win)
    out = stft(x.to(cuda), n_fft, hop, win)
    assert list(out.shape) == list(ref.shape) and out.shape[1] == 1 + 25500 // hop
    assert rel_err(out, ref) < 1e-4


def test_multi_resolution_stft_loss(cuda):
    from oracle import stft as ost
    from parakeet_b200.modules.stft_loss import MultiResolutionSTFTLoss
    x, y = torch.randn(2, 25500), torch.randn(2, 25500)
    sc_ref, mag_ref = ost.multi_resolution_stft_loss(x, y)
    sc, mag = MultiResolutionSTFTLoss(device=cuda)(x.to(cuda), y.to(cuda))
    assert abs(float(sc) - float(sc_ref)) < 1e-3 * float(sc_ref)
    assert abs(float(mag) - float(mag_ref)) < 1e-3 * float(mag_ref)


def test_log_mel_and_energy_features(cuda):
    """LogMelFBank / Energy of data/get_feats.py at the CSMSC settings (24 kHz, n_fft 2048, hop 300, win 1200)."""
    from oracle import stft as ost
    from parakeet_b200.modules.audio import Energy, LogMelFBank, MelScale, STFT
    wav = (np.random.default_rng(0).standard_normal(24000) * 0.1).astype(np.float32)
    ref = ost.log_mel_fbank(wav, 24000, 2048, 300, 1200, "hann", 80, 80, 7600)
    out = LogMelFBank(24000, 2048, 300, 1200, "hann", 80, 80, 7600, device=cuda).get_log_mel_fbank(wav)
    assert list(out.shape) == list(ref.shape) == [81, 80]
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-3          # log10 domain, values ~[-3, 1]
    e_ref = ost.energy(wav, 2048, 300, 1200)
    e = Energy(24000, 2048, 300, 1200, device=cuda).get_energy(wav)
    assert rel_err(e, torch.from_numpy(e_ref)) < 1e-4
    # MelScale (matmul form) on a magnitude spectrogram
    st = STFT(2048, 300, 1200, "hann", device=cuda)
    mag = st.magnitude(torch.from_numpy(wav)[None].to(cuda))
    mel = MelScale(24000, 2048, 80, 80, 7600, device=cuda)(mag)
    assert rel_err(mel, torch.from_numpy(ost.mel_filterbank(24000, 2048, 80, 80, 7600)) @ mag.cpu()[0]) < 1e-4


def test_stft_full_size_parseval(cuda):
    """Size-independent property at full feature-extraction size (32 x 5 s of 24 kHz audio): with a rectangular window
    and hop = n_fft, center=False, Parseval holds per frame: sum_k c_k |X_k|^2 = N * sum_n x_n^2."""
    from parakeet_b200.modules.audio import STFT
    x = torch.randn(32, 120000, device=cuda)
    N = 2048
    st = STFT(N, N, N, "boxcar", center=False, device=cuda)
    p = st.power(x)                                               # (B, bins, frames)
    w = torch.full((N // 2 + 1,), 2.0, device=cuda)
    w[0] = w[-1] = 1.0
    lhs = (p * w[None, :, None]).sum(1)
    frames = p.shape[-1]
    rhs = N * (x[:, :frames * N].reshape(32, frames, N) ** 2).sum(-1)
    assert rel_err(lhs, rhs) < 1e-4
