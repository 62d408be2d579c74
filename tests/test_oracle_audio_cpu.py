"""CPU cross-checks of the STFT / mel oracle (SURVEY.md 8c items 1 and the librosa-free mel filterbank)."""
import numpy as np
import torch

from oracle import stft as ost


def test_dft_conv_equals_torch_stft_equals_numpy_rfft():
    x = torch.randn(2, 3000)
    re, im = ost.stft_dft_conv(x, 512, 128, 400, "hann")
    w = torch.tensor(ost.signal.get_window("hann", 400, fftbins=True), dtype=torch.float32)
    X = torch.stft(x, 512, 128, 400, window=w, center=True, pad_mode="reflect", return_complex=True)
    assert (re - X.real).abs().max() < 1e-3 and (im - X.imag).abs().max() < 1e-3
    xp = np.pad(x[0].numpy(), 256, mode="reflect")
    win = ost.make_window("hann", 400, 512)
    for t in (0, 5, 23):
        ref = np.fft.rfft(xp[t * 128:t * 128 + 512] * win)
        assert np.abs(re[0, :, t].numpy() - ref.real).max() < 1e-3 and np.abs(im[0, :, t].numpy() - ref.imag).max() < 1e-3
    assert re.shape[-1] == 1 + 3000 // 128


def test_mel_filterbank_matches_torchaudio_slaney():
    import torchaudio
    fb = ost.mel_filterbank(24000, 2048, 80, 80, 7600)
    ta = torchaudio.functional.melscale_fbanks(1025, 80.0, 7600.0, 80, 24000, norm="slaney", mel_scale="slaney").T.numpy()
    assert np.abs(fb - ta).max() < 1e-6
    from parakeet_b200.modules.audio import mel_filterbank
    assert np.abs(mel_filterbank(22050, 1024, 80, 0, 8000) - ost.mel_filterbank(22050, 1024, 80, 0, 8000)).max() == 0


def test_mr_stft_frames():
    x = torch.randn(1, 25500)
    assert [ost.stft_magnitude(x, f, h, w).shape[1] for f, h, w in ((1024, 120, 600), (2048, 240, 1200), (512, 50, 240))] == [213, 107, 511]
