"""The CUDA path against vectors computed by the REFERENCE'S OWN CODE (tests/golden/ref_executed_models.npz, written by
scripts/make_golden_ref.py: the reference's model classes executed on a torch-backed stand-in for the Paddle primitives they
call, with the same seeded state dicts these tests load).  Same structure as the oracle-golden tests of test_gpu_models.py."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3   # north_star: "within 1e-3 rel fp32" (max-abs error / max-abs reference per tensor)


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLD, "ref_executed_models.npz"))


def test_fastspeech2_cuda_vs_executed_reference(cuda, g):
    from oracle import fastspeech2 as ofs
    from parakeet_b200.models import FastSpeech2, FastSpeech2Loss
    m = FastSpeech2(80, 80, **ofs.LJSPEECH_MODEL_CFG, device=cuda)
    m.set_state_dict(ofs.synth_params(1))
    text = torch.from_numpy(g["fs2_inf_text"]).to(cuda)
    out = m.inference(text)
    assert tuple(out.shape) == g["fs2_inf_mel"].shape                      # integer durations: same number of frames
    assert rel_err(out, torch.from_numpy(g["fs2_inf_mel"])) < TOL
    out13 = m.inference(text, alpha=1.3)
    assert tuple(out13.shape) == g["fs2_inf_mel_alpha"].shape and rel_err(out13, torch.from_numpy(g["fs2_inf_mel_alpha"])) < TOL
    b = {k: torch.from_numpy(g[f"fs2_fwd_{k}"]).to(cuda) for k in ("text", "text_lengths", "speech", "speech_lengths", "durations", "pitch", "energy")}
    before, after, d_outs, p_outs, e_outs, ys, olens = m(b["text"], b["text_lengths"], b["speech"], b["speech_lengths"],
                                                          b["durations"], b["pitch"], b["energy"])
    for name, t in zip(("before", "after", "d_outs", "p_outs", "e_outs"), (before, after, d_outs, p_outs, e_outs)):
        assert rel_err(t, torch.from_numpy(g[f"fs2_fwd_out_{name}"])) < TOL, name
    losses = FastSpeech2Loss(device=cuda)(after, before, d_outs, p_outs, e_outs, ys, b["durations"], b["pitch"], b["energy"],
                                          b["text_lengths"], olens)
    got = [float(v) for v in losses]
    assert np.allclose(got[:4], g["fs2_loss"], rtol=1e-3), (got, g["fs2_loss"])


def test_pwg_cuda_vs_executed_reference(cuda, g):
    from oracle import pwg as opwg
    from parakeet_b200.models import PWGGenerator
    gen = PWGGenerator(**opwg.DEFAULT_GENERATOR_PARAMS, device=cuda)
    gen.set_state_dict(opwg.synth_params(2, weight_norm=True))
    y = gen(torch.from_numpy(g["pwg_x"]).to(cuda), torch.from_numpy(g["pwg_c"]).to(cuda))
    assert rel_err(y, torch.from_numpy(g["pwg_y"])) < TOL
    assert rel_err(y, torch.from_numpy(g["pwg_y_weight_norm"])) < TOL          # the reference's own weight_norm(g, v) path


def test_waveflow_cuda_vs_executed_reference(cuda, g):
    from oracle import waveflow as owf
    from parakeet_b200.models import ConditionalWaveFlow
    wf = ConditionalWaveFlow([16, 16], 8, 8, 16, 64, 80, (3, 3), device=cuda)
    wf.set_state_dict(owf.synth_params(4))
    mel, z = torch.from_numpy(g["wf_mel"]).to(cuda), torch.from_numpy(g["wf_z"]).to(cuda)
    assert rel_err(wf.encode(mel), torch.from_numpy(g["wf_cond"])) < 1e-4
    out = wf.infer(mel, z=z)
    assert tuple(out.shape) == g["wf_x"].shape and rel_err(out, torch.from_numpy(g["wf_x"])) < TOL
    # W = 335 columns: the +-128 taps of the widest layer are inside the row (the vector above has W = 127)
    mel2, z2 = torch.from_numpy(g["wf2_mel"]).to(cuda), torch.from_numpy(g["wf2_z"]).to(cuda)
    out2 = wf.infer(mel2, z=z2)
    assert tuple(out2.shape) == g["wf2_x"].shape and rel_err(out2, torch.from_numpy(g["wf2_x"])) < TOL
