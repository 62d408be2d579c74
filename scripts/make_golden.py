"""Generate tests/golden/*.npz from the CPU oracle (seeded, small).  Re-run only when the oracle changes on purpose;
the CPU tests check the oracle against these files so that accidental drift is caught, the GPU tests check the CUDA
path against them without needing the oracle's runtime.

The reference itself cannot run here (PaddlePaddle absent), so these vectors pin the ORACLE, not the
reference: parity is "unpinned" in the sense of the task statement.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fastspeech2 as ofs  # noqa: E402
from oracle import pwg as opwg  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
torch.set_num_threads(1)

# --- Parallel WaveGAN: B=2, 6 mel frames -> 1800 samples, baker generator params, seed 2 (cfg2's weights) ---
params = opwg.fold_weight_norm(opwg.synth_params(2, weight_norm=True))
x, c = opwg.synth_inputs(2, batch=2, mel_frames=6)
with torch.no_grad():
    y, inter = opwg.generator_forward(params, x, c, return_intermediates=True)
np.savez_compressed(os.path.join(out, "pwg_small.npz"), x=x.numpy(), c=c.numpy(), y=y.numpy(),
                    c_up_checksum=np.float64(inter["c_up"].double().sum().item()),
                    skips_checksum=np.float64(inter["skips"].double().sum().item()))

# --- FastSpeech2 (LJSpeech yaml, V=80), seed 1 (cfg1's weights): T=24 single-utterance inference ---
fp = ofs.synth_params(1)
xs, il = ofs.synth_text(11, [24])
with torch.no_grad():
    b, a, d, p, e = ofs.fs2_forward(fp, None, xs, il, is_inference=True)
np.savez_compressed(os.path.join(out, "fs2_infer_small.npz"), text=xs.numpy(), after=a.numpy(), before=b.numpy(),
                    durations=d.numpy(), pitch=p.numpy(), energy=e.numpy())
# --- FastSpeech2 teacher-forced padded batch (cfg5-shaped, tiny) ---
batch = ofs.synth_train_batch(5, [9, 14, 11], dur_range=(1, 4))
with torch.no_grad():
    ref = ofs.fs2_forward(fp, None, batch["text"], batch["text_lengths"], batch["speech_lengths"], batch["durations"],
                          batch["pitch"], batch["energy"])
    losses = ofs.fs2_loss(ref[1], ref[0], ref[2], ref[3], ref[4], batch["speech"], batch["durations"], batch["pitch"],
                          batch["energy"], batch["text_lengths"], batch["speech_lengths"])
np.savez_compressed(os.path.join(out, "fs2_forward_small.npz"), **{k: v.numpy() for k, v in batch.items()},
                    before=ref[0].numpy(), after=ref[1].numpy(), d_outs=ref[2].numpy(), p_outs=ref[3].numpy(),
                    e_outs=ref[4].numpy(), losses=np.array([float(v) for v in losses]))
# --- length regulator: the reference's own test case (tests/unit/test_expansion.py:20-24) + a ragged one ---
enc = torch.arange(2 * 4 * 3, dtype=torch.float32).reshape(2, 4, 3) + 1
dur = torch.tensor([[1, 2, 2, 1], [3, 1, 4, 0]], dtype=torch.int64)
np.savez_compressed(os.path.join(out, "length_regulator.npz"), enc=enc.numpy(), dur=dur.numpy(),
                    out=ofs.length_regulator_expand(enc, dur).numpy())
for f in sorted(os.listdir(out)):
    print(f, os.path.getsize(os.path.join(out, f)))
