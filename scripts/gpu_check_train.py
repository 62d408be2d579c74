import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import fastspeech2 as ofs
from parakeet_b200.models import FastSpeech2
from parakeet_b200.training import FastSpeech2TrainStep
params = ofs.synth_params(1)
batch = ofs.synth_train_batch(5, [9, 14, 11], dur_range=(1, 4))
losses_ref, grads_ref, stats_ref = ofs.train_step_grads(params, None, batch, stop_gradient_from_pitch_predictor=True)
m = FastSpeech2(80, 80, **ofs.LJSPEECH_MODEL_CFG, stop_gradient_from_pitch_predictor=True, device="cuda")
m.set_state_dict(params)
ts = FastSpeech2TrainStep(m, learning_rate=1e-3, dropout=False)
losses = ts.forward_backward(batch)
print("losses", [float(v) for v in losses], losses_ref)
rows = []
for k, gref in grads_ref.items():
    g = ts.grads[k].detach().double().cpu(); r = gref.double()
    mx = (g - r).abs().max().item() / max(r.abs().max().item(), 1e-12)
    l2 = (g - r).norm().item() / max(r.norm().item(), 1e-12)
    rows.append((mx, l2, k, r.abs().max().item()))
rows.sort(reverse=True)
print("worst by max-rel:")
for r in rows[:25]: print(f"  max-rel {r[0]:.3e}  l2-rel {r[1]:.3e}  |ref|max {r[3]:.3e}  {r[2]}")
print("count max-rel>2e-3:", sum(r[0] > 2e-3 for r in rows), "of", len(rows), "; l2-rel>2e-3:", sum(r[1] > 2e-3 for r in rows))
