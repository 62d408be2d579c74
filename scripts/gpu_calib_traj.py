import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import fastspeech2 as ofs
from parakeet_b200.models import FastSpeech2
from parakeet_b200.training import FastSpeech2TrainStep
g = torch.Generator().manual_seed(53)
lengths = torch.randint(60, 141, (4,), generator=g).tolist()
params = ofs.synth_params(1)
batch = ofs.synth_train_batch(52, lengths)
lr = float(os.environ.get("LR", 2e-5))
p_ref, state, loss_ref = dict(params), {}, []
for _ in range(3):
    losses, grads, stats = ofs.train_step_grads(p_ref, None, batch, stop_gradient_from_pitch_predictor=True)
    loss_ref.append(losses["loss"])
    new = ofs.adam_step({k: p_ref[k] for k in grads}, grads, state, lr=lr)
    p_ref = {**p_ref, **new, **stats}
m = FastSpeech2(80, 80, **ofs.LJSPEECH_MODEL_CFG, stop_gradient_from_pitch_predictor=True, device="cuda")
m.set_state_dict(params)
ts = FastSpeech2TrainStep(m, learning_rate=lr, dropout=False)
loss_got = [float(ts.step(batch).sum()) for _ in range(3)]
print("loss", loss_got, loss_ref)
sd = m.state_dict()
rows = []
for k, v in p_ref.items():
    got, ref, init = sd[k].detach().double().cpu(), v.double(), params[k].double()
    moved = (ref - init).abs().max().item()
    e = (got - ref).norm().item() / max((ref - init).norm().item(), 1e-12)
    rows.append((e, moved, (got - init).abs().max().item(), k))
rows.sort(reverse=True)
for r in rows[:40]:
    print("%.3e moved %.3e gotmoved %.3e %s" % r)
print("n", len(rows), "e>5e-2:", sum(r[0] > 5e-2 for r in rows), "e>1e-2:", sum(r[0] > 1e-2 for r in rows))
