"""GPU diagnostic for the EXPERIMENTAL frame-rate conditioning path (csrc/pwg_fc.cu, PK_PWG_FRAME_COND=1):
generator output vs the default path and vs the oracle, plus timings.  Not part of the test suite until it passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pwg as opwg
from parakeet_b200.models import PWGGenerator

dev = "cuda"
def err(a, b):
    return ((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()

cfg = dict(opwg.DEFAULT_GENERATOR_PARAMS)
params = opwg.synth_params(2, weight_norm=True)
gen = PWGGenerator(**cfg, device=dev)
gen.set_state_dict(params)
folded = opwg.fold_weight_norm(params)
for frames, batch in ((40, 2), (7, 3), (1, 1)):
    x, c = opwg.synth_inputs(2, batch=batch, mel_frames=frames)
    with torch.no_grad():
        ref = opwg.generator_forward(folded, x, c)
    os.environ["PK_PWG_FRAME_COND"] = "0"
    y0 = gen(x.to(dev), c.to(dev)).clone()
    os.environ["PK_PWG_FRAME_COND"] = "1"
    y1 = gen(x.to(dev), c.to(dev)).clone()
    torch.cuda.synchronize()
    print(f"frames {frames} batch {batch}: default err {err(y0, ref):.2e}  frame-cond err {err(y1, ref):.2e}  fc vs default {err(y1, y0):.2e}", flush=True)
# ragged batch
x, c = opwg.synth_inputs(3, batch=3, mel_frames=40)
lens = torch.tensor([40 * 300, 25 * 300, 33 * 300], dtype=torch.int32, device=dev)
os.environ["PK_PWG_FRAME_COND"] = "0"; y0 = gen(x.to(dev), c.to(dev), lens=lens).clone()
os.environ["PK_PWG_FRAME_COND"] = "1"; y1 = gen(x.to(dev), c.to(dev), lens=lens).clone()
print("ragged: fc vs default", err(y1, y0), flush=True)
# timing at the bench configuration
x, c = opwg.synth_inputs(2, batch=32, mel_frames=400)
x, c = x.to(dev), c.to(dev)
for flag in ("0", "1"):
    os.environ["PK_PWG_FRAME_COND"] = flag
    for _ in range(3):
        gen(x, c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gen(x, c)
    e1.record(); torch.cuda.synchronize()
    print(f"PK_PWG_FRAME_COND={flag}: {e0.elapsed_time(e1) / 5:.2f} ms per batch", flush=True)
