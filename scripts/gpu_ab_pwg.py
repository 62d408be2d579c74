"""A/B timing of PWG generator variants at the bench configuration (each variant in its own process: the switches are read once)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os
sys.path.insert(0, %r)
import torch
from oracle import pwg as opwg
from parakeet_b200.models import PWGGenerator
gen = PWGGenerator(**opwg.DEFAULT_GENERATOR_PARAMS, device="cuda")
params = opwg.synth_params(2, weight_norm=True)
gen.set_state_dict(params)
xs, cs = opwg.synth_inputs(2, batch=2, mel_frames=40)
with torch.no_grad():
    ref = opwg.generator_forward(opwg.fold_weight_norm(params), xs, cs)
y = gen(xs.cuda(), cs.cuda())
err = ((y.double().cpu() - ref.double()).abs().max() / ref.double().abs().max()).item()
x, c = opwg.synth_inputs(2, batch=32, mel_frames=400)
x, c = x.cuda(), c.cuda()
for _ in range(4): gen(x, c)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): gen(x, c)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 8)
print("%%-40s err %%.2e  %%.2f ms per batch" %% (os.environ.get("TAG"), err, best), flush=True)
''' % ROOT
VARIANTS = (("resid=mma", {"PK_PWG_RESID": "mma"}), ("resid=gate", {"PK_PWG_RESID": "gate"}), ("resid=ldg", {"PK_PWG_RESID": "ldg"}),
            ("resid=mma again", {"PK_PWG_RESID": "mma"}), ("resid=gate again", {"PK_PWG_RESID": "gate"}))
if "--all" in sys.argv:
    VARIANTS += (("round-1 kernel", {"PK_PWG_FRAME_COND": "0"}),)
for tag, env in VARIANTS:
    e = dict(os.environ, TAG=tag, **env)
    subprocess.run([sys.executable, "-c", CODE], env=e)
