import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pwg as opwg
from parakeet_b200.models import PWGGenerator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gen = PWGGenerator(**opwg.DEFAULT_GENERATOR_PARAMS, device="cuda")
gen.set_state_dict(opwg.synth_params(2, weight_norm=True))
x, c = opwg.synth_inputs(2, batch=B, mel_frames=400)
y = gen(x.cuda(), c.cuda())
torch.cuda.synchronize()
print(y.std().item())
