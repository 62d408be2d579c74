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
# phase timing
gen._prof = torch.zeros(40, dtype=torch.int64, device="cuda")
y = gen(x.cuda(), c.cuda()); torch.cuda.synchronize()
pr = gen._prof.cpu().tolist(); nt = max(pr[32], 1)
names = {0: "prod wait_empty", 1: "prod issue", 8: "mma wait acc1_empty", 9: "mma wait full(G1)", 10: "mma issue", 11: "mma wait z_full",
         12: "mma wait acc2_empty", 13: "mma wait full(G2)", 14: "mma other",
         16: "epi0 wait acc1_full", 17: "epi0 ld+gate", 18: "epi0 wait z_empty", 19: "epi0 z store", 20: "epi0 wait acc2_full", 21: "epi0 E2", 22: "epi0 barrier+loop",
         24: "epi1 wait acc1_full", 25: "epi1 ld+gate", 26: "epi1 wait z_empty", 27: "epi1 z store", 28: "epi1 wait acc2_full", 29: "epi1 E2", 30: "epi1 barrier+loop"}
print("tiles (all CTAs, all layers):", nt)
for k, n in names.items():
    print(f"  {n:24s} {pr[k] / nt:9.0f} cycles/tile")
