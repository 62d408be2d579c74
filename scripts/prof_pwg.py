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
gen._prof = torch.zeros(64, dtype=torch.int64, device="cuda")
y = gen(x.cuda(), c.cuda()); torch.cuda.synchronize()
pr = gen._prof.cpu().tolist(); nt = max(pr[32], 1)
names = {0: "prod wait_empty", 1: "prod issue", 8: "mma wait acc1_empty", 9: "mma wait full(G1)", 10: "mma issue", 11: "mma wait z_full",
         12: "mma wait acc2_empty", 13: "mma wait full(G2)", 14: "mma other",
         16: "gate0 wait acc1_full", 17: "gate0 ld+math", 18: "gate0 wait g2 stage", 19: "gate0 z store", 22: "gate0 loop",
         40: "skip wait acc2_full", 41: "skip rest", 42: "skip tmem ld", 43: "skip sts", 44: "skip lds", 45: "skip red/st", 46: "skip loop+prefetch",
         48: "out wait acc2_full", 49: "out rest", 50: "out tmem ld", 51: "out sts", 52: "out lds", 53: "out global", 54: "out loop+prefetch"}
if os.environ.get("PK_PWG_FRAME_COND", "0") == "1":     # pwg_fc.cu (4-stage ring, z in tensor memory): different buckets
    names.update({8: "mma wait full tap-d", 9: "mma wait full tap+d", 14: "mma wait full cond", 15: "mma wait full centre", 10: "mma issue G1 + resid", 11: "mma wait z_full", 12: "mma wait acc2_empty",
                  13: "mma issue G2", 17: "gate0 ld+math+st", 18: "gate0 (unused)", 19: "gate0 wait st + arrive"})
print("tiles (all CTAs, all layers):", nt)
for k, n in names.items():
    print(f"  {n:24s} {pr[k] / nt:9.0f} cycles/tile")
