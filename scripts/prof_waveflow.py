"""ConditionalWaveFlow.infer (cfg4 shapes) for a launch-list profile (eager launches; run under
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PK_CUDA_GRAPHS", "0")
import torch
from parakeet_b200.models import ConditionalWaveFlow
dev = "cuda"
B, FRAMES = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 400
wf = ConditionalWaveFlow([16, 16], 8, 8, 16, 64, 80, (3, 3), device=dev, seed=4)
g = torch.Generator().manual_seed(4)
mel = (torch.randn(B, 80, FRAMES, generator=g) * 0.5 - 3).to(dev)
z = torch.randn(B, 256 * FRAMES - 272, generator=g).to(dev)
wf.infer(mel, z=z); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); wf.infer(mel, z=z); e1.record(); torch.cuda.synchronize()
print("eager ms/call", e0.elapsed_time(e1), flush=True)
torch.cuda.cudart().cudaProfilerStart()
wf.infer(mel, z=z)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
