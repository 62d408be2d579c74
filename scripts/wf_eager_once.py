"""One eager ConditionalWaveFlow.inverse at cfg4 shapes (for an ncu launch list; no graphs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parakeet_b200.models import ConditionalWaveFlow
dev = "cuda"
B, FRAMES = 16, 400
wf = ConditionalWaveFlow([16, 16], 8, 8, 16, 64, 80, (3, 3), device=dev, seed=4)
g = torch.Generator().manual_seed(4)
mel = (torch.randn(B, 80, FRAMES, generator=g) * 0.5 - 3).to(dev)
z = torch.randn(B, 256 * FRAMES - 272, generator=g).to(dev)
cond = wf.encode(mel)
torch.cuda.synchronize()
y = wf.inverse(z, cond)
torch.cuda.synchronize()
print("done", tuple(y.shape))
