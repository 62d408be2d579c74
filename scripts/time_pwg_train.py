import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parakeet_b200.models import PWGDiscriminator, PWGGenerator
from parakeet_b200.training import PWGTrainStep
from parakeet_b200 import _lib
dev = "cuda"
for graphs in (False, True):
    gen = PWGGenerator(layers=30, stacks=3, upsample_scales=[4, 5, 3, 5], device=dev, seed=5)
    dis = PWGDiscriminator(device=dev, seed=6)
    ts = PWGTrainStep(gen, dis, discriminator_train_start_steps=0, use_graphs=graphs)
    ts.iteration = 1
    g = torch.Generator().manual_seed(9)
    wav = (torch.randn(6, 1, 85 * 300, generator=g) * 0.3).to(dev)
    mel = torch.randn(6, 80, 89, generator=g).to(dev)
    for i in range(7):
        n0 = _lib.launch_count()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = ts.update_core((wav, mel))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"graphs={graphs} step {i}: {1e3 * (t1 - t0):8.1f} ms  launches {_lib.launch_count() - n0}  replays {ts._graphs.replays}  mem {torch.cuda.memory_allocated() / 1e9:.1f} GB", flush=True)
