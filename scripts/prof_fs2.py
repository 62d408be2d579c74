"""FastSpeech2 batch inference (cfg3 shapes) for a launch-list profile:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/prof_fs2.py
Without ncu it prints the CUDA-event time per call."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parakeet_b200.models import FastSpeech2
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
fs = FastSpeech2(80, 80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536, positionwise_layer_type="conv1d",
                 positionwise_conv_kernel_size=3, duration_predictor_layers=2, duration_predictor_chans=256,
                 duration_predictor_kernel_size=3, postnet_layers=5, postnet_filts=5, postnet_chans=256, pitch_predictor_layers=5,
                 pitch_predictor_chans=256, pitch_predictor_kernel_size=5, pitch_embed_kernel_size=1, energy_predictor_layers=2,
                 energy_predictor_chans=256, energy_predictor_kernel_size=3, energy_embed_kernel_size=1, device=dev, seed=1)
sd = dict(fs.state_dict()); sd["duration_predictor.linear.bias"] = torch.tensor([math.log(8.0)]); fs.set_state_dict(sd)
g = torch.Generator().manual_seed(3)
lengths = torch.randint(60, 141, (B,), generator=g).tolist()
ids = torch.zeros(B, max(lengths), dtype=torch.int64)
for i, n in enumerate(lengths):
    ids[i, :n] = torch.randint(1, 79, (n,), generator=g)
il = torch.tensor(lengths, dtype=torch.int64)
ids, il = ids.to(dev), il.to(dev)
for _ in range(3):
    mel, olens, _ = fs.batch_inference(ids, il)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    mel, olens, _ = fs.batch_inference(ids, il)
e1.record(); torch.cuda.synchronize()
print("frames", int(olens.sum()), "ms/call", e0.elapsed_time(e1) / 5, flush=True)
torch.cuda.cudart().cudaProfilerStart()
mel, olens, _ = fs.batch_inference(ids, il)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
