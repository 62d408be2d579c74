"""GPU diagnostic: FastSpeech2 (CUDA) vs the torch-CPU oracle."""
import sys, os, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import fastspeech2 as ofs
from parakeet_b200.models import FastSpeech2
from parakeet_b200 import ops

dev = "cuda"
def err(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()

torch.set_num_threads(os.cpu_count())
cfg = dict(ofs.LJSPEECH_MODEL_CFG)
params = ofs.synth_params(1)
model = FastSpeech2(80, 80, **cfg, device=dev)
model.set_state_dict(params)
bad = False

# 0. unit: layer norm / embed / softmax quick checks
x = torch.randn(3, 50, 384)
g, b = torch.randn(384), torch.randn(384)
y, ys = ops.layer_norm(x.to(dev), g.to(dev), b.to(dev), want_f32=True)
print("layer_norm err:", err(y, torch.nn.functional.layer_norm(x, (384,), g, b)), err(ys.float(), y))

# 1. cfg1: single utterance inference, T=100
xs, il = ofs.synth_text(1, [100])
t0 = time.time()
with torch.no_grad():
    b_ref, a_ref, d_ref, p_ref, e_ref, inter = ofs.fs2_forward(params, cfg, xs, il, is_inference=True, return_intermediates=True)
t_cpu = time.time() - t0
out = model.inference(xs[0].to(dev)); torch.cuda.synchronize()
print("cfg1 frames:", a_ref.shape[1], "gpu frames:", out.shape[0], f"(cpu oracle {t_cpu:.2f}s)")
if out.shape[0] == a_ref.shape[1]:
    e = err(out, a_ref[0]); bad |= e > 1e-3
    print("cfg1 inference err:", e)
else:
    bad = True
# intermediates through _forward
bo, ao, do, po, eo, ol = model._forward(xs.to(dev), il.to(dev), is_inference=True)
print("  durations equal:", torch.equal(do.cpu(), d_ref), " p_outs err:", err(po, p_ref), " e_outs err:", err(eo, e_ref),
      " before err:", err(bo, b_ref))

# 2. teacher-forced padded batch forward (reference training-forward semantics, eval-mode arithmetic)
batch = ofs.synth_train_batch(5, [60, 100, 83, 71])
with torch.no_grad():
    ref = ofs.fs2_forward(params, cfg, batch["text"], batch["text_lengths"], batch["speech_lengths"], batch["durations"],
                          batch["pitch"], batch["energy"])
gb = {k: v.to(dev) for k, v in batch.items()}
o = model(gb["text"], gb["text_lengths"], gb["speech"], gb["speech_lengths"], gb["durations"], gb["pitch"], gb["energy"])
torch.cuda.synchronize()
names = ["before", "after", "d_outs", "p_outs", "e_outs"]
for n, a, r in zip(names, o[:5], ref):
    e = err(a, r); bad |= e > 1e-3
    print(f"forward {n}: err {e:.3e} shape {tuple(a.shape)}")

# 3. batched independent inference == per-utterance inference
lengths = [60, 100, 83, 71, 140, 97]
xs, il = ofs.synth_text(7, lengths)
mel, olens, d = model.batch_inference(xs.to(dev), il.to(dev)); torch.cuda.synchronize()
for i, n in enumerate(lengths):
    with torch.no_grad():
        r = ofs.fs2_inference(params, cfg, xs[i, :n])
    L = int(olens[i])
    if L != r.shape[0]:
        print(f"utt {i}: length mismatch gpu {L} ref {r.shape[0]}"); bad = True; continue
    e = err(mel[i, :L], r); bad |= e > 1e-3
    tail = mel[i, L:].abs().max().item() if L < mel.shape[1] else 0.0
    print(f"batch_inference utt {i} (T={n}, L={L}): err {e:.3e}  tail max {tail:.1e}")

# 4. timing: batch 32, T~U{60..140}
if "--time" in sys.argv:
    g = torch.Generator().manual_seed(3)
    lengths = torch.randint(60, 141, (32,), generator=g).tolist()
    xs, il = ofs.synth_text(3, lengths)
    xs, il = xs.to(dev), il.to(dev)
    for _ in range(3): mel, olens, d = model.batch_inference(xs, il)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 5
    ev[0].record()
    for _ in range(n): mel, olens, d = model.batch_inference(xs, il)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    frames = int(olens.sum())
    print(f"batch32: {ms:.2f} ms, {frames} frames -> {frames / ms * 1e3:.0f} mel-frames/s")
print("FAILED" if bad else "ALL OK")
sys.exit(1 if bad else 0)
