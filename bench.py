#!/usr/bin/env python
"""bench.py - headline benchmark of parakeet_b200 (contract in the task statement).

Workload (BASELINE.json configs[1]): Parallel WaveGAN generator inference, batch 32, 80-mel x 400 frames -> 3.84 M
samples of 24 kHz audio per step, CSMSC generator (30 residual layers, 64/128 channels, upsample [4,5,3,5]), random
weights of that architecture, synthetic N(0,1) mel + noise.  One step = one pass of the generator over one batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); utterances are independent, so every rank runs its own batch of 32
with no data-path collective (weak scaling) and `value` is the whole-job aggregate.
`--impl reference` times the reference algorithm's CPU path (the torch-CPU oracle restatement; PaddlePaddle itself is not
installable here) with all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH, FRAMES, HOP = 32, 400, 300
FLOP_PER_SAMPLE_LAYER = 2 * (128 * (3 * 64 + 80) + 128 * 64)       # 86 016: conv k3 + aux 1x1 + skip/out 1x1
FLOP_PER_SAMPLE = 30 * FLOP_PER_SAMPLE_LAYER + 2 * 64 * 64 + 2 * 64 + 2 * 64  # 2 588 928


def peaks():
    p = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(hbm_gbs=m["hbm_gbs"], bf16_tflops=m["bf16_tflops"],
                 bf16_tflops_sustained=m.get("bf16_tflops_sustained", m["bf16_tflops"]), source="measured (MEASURED_PEAKS.json)")
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_reference_step(params, x, c):
    import torch
    from oracle import pwg as opwg
    with torch.no_grad():
        return opwg.generator_forward(params, x, c)


def pick_cpu_threads(params, x, c):
    """The CPU path is timed at its best thread count (oversubscribing small convs on a 100+ core host is slower)."""
    import torch
    cores = os.cpu_count() or 1
    best, best_t = cores, None
    xs, cs = x[:1, :, :HOP * 40].contiguous(), c[:1, :, :44].contiguous()
    for n in sorted({cores, max(cores // 2, 1), max(cores // 4, 1), 32, 16, 8} & set(range(1, cores + 1)), reverse=True):
        torch.set_num_threads(n)
        cpu_reference_step(params, xs, cs)
        t0 = time.perf_counter()
        cpu_reference_step(params, xs, cs)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def run_reference(args, rank):
    """The reference algorithm's CPU implementation (oracle restatement) on the host cores, bounded sample per step."""
    if rank != 0:
        return
    import torch
    from oracle import pwg as opwg
    params = opwg.fold_weight_norm(opwg.synth_params(2, weight_norm=True))
    b_s = 1                                                   # bounded sample: 1 of the 32 utterances per step (~5 s)
    x, c = opwg.synth_inputs(2, batch=b_s, mel_frames=FRAMES)
    cores = pick_cpu_threads(params, x, c)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_step(params, x, c)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(params, x, c)
    dt = time.perf_counter() - t0
    v = b_s * FRAMES * HOP * args.steps / dt
    sample = (f"{b_s} of {BATCH} utterances (400 mel frames each) per step, torch-CPU fp32, best of several thread counts = "
              f"{cores} of {os.cpu_count()} host threads")
    print(json.dumps({
        "impl": "reference", "metric": "audio-samples/sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "pwg_generator_b32_mel400_24k", "global_batch": BATCH * args.gpus, "mel_frames": FRAMES, "hop": HOP},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the FastSpeech2 / end-to-end extras and the CPU baseline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from parakeet_b200 import _lib
    from parakeet_b200.models import PWGGenerator

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    # CSMSC generator_params (examples/GANVocoder/parallelwave_gan/baker/conf/default.yaml:23-45), random-init weights
    gen = PWGGenerator(layers=30, stacks=3, residual_channels=64, gate_channels=128, skip_channels=64, aux_channels=80,
                       aux_context_window=2, upsample_scales=[4, 5, 3, 5], use_weight_norm=True, device=dev, seed=2)
    gen.remove_weight_norm()                     # as synthesize.py does before inference
    g_in = torch.Generator().manual_seed(1002 + rank)
    c_h = torch.randn(BATCH, 80, FRAMES, generator=g_in)
    c_h = torch.cat([c_h[:, :, :1].expand(-1, -1, 2), c_h, c_h[:, :, -1:].expand(-1, -1, 2)], dim=-1).contiguous()  # replicate pad
    x_h = torch.randn(BATCH, 1, FRAMES * HOP, generator=g_in)
    x_h, c_h = x_h.pin_memory(), c_h.pin_memory()
    x, c = x_h.to(dev), c_h.to(dev)
    samples_per_step = BATCH * FRAMES * HOP

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---------------- device-resident timed region ----------------
    for _ in range(args.warmup):
        y = gen(x, c)
    gen._layer_events = []                     # (start, end) CUDA events around the 30 residual-layer launches of each step
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.time()
    while not sampler.lines and time.time() - t_wait < 3.0:     # nvidia-smi needs a moment to emit its first sample
        y = gen(x, c)
        torch.cuda.synchronize()
    barrier()
    first_line = len(sampler.lines)
    launches0 = lib.pk_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        y = gen(x, c)
    e1.record()
    barrier()
    launches = lib.pk_launch_count() - launches0
    sampler.lines = sampler.lines[max(first_line - 1, 0):]      # samples taken during the timed region (+ the one straddling its start)
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    layer_ms = [a.elapsed_time(b) for a, b in gen._layer_events]
    gen._layer_events = None
    ms_per_step = ms_total / args.steps
    value = world * samples_per_step * args.steps / (ms_total * 1e-3)

    # ---------------- end to end through the public API with host buffers ----------------
    wav_h = torch.empty(BATCH, 1, FRAMES * HOP, dtype=torch.float32).pin_memory()
    for _ in range(2):
        wav_h.copy_(gen(x_h.to(dev, non_blocking=True), c_h.to(dev, non_blocking=True)), non_blocking=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        xd = x_h.to(dev, non_blocking=True)
        cd = c_h.to(dev, non_blocking=True)
        wav_h.copy_(gen(xd, cd), non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * samples_per_step * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    layer_launch_ms = sum(layer_ms) / max(len(layer_ms), 1) / 30.0            # average residual-layer kernel duration
    flops_per_launch = FLOP_PER_SAMPLE_LAYER * samples_per_step               # algorithmic (one pass), 330 GFLOP
    achieved_tf = flops_per_launch / (layer_launch_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "pk::pwg_layer_pair_kernel", "achieved": achieved_tf, "peak": pk["bf16_tflops_sustained"],
                "unit": "TFLOP/s", "frac": achieved_tf / pk["bf16_tflops_sustained"],
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this exact config (B=32, 400 frames), from
                # `ncu --set full` (profiles/r01_pwg_layer_pair_b32_ncu.txt); algorithmic bytes are 5.16e9
                "traffic": 5.122e9, "traffic_unit": "bytes/launch",
                "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "launch_ms": layer_launch_ms, "launches_per_step": 30,
                "note": "algorithmic FLOPs; split-bf16 operands execute 3 tensor-core passes per product (+ the residual pass)",
                "hbm_algorithmic_gbs": (samples_per_step * (256 + 256 + 512 + 320)) / (layer_launch_ms * 1e-3) / 1e9,
                "hbm_peak_gbs": pk["hbm_gbs"]}

    out = {"metric": "audio-samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16x3 (split-bf16 operands, fp32 accumulate; fp32-grade, 1e-3 parity)", "data": "synthetic",
           "config": {"workload": "pwg_generator_b32_mel400_24k", "global_batch": BATCH * world, "per_gpu_batch": BATCH,
                      "mel_frames": FRAMES, "hop": HOP, "parallelism": f"batch-sharded x{world}, no collective",
                      "l2": "inputs and working set (4.2 GB) larger than L2; no flush needed"},
           "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                   "h2d_bytes_per_step": int(x_h.numel() * 4 + c_h.numel() * 4), "d2h_bytes_per_step": int(wav_h.numel() * 4),
                   "api": "PWGGenerator.forward(x, c) with pinned host tensors"},
           "roofline": roofline}

    if not args.no_extra and world == 1:      # single-GPU side metrics; the scaling runs report the headline only
        # FastSpeech2 (cfg3's acoustic half) and the FS2 -> PWG pipeline, reported alongside the headline
        try:
            import math
            from parakeet_b200.models import FastSpeech2
            # LJSpeech yaml (examples/fastspeech2/ljspeech/conf/default.yaml:33-75), vocab 80, random-init weights
            fs = FastSpeech2(80, 80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536,
                             positionwise_layer_type="conv1d", positionwise_conv_kernel_size=3, duration_predictor_layers=2,
                             duration_predictor_chans=256, duration_predictor_kernel_size=3, postnet_layers=5, postnet_filts=5,
                             postnet_chans=256, pitch_predictor_layers=5, pitch_predictor_chans=256, pitch_predictor_kernel_size=5,
                             pitch_embed_kernel_size=1, energy_predictor_layers=2, energy_predictor_chans=256,
                             energy_predictor_kernel_size=3, energy_embed_kernel_size=1, device=dev, seed=1)
            sd = dict(fs.state_dict())
            sd["duration_predictor.linear.bias"] = torch.tensor([math.log(8.0)])   # predicted durations ~7 frames / phoneme
            fs.set_state_dict(sd)
            g = torch.Generator().manual_seed(3)
            lengths = torch.randint(60, 141, (BATCH,), generator=g).tolist()
            ids = torch.zeros(BATCH, max(lengths), dtype=torch.int64)
            for i, n in enumerate(lengths):
                ids[i, :n] = torch.randint(1, 79, (n,), generator=g)
            il = torch.tensor(lengths, dtype=torch.int64)
            ids, il = ids.to(dev), il.to(dev)
            for _ in range(3):
                mel, olens, _ = fs.batch_inference(ids, il)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                mel, olens, _ = fs.batch_inference(ids, il)
            e1.record()
            torch.cuda.synchronize()
            fs_ms = e0.elapsed_time(e1) / args.steps
            frames = int(olens.sum())

            def tts():
                mel, olens, _ = fs.batch_inference(ids, il)
                L = mel.shape[1]
                cc = mel.transpose(1, 2)
                cc = torch.cat([cc[:, :, :1].expand(-1, -1, 2), cc, cc[:, :, -1:].expand(-1, -1, 2)], dim=-1).contiguous()
                noise = torch.randn(BATCH, 1, L * HOP, device=dev)
                return gen(noise, cc, lens=(olens * HOP).to(torch.int32)), olens
            for _ in range(2):
                tts()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(max(args.steps // 2, 1)):
                wav, olens = tts()
            e1.record()
            torch.cuda.synchronize()
            tts_ms = e0.elapsed_time(e1) / max(args.steps // 2, 1)
            out["extra"] = {"fastspeech2_b32": {"mel_frames_per_s": frames / (fs_ms * 1e-3), "ms_per_step": fs_ms, "frames": frames},
                            "fs2_pwg_e2e_b32": {"samples_per_s": frames * HOP / (tts_ms * 1e-3), "mel_frames_per_s": frames / (tts_ms * 1e-3),
                                                "ms_per_step": tts_ms, "note": "cfg3, ragged batch, per-GPU"}}
            # WaveFlow (cfg4): 64 channels, 8 flows x 8 layers, n_group 16, batch 16, 400 mel frames -> 102 128 samples each
            from parakeet_b200.models import ConditionalWaveFlow
            wf = ConditionalWaveFlow([16, 16], 8, 8, 16, 64, 80, (3, 3), device=dev, seed=4)
            sdw = dict(wf.state_dict())
            gw = torch.Generator().manual_seed(4)
            for k_ in sdw:
                if "output_proj" in k_:      # the reference zero-initialises these (identity flow); use small random values
                    sdw[k_] = (torch.rand(sdw[k_].shape, generator=gw) * 2 - 1) * 0.05
            wf.set_state_dict(sdw)
            melw = (torch.randn(16, 80, FRAMES, generator=gw) * 0.5 - 3).to(dev)
            zw = torch.randn(16, 256 * FRAMES - 272, generator=gw).to(dev)
            for _ in range(3):               # eager call, graph capture, first replay
                wf.infer(melw, z=zw)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                aw = wf.infer(melw, z=zw)
            e1.record()
            torch.cuda.synchronize()
            wf_ms = e0.elapsed_time(e1) / 3
            out["extra"]["waveflow_b16_c64"] = {"samples_per_s": aw.numel() / (wf_ms * 1e-3), "ms_per_step": wf_ms,
                                                "note": "cfg4; one CUDA graph of ~2 300 kernel nodes per call (fused GEMM epilogues)"}
        except Exception as ex:  # extras must never break the headline line
            out.setdefault("extra", {})["error"] = repr(ex)
    if not args.no_extra and world == 1:
        # CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample (2 of 32 utterances, ~10-20 s)
        params = {k: v.detach().cpu() for k, v in gen.state_dict().items()}   # same weights, oracle arithmetic
        xs, cs = x_h[:2].clone(), c_h[:2].clone()
        cores = pick_cpu_threads(params, xs, cs)
        t0 = time.perf_counter()
        cpu_reference_step(params, xs, cs)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 2 * FRAMES * HOP / dt, "unit": "samples/s", "cores": cores, "kind": "port",
                               "sample": f"2 of 32 utterances x 400 frames, torch-CPU fp32 oracle, one pass, best thread count {cores} of {os.cpu_count()}"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
