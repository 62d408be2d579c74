"""CPU emulation of the whole 30-layer generator with frame-rate conditioning (tile-wise band-table x P-window products,
split-bf16 operands) against the current formulation and fp64: precision of DESIGN.md 7.2 (1.4e-5 vs 1.2e-5)."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from oracle import pwg
from parakeet_b200.models import _pwg_frame_cond as fc
torch.manual_seed(0)
P = pwg.fold_weight_norm(pwg.synth_params(2, weight_norm=True))
frames = 12
x, c = pwg.synth_inputs(2, batch=2, mel_frames=frames)
P64 = {k: v.double() for k, v in P.items()}
cfg = pwg.DEFAULT_GENERATOR_PARAMS
scales = cfg["upsample_scales"]; hop = 300
def bf(v): return v.float().bfloat16().double()
def split(v):
    hi = bf(v); lo = bf(v - hi); return hi, lo
def x3(a, b, f):   # bf16x3 product of two operands through bilinear f
    ah, al = split(a); bh, bl = split(b)
    return f(ah, bh) + f(al, bh) + f(ah, bl)
firs = [P64[f"upsample_net.upsample.up_layers.{2*i+1}.weight"].reshape(-1) for i in range(4)]
table = fc.tile_band_table(firs, scales, frames)          # (T, 16) tile-relative
T = frames * hop
m1 = F.conv1d(c.double(), P64["upsample_net.conv_in.weight"])       # (B, 80, frames)
c_up = pwg.upsample_net(P64, m1, scales)
def run(mode):
    layers, lps = cfg["layers"], cfg["layers"] // cfg["stacks"]
    h = F.conv1d(x.double(), P64["first_conv.weight"], P64["first_conv.bias"])
    skips = 0
    for i in range(layers):
        pre = f"conv_layers.{i}."; d = 2 ** (i % lps)
        w = P64[pre + "conv.weight"]
        if mode == "exact":
            t = F.conv1d(h, w, P64[pre + "conv.bias"], padding=d, dilation=d) + F.conv1d(c_up, P64[pre + "conv1x1_aux.weight"])
        else:
            t = x3(h, w, lambda a, b: F.conv1d(a, b, padding=d, dilation=d)) + P64[pre + "conv.bias"][None, :, None]
            wa = P64[pre + "conv1x1_aux.weight"][:, :, 0]
            if mode == "cur":
                t = t + x3(c_up, P64[pre + "conv1x1_aux.weight"], lambda a, b: F.conv1d(a, b))
            else:   # frame-rate conditioning: P via bf16x3 GEMM of split operands, then split again; table split; tile-wise product
                Pm = x3(m1, wa, lambda a, b: torch.einsum("bcj,nc->bjn", a, b))           # (B, frames, 128)
                Ppad = torch.zeros(Pm.shape[0], frames + 32, 128, dtype=torch.float64); Ppad[:, 16:16 + frames] = Pm
                aux = torch.zeros(Pm.shape[0], T, 128, dtype=torch.float64)
                for t0 in range(0, T, 128):
                    j0 = t0 // hop - 2
                    aux[:, t0:t0 + 128] = x3(table[t0:t0 + 128], Ppad[:, 16 + j0:16 + j0 + 16], lambda a, b: torch.einsum("tk,bkn->btn", a, b))
                t = t + aux.transpose(1, 2)
        a, b = torch.chunk(t, 2, dim=1)
        z = torch.tanh(a) * torch.sigmoid(b)
        f11 = (lambda a, b: F.conv1d(a, b))
        if mode == "exact":
            sk = F.conv1d(z, P64[pre + "conv1x1_skip.weight"]); ou = F.conv1d(z, P64[pre + "conv1x1_out.weight"])
        else:
            sk = x3(z, P64[pre + "conv1x1_skip.weight"], f11); ou = x3(z, P64[pre + "conv1x1_out.weight"], f11)
        skips = skips + sk + P64[pre + "conv1x1_skip.bias"][None, :, None]
        h = (ou + P64[pre + "conv1x1_out.bias"][None, :, None] + h) * math.sqrt(0.5)
    y = F.relu(skips * math.sqrt(1.0 / layers))
    y = F.relu(F.conv1d(y, P64["last_conv_layers.1.weight"], P64["last_conv_layers.1.bias"]))
    return F.conv1d(y, P64["last_conv_layers.3.weight"], P64["last_conv_layers.3.bias"])
ref = run("exact")
for m in ("cur", "fc"):
    y = run(m)
    print(m, "%.2e" % ((y - ref).abs().max() / ref.abs().max()).item())
