"""Stage-by-stage GPU diagnostic for the frame-rate conditioning path: P GEMM vs torch, then single fc layers with a sync each."""
import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from oracle import pwg as opwg
from parakeet_b200 import _lib, ops
from parakeet_b200.models import PWGGenerator
from parakeet_b200.models import _pwg_frame_cond as fc
from parakeet_b200.ops import Split

dev = "cuda"
def err(a, b):
    return ((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()

cfg = dict(opwg.DEFAULT_GENERATOR_PARAMS)
params = opwg.synth_params(2, weight_norm=True)
gen = PWGGenerator(**cfg, device=dev)
gen.set_state_dict(params)
folded = opwg.fold_weight_norm(params)
frames, batch = int(os.environ.get("FRAMES", 40)), int(os.environ.get("BATCH", 2))
x, c = opwg.synth_inputs(2, batch=batch, mel_frames=frames)
with torch.no_grad():
    ref, inter = opwg.generator_forward(folded, x, c, return_intermediates=True)
os.environ["PK_PWG_FRAME_COND"] = "0"
y0 = gen(x.to(dev), c.to(dev)).clone()
torch.cuda.synchronize()
print("default path err", err(y0, ref), flush=True)

# --- the fc path, by hand ---
L = _lib.lib()
pk = gen._pack()
B, T = batch, frames * 300
ws = gen._workspace(B, T)
hop, A, NL = 300, 80, 30
fp = {k: v.detach().float().cpu() for k, v in gen._folded().items()}
aux_all = torch.cat([fp[f"conv_layers.{i}.conv1x1_aux.weight"][:, :, 0] for i in range(NL)], dim=0)
aux_s = Split.from_f32(aux_all.contiguous().to(dev).unsqueeze(0))
m1 = ws["conv_in"].clone()                      # (B, frames, aux) from the default run above
m1s = Split.from_f32(m1)
Fp = max((frames + 7) // 8 * 8, 64)
P = Split.zeros((B, NL * 128, Fp), dev)
a_spec = dict(rows=NL * 128, cols=A, ld=A, batch_stride=0, batches=1, bmul=0, hmul=0, col0=0, colh=0)
b_spec = dict(rows=frames, cols=A, ld=A, batch_stride=frames * A, batches=B, bmul=1, hmul=0, col0=0, colh=0)
ops.batched_matmul_nt(aux_s, m1s, batch=B, heads=1, m=NL * 128, n=frames, k=A, a_spec=a_spec, b_spec=b_spec,
                      y_split=P, y_batch_stride=NL * 128 * Fp, y_head_stride=0, y_ld=Fp)
torch.cuda.synchronize()
P_ref = torch.einsum("nk,bfk->bnf", aux_all.double(), m1.double().cpu())
print("P gemm err", err((P.hi.float() + P.lo.float())[:, :, :frames], P_ref), flush=True)

firs, off = [], 0
for s_ in gen.upsample_scales:
    firs.append(torch.from_numpy(pk["fir_host"][off:off + 2 * s_ + 1].copy()))
    off += 2 * s_ + 1
tb = torch.zeros(frames * hop, 64, dtype=torch.float32)
tb[:, :fc.KWIN] = fc.tile_band_table(firs, gen.upsample_scales, frames).float()
U = Split.from_f32(tb.to(dev).unsqueeze(0))

# layer input: first conv output from the default run is gone (ping-pong); recompute
xin = Split.zeros((B, T, 64), dev)
_lib.check(L.pk_pwg_first_conv(x.to(dev).contiguous().data_ptr(), pk["first_w"].data_ptr(), pk["first_b"].data_ptr(), None, B, T,
                               xin.hi.data_ptr(), xin.lo.data_ptr(), None), "first")
torch.cuda.synchronize()
x0_ref = inter["x_layers"][0] if "x_layers" in inter else None
yout = Split.zeros((B, T, 64), dev)
skip = torch.zeros(B, T, 64, device=dev)
args = _lib.PwgLayerFcArgs()
args.batch, args.t, args.hop = B, T, hop
args.lens = None
args.u_hi, args.u_lo, args.u_batches = U.hi.data_ptr(), U.lo.data_ptr(), 1
args.p_hi, args.p_lo, args.p_rows, args.p_ld, args.p_frames = P.hi.data_ptr(), P.lo.data_ptr(), NL * 128, Fp, frames
args.skip = skip.data_ptr()
args.prof = None
src, dst = xin, yout
nlayers = int(os.environ.get("NLAYERS", 3))
for i in range(nlayers):
    lay = pk["layers"][i]
    args.dilation, args.p_row0 = lay["dil"], i * 128
    args.x_hi, args.x_lo, args.y_hi, args.y_lo = src.hi.data_ptr(), src.lo.data_ptr(), dst.hi.data_ptr(), dst.lo.data_ptr()
    args.w1_hi, args.w1_lo = lay["w1"].hi.data_ptr(), lay["w1"].lo.data_ptr()
    args.w2_hi, args.w2_lo = lay["w2"].hi.data_ptr(), lay["w2"].lo.data_ptr()
    args.bias1, args.bias2 = lay["b1"].ctypes.data, lay["b2"].ctypes.data
    args.skip_init = 1 if i == 0 else 0
    print("launch fc layer", i, flush=True)
    _lib.check(L.pk_pwg_residual_layer_fc(C.byref(args), None), "fc")
    torch.cuda.synchronize()
    got = (dst.hi.float() + dst.lo.float()).transpose(1, 2)
    print(f"layer {i}: x err vs oracle {err(got, inter['x_layers'][i]):.3e}", flush=True)
    src, dst = dst, src
print("done", flush=True)
