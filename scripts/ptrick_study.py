"""CPU study for the next round: conditioning as h_aux = U (W_aux m') instead of W_aux (U m').

The ConvInUpsampleNet is linear and acts per channel, so the 1x1 aux convolution commutes with it:
    conv1x1_aux(upsample(m'))[t, n] = sum_j U[t, j] * P[j, n],   P = W_aux m'  (frame rate, 128 channels per layer)
U is the (T x frames) matrix of the 4-stage stretch/FIR cascade INCLUDING its zero-padding edge effects.  This script
measures what a kernel needs to know: the band width of U, its period, how far the edge effects reach, and the error of
the re-associated product in fp64 / with split-bf16 operands.  Uses the oracle only (no GPU).
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import pwg

torch.set_grad_enabled(False)
cfg = pwg.DEFAULT_GENERATOR_PARAMS
scales = cfg["upsample_scales"]
hop = 1
for s in scales:
    hop *= s
params = {k: v.double() for k, v in pwg.fold_weight_norm(pwg.synth_params(2, weight_norm=True)).items()}


def upsample_matrix(frames):
    """U (frames*hop x frames): column j = response of the stretch/FIR cascade to an impulse at frame j (one channel)."""
    eye = torch.eye(frames, dtype=torch.float64).unsqueeze(0)           # (1, frames 'channels', frames): channel j = impulse at j
    up = pwg.upsample_net(params, eye, scales)                          # per-channel operator -> (1, frames, T)
    return up[0].transpose(0, 1).contiguous()                           # (T, frames)


Fr = 24
U = upsample_matrix(Fr)
T = Fr * hop
nz = (U.abs() > 0)
first = torch.where(nz.any(1), nz.float().argmax(1), torch.full((T,), -1))
last = Fr - 1 - torch.where(nz.any(1), nz.flip(1).float().argmax(1), torch.full((T,), -1))
width = (last - first + 1)
print(f"hop {hop}, frames {Fr}: band width per row: max {int(width.max())}, min {int(width.min())}")
rel = first - torch.arange(T) // hop
print("first non-zero frame relative to t // hop: min", int(rel.min()), "max", int((last - torch.arange(T) // hop).max()))
# period / shift invariance in the interior
ok_rows = []
for t in range(T - hop):
    a = U[t, : Fr - 1]
    b = U[t + hop, 1:]
    ok_rows.append(bool(torch.equal(a, b)))
ok = torch.tensor(ok_rows)
bad = (~ok).nonzero().flatten()
lo_edge = int(bad[bad < T // 2].max()) + 1 if (bad < T // 2).any() else 0
hi_edge = T - hop - int(bad[bad >= T // 2].min()) if (bad >= T // 2).any() else 0
print(f"shift invariance U[t+hop, j+1] == U[t, j] fails only within {lo_edge} samples of the start and {hi_edge + hop} of the end")

# the re-associated product, exact and with split-bf16 operands
g = torch.Generator().manual_seed(0)
mel = torch.randn(1, 80, Fr + 4, generator=g, dtype=torch.float64)
m1 = F.conv1d(mel, params["upsample_net.conv_in.weight"])               # (1, 80, Fr)
c_up = pwg.upsample_net(params, m1, scales)                              # (1, 80, T)
w_aux = params["conv_layers.7.conv1x1_aux.weight"][:, :, 0]              # (128, 80)
ref = torch.einsum("nc,ct->tn", w_aux, c_up[0])                          # (T, 128)
P = torch.einsum("nc,cj->jn", w_aux, m1[0])                              # (Fr, 128)
alt = U @ P
print("fp64 |U (W m') - W (U m')| / max:", float((alt - ref).abs().max() / ref.abs().max()))


def split(v):
    hi = v.float().bfloat16().double()
    lo = (v - hi).float().bfloat16().double()
    return hi, lo


Uh, Ul = split(U)
Ph, Pl = split(P)
x3 = Uh @ Ph + Ul @ Ph + Uh @ Pl
print("bf16x3 re-associated product error:", float((x3 - ref).abs().max() / ref.abs().max()))
ch, cl = split(c_up[0]); wh, wl = split(w_aux)
cur = torch.einsum("nc,ct->tn", wh, ch) + torch.einsum("nc,ct->tn", wh, cl) + torch.einsum("nc,ct->tn", wl, ch)
print("bf16x3 current formulation error:  ", float((cur - ref).abs().max() / ref.abs().max()))
# K window of a 128-sample tile
worst = 0
for m0 in range(0, T, 128):
    rows = slice(m0, min(m0 + 128, T))
    cols = nz[rows].any(0).nonzero().flatten()
    worst = max(worst, int(cols.max() - cols.min() + 1))
print("frames touched by one 128-sample tile (K of the aux chunk):", worst)


# ----------------------------------------------------------------------------------------------------------------
# Band tables a kernel can consume: row t of U for an utterance of F frames, as K = 8 coefficients starting at frame
# j0(t) = t // hop - 2, assembled from three small tables (interior period, first tile(s), last tile(s)).
# ----------------------------------------------------------------------------------------------------------------
K = 8
EDGE = 128          # rows nearest to either end that get their own coefficients (edge effects reach 95 samples)


def band_rows(Umat):
    Tn, Fn = Umat.shape
    rows = torch.zeros(Tn, K, dtype=Umat.dtype)
    for t in range(Tn):
        j0 = t // hop - 2
        for k in range(K):
            j = j0 + k
            if 0 <= j < Fn:
                rows[t, k] = Umat[t, j]
        # everything outside the window must be zero
        lo, hi = max(j0, 0), min(j0 + K, Fn)
        assert Umat[t, :lo].abs().max() == 0 if lo > 0 else True
        assert Umat[t, hi:].abs().max() == 0 if hi < Fn else True
    return rows


def build_tables(frames_ref=12):
    B_ = band_rows(upsample_matrix(frames_ref))
    Tn = frames_ref * hop
    mid = (frames_ref // 2) * hop
    return dict(interior=B_[mid:mid + hop].clone(), begin=B_[:EDGE].clone(), end=B_[Tn - EDGE:].clone())


def assemble(tables, frames):
    Tn = frames * hop
    rows = torch.empty(Tn, K, dtype=torch.float64)
    for t in range(Tn):
        if t < EDGE:
            rows[t] = tables["begin"][t]
        elif t >= Tn - EDGE:
            rows[t] = tables["end"][t - (Tn - EDGE)]
        else:
            rows[t] = tables["interior"][t % hop]
    return rows


tabs = build_tables()
for frames in (3, 4, 7, 24):
    want = band_rows(upsample_matrix(frames))
    got = assemble(tabs, frames)
    # near the start the window is clipped at frame 0 / near the end at frame F-1: clipped entries are multiplied by
    # out-of-range (zero) rows of P in the kernel (TMA zero fill), so compare only in-range entries
    Tn = frames * hop
    ok = True
    for t in range(Tn):
        j0 = t // hop - 2
        for k in range(K):
            if 0 <= j0 + k < frames and want[t, k] != got[t, k]:
                ok = False
    print(f"frames {frames:3d}: U rebuilt from (interior[{hop}], begin[{EDGE}], end[{EDGE}]) tables bit-exactly: {ok}")
