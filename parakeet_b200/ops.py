"""Torch-tensor level wrappers over the C-ABI (device memory + streams are torch's; all math is in the .so)."""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import PK_ACT_NONE, PK_ACT_RELU, PK_ACT_TANH, ConvGemmArgs, Operand  # noqa: F401

ACTS = {None: PK_ACT_NONE, "none": PK_ACT_NONE, "relu": PK_ACT_RELU, "tanh": PK_ACT_TANH}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.PkError("parakeet_b200 ops need CUDA tensors (no CPU fallback)")


class Split:
    """split-bf16 tensor: value = hi + lo (two bf16 planes of identical shape / strides)."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo):
        assert hi.dtype == torch.bfloat16 and lo.dtype == torch.bfloat16 and hi.shape == lo.shape
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape

    # Both planes come from ONE allocation, lo right after hi: kernels that stream a tile of both planes can then fetch them
    # with a single 4-D TMA box (the plane is the outermost dimension; csrc/pwg_fc.cu) instead of two loads.
    @staticmethod
    def empty(shape, device):
        buf = torch.empty((2,) + tuple(shape), dtype=torch.bfloat16, device=device)
        return Split(buf[0], buf[1])

    @staticmethod
    def zeros(shape, device):
        buf = torch.zeros((2,) + tuple(shape), dtype=torch.bfloat16, device=device)
        return Split(buf[0], buf[1])

    @staticmethod
    def from_f32(x):
        """fp32 CUDA tensor -> split planes (pk_split_f32)."""
        _require_cuda(x)
        x = x.contiguous().float()
        out = Split.empty(x.shape, x.device)
        _lib.check(_lib.lib().pk_split_f32(_ptr(x), _ptr(out.hi), _ptr(out.lo), x.numel(), _stream()), "pk_split_f32")
        return out

    def float(self):
        return self.hi.float() + self.lo.float()


def pack_weight(w, device=None):
    """Conv1D weight [out, in, k] (Paddle/torch layout) or Linear weight given as [out, in] ->
    K-major GEMM operand [out, k * Kp] (tap-major, each tap's channels zero-padded to a multiple of 64), split-bf16.
    Host-side, done once at load time."""
    w = w.detach().float().cpu()
    if w.dim() == 2:
        w = w.unsqueeze(-1)
    n, k, taps = w.shape
    kp = (k + 63) // 64 * 64
    packed = torch.zeros(n, taps, kp, dtype=torch.float32)
    packed[:, :, :k] = w.permute(0, 2, 1)
    packed = packed.reshape(n, taps * kp)
    hi = packed.to(torch.bfloat16)
    lo = (packed - hi.float()).to(torch.bfloat16)
    dev = device or "cuda"
    return Split(hi.to(dev).contiguous(), lo.to(dev).contiguous())


def _operand(s, rows, cols, ld, batch_stride, batches, bmul=1, hmul=0, col0=0, colh=0):
    return Operand(hi=s.hi.data_ptr(), lo=s.lo.data_ptr(), batch_stride=batch_stride, ld=ld, rows=rows, cols=cols,
                   batches=batches, bmul=bmul, hmul=hmul, col0=col0, colh=colh)


def conv_gemm(a, w, *, n, k, taps=1, dil=1, pad=None, bias=None, act=None, residual=None, lens=None, scale=1.0,
              out_f32=True, out_split=False, passes=3, simt=False, y_f32=None, y_split=None, epilogue=None):
    """Channels-last Conv1D / Linear.  a: Split (B, T, C_total); w: packed weight Split [n, taps*Kp].

    Returns (y_f32 or None, y_split or None), each (B, T, n).
    `epilogue` selects a fused pair epilogue of pk_conv_gemm_ex (n == 2 * C, the GEMM result itself is not written):
      dict(mode="gate", channels=C, residual=(tensor (B, T, >= 2C) fp32 view with last stride 1) or None) -> y_split (B, T, C)
      dict(mode="wf_update", channels=C, state=, skip=, skip_init=bool, buf=Split or None, buf_col0=int)
    """
    _require_cuda(a.hi, w.hi)
    B, T, Ctot = a.hi.shape
    if pad is None:
        pad = (taps - 1) // 2
    dev = a.hi.device
    if epilogue is not None:
        out_f32 = False
        out_split = False
        y_f32 = None
        if epilogue["mode"] == "gate" and y_split is None:
            y_split = Split.empty((B, T, epilogue["channels"]), dev)
    if out_f32 and y_f32 is None:
        y_f32 = torch.empty(B, T, n, dtype=torch.float32, device=dev)
    if out_split and y_split is None:
        y_split = Split.empty((B, T, n), dev)
    args = ConvGemmArgs()
    args.a = _operand(a, rows=T, cols=Ctot, ld=a.hi.stride(1), batch_stride=a.hi.stride(0), batches=B)
    args.b = _operand(w, rows=w.hi.shape[0], cols=w.hi.shape[1], ld=w.hi.stride(0), batch_stride=0, batches=1, bmul=0)
    args.batch, args.heads, args.m, args.n, args.k = B, 1, T, n, k
    args.taps, args.dil, args.pad = taps, dil, pad
    args.scale = scale
    args.bias = bias.data_ptr() if bias is not None else None
    args.act = ACTS[act]
    args.residual = residual.data_ptr() if residual is not None else None
    args.lens = lens.data_ptr() if lens is not None else None
    args.y_f32 = y_f32.data_ptr() if y_f32 is not None else None
    args.y_hi = y_split.hi.data_ptr() if y_split is not None else None
    args.y_lo = y_split.lo.data_ptr() if y_split is not None else None
    args.y_batch_stride, args.y_head_stride, args.y_ld = T * n, 0, n
    args.passes = passes
    if epilogue is not None:
        ep = _lib.GemmEpilogue()
        Cc = int(epilogue["channels"])
        ep.channels = Cc
        if epilogue["mode"] == "gate":
            ep.mode = _lib.PK_EPI_GATE
            args.y_batch_stride, args.y_ld = T * y_split.hi.stride(1), y_split.hi.stride(1)
            res = epilogue.get("residual")
            if res is not None:
                assert res.dtype == torch.float32 and res.stride(-1) == 1 and res.shape[0] == B and res.shape[1] == T
                ep.residual, ep.residual_batch_stride, ep.residual_ld = res.data_ptr(), res.stride(0), res.stride(1)
        else:
            ep.mode = _lib.PK_EPI_WF_UPDATE
            state, skip = epilogue["state"], epilogue["skip"]
            assert state.is_contiguous() and skip.is_contiguous() and tuple(state.shape) == (B, T, Cc) == tuple(skip.shape)
            ep.state, ep.skip, ep.skip_init = state.data_ptr(), skip.data_ptr(), 1 if epilogue.get("skip_init") else 0
            buf = epilogue.get("buf")
            if buf is not None:
                assert buf.hi.is_contiguous() and buf.hi.shape[0] == B and buf.hi.shape[1] == T
                ep.buf_hi, ep.buf_lo, ep.buf_ld, ep.buf_col0 = buf.hi.data_ptr(), buf.lo.data_ptr(), buf.hi.stride(1), int(epilogue.get("buf_col0", 0))
        _lib.check(_lib.lib().pk_conv_gemm_ex(C.byref(args), C.byref(ep), _stream()), "pk_conv_gemm_ex")
        return None, y_split
    fn = _lib.lib().pk_conv_gemm_simt if simt else _lib.lib().pk_conv_gemm
    _lib.check(fn(C.byref(args), _stream()), "pk_conv_gemm")
    return y_f32, y_split


def batched_matmul_nt(a, b, *, batch, heads, m, n, k, a_spec, b_spec, scale=1.0, y_f32=None, y_split=None,
                      y_batch_stride=None, y_head_stride=None, y_ld=None, lens=None, passes=3, simt=False):
    """y[b,h] = scale * A[b,h] (m x k) . B[b,h]^T (n x k); operand addressing given by a_spec / b_spec dicts
    (rows, cols, ld, batch_stride, batches, bmul, hmul, col0, colh)."""
    args = ConvGemmArgs()
    args.a = _operand(a, **a_spec)
    args.b = _operand(b, **b_spec)
    args.batch, args.heads, args.m, args.n, args.k = batch, heads, m, n, k
    args.taps, args.dil, args.pad = 1, 1, 0
    args.scale = scale
    args.act = PK_ACT_NONE
    args.lens = lens.data_ptr() if lens is not None else None
    args.y_f32 = y_f32.data_ptr() if y_f32 is not None else None
    args.y_hi = y_split.hi.data_ptr() if y_split is not None else None
    args.y_lo = y_split.lo.data_ptr() if y_split is not None else None
    args.y_batch_stride, args.y_head_stride, args.y_ld = y_batch_stride, y_head_stride, y_ld
    args.passes = passes
    fn = _lib.lib().pk_conv_gemm_simt if simt else _lib.lib().pk_conv_gemm
    _lib.check(fn(C.byref(args), _stream()), "pk_conv_gemm")


def length_regulator_lens(dur):
    """dur: int64 (B, T) CUDA -> int32 (B,) total frames per utterance (device tensor, no sync)."""
    _require_cuda(dur)
    dur = dur.contiguous()
    B, T = dur.shape
    out = torch.empty(B, dtype=torch.int32, device=dur.device)
    _lib.check(_lib.lib().pk_length_regulator_lens(_ptr(dur), B, T, _ptr(out), _stream()), "pk_length_regulator_lens")
    return out


def length_regulate(x, dur, t_out, want_f32=True, want_split=False):
    """x fp32 (B, T, C), dur int64 (B, T) -> (B, t_out, C) repeat-expanded; rows past sum(d) are zero."""
    _require_cuda(x, dur)
    x = x.contiguous()
    dur = dur.contiguous()
    B, T, Cc = x.shape
    y = torch.empty(B, t_out, Cc, dtype=torch.float32, device=x.device) if want_f32 else None
    ys = Split.empty((B, t_out, Cc), x.device) if want_split else None
    if t_out == 0:   # every duration is zero: empty output, as in the reference (t_dec = 0)
        return y, ys
    _lib.check(_lib.lib().pk_length_regulate(_ptr(x), _ptr(dur), B, T, Cc, t_out, _ptr(y),
                                             _ptr(ys.hi) if ys else None, _ptr(ys.lo) if ys else None, _stream()),
               "pk_length_regulate")
    return y, ys


# ----------------------------------------------------------------------------------------------------------------
# FastSpeech2 row-wise kernels
# ----------------------------------------------------------------------------------------------------------------
def embed_pe(ids, table, x_in, alpha, lens, padding_idx=0):
    """Embedding(padding_idx)+ScaledPositionalEncoding (ids given) or ScaledPositionalEncoding only (x_in given)."""
    if ids is not None:
        B, T = ids.shape
        d = table.shape[1]
        dev = ids.device
        ids = ids.contiguous()
    else:
        B, T, d = x_in.shape
        dev = x_in.device
        x_in = x_in.contiguous()
    y = torch.empty(B, T, d, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().pk_embed_pe(_ptr(ids), _ptr(table), table.shape[0] if table is not None else 0, padding_idx,
                                      _ptr(x_in), _ptr(alpha), _ptr(lens), B, T, d, _ptr(y), _stream()), "pk_embed_pe")
    return y


def layer_norm(x, gamma, beta, lens=None, want_f32=False, want_split=True, eps=1e-5):
    B, T, d = x.shape
    y = torch.empty_like(x) if want_f32 else None
    ys = Split.empty((B, T, d), x.device) if want_split else None
    _lib.check(_lib.lib().pk_layer_norm(_ptr(x), _ptr(gamma), _ptr(beta), eps, _ptr(lens), B, T, d, _ptr(y),
                                        _ptr(ys.hi) if ys else None, _ptr(ys.lo) if ys else None, _stream()), "pk_layer_norm")
    return y, ys


def masked_softmax(s, key_lens, batch, heads, rows, keys):
    """s fp32 (batch*heads, rows, ld) -> split planes of the same shape."""
    ld = s.shape[-1]
    p = Split.empty(tuple(s.shape), s.device)
    _lib.check(_lib.lib().pk_masked_softmax(_ptr(s), _ptr(key_lens), batch, heads, rows, keys, ld, _ptr(p.hi), _ptr(p.lo),
                                            _stream()), "pk_masked_softmax")
    return p


def transpose_heads(src, col0, dk, heads, ld_dst):
    B, T, ld_src = src.hi.shape
    dst = Split.empty((B * heads, dk, ld_dst), src.hi.device)
    _lib.check(_lib.lib().pk_transpose_heads(_ptr(src.hi), _ptr(src.lo), B, T, ld_src, col0, dk, heads, ld_dst, _ptr(dst.hi),
                                             _ptr(dst.lo), _stream()), "pk_transpose_heads")
    return dst


def l2_normalize_axis1(x, eps=1e-12):
    """paddle F.normalize(x) (p=2, axis=1): (B, D) over D; (B, T, D) over T (the reference's batched tone path)."""
    x = x.contiguous().float()
    outer, n = x.shape[0], x.shape[1]
    inner = x.numel() // (outer * n)
    y = torch.empty_like(x)
    _lib.check(_lib.lib().pk_l2_normalize(_ptr(x), outer, n, inner, float(eps), _ptr(y), _stream()), "pk_l2_normalize")
    return y


def fused_attention(qkv, heads, key_lens=None, row_lens=None, ctx=None):
    """qkv Split (B, T, 3A) -> ctx Split (B, T, A): softmax(q k^T / sqrt(d_k), key mask) v per head, one kernel
    (pk_fused_attention) after the per-head transpose of v."""
    B, T, ld = qkv.hi.shape
    A = ld // 3
    dk = A // heads
    Tp = (T + 63) // 64 * 64
    vt = transpose_heads(qkv, col0=2 * A, dk=dk, heads=heads, ld_dst=Tp)
    if ctx is None:
        ctx = Split.empty((B, T, A), qkv.hi.device)
    _lib.check(_lib.lib().pk_fused_attention(_ptr(qkv.hi), _ptr(qkv.lo), _ptr(vt.hi), _ptr(vt.lo), B, T, heads, dk, Tp, _ptr(key_lens),
                                             _ptr(row_lens), 1.0 / math.sqrt(dk), _ptr(ctx.hi), _ptr(ctx.lo), _stream()),
               "pk_fused_attention")
    return ctx


def duration_post(x, lens, offset=1.0):
    B, T = x.shape
    x = x.contiguous()
    d_f = torch.empty(B, T, dtype=torch.float32, device=x.device)
    d_i = torch.empty(B, T, dtype=torch.int64, device=x.device)
    _lib.check(_lib.lib().pk_duration_post(_ptr(x), _ptr(lens), B, T, offset, _ptr(d_f), _ptr(d_i), _stream()),
               "pk_duration_post")
    return d_f, d_i


def duration_scale(d, alpha):
    d = d.contiguous()
    out = torch.empty_like(d)
    _lib.check(_lib.lib().pk_duration_scale(_ptr(d), alpha, d.numel(), _ptr(out), _stream()), "pk_duration_scale")
    return out


def mask_rows_(x, lens):
    B, T = x.shape[:2]
    inner = x.numel() // (B * T)
    _lib.check(_lib.lib().pk_mask_rows(_ptr(x), _ptr(lens), B, T, inner, _stream()), "pk_mask_rows")
    return x


def variance_embed_add(hs, pitch, energy, wp, bp, we, be, lens=None):
    B, T, c = hs.shape
    hs, pitch, energy = hs.contiguous(), pitch.contiguous(), energy.contiguous()   # locals: must outlive the launch call
    y = torch.empty_like(hs)
    _lib.check(_lib.lib().pk_variance_embed_add(_ptr(hs), _ptr(pitch), _ptr(energy), _ptr(wp), _ptr(bp),
                                                wp.shape[-1], _ptr(we), _ptr(be), we.shape[-1], _ptr(lens), B, T, c, _ptr(y),
                                                _stream()), "pk_variance_embed_add")
    return y


def zscore(x, mu, sigma, inverse=False):
    x = x.contiguous().float()
    y = torch.empty_like(x)
    _lib.check(_lib.lib().pk_zscore(_ptr(x), _ptr(mu), _ptr(sigma), x.shape[-1], x.numel(), 1 if inverse else 0, _ptr(y), _stream()),
               "pk_zscore")
    return y


# ----------------------------------------------------------------------------------------------------------------
# training-step helpers (train.cu)
# ----------------------------------------------------------------------------------------------------------------
def transpose_planes(src, *, z, rows, src_zstride, ld_src, c0, cols, shift, r_out, dst, dst_zstride, ld_dst):
    """dst[z*dst_zstride + c*ld_dst + r] = src[z*src_zstride + (r+shift)*ld_src + c0 + c]; src / dst are Split (views allowed)."""
    _lib.check(_lib.lib().pk_transpose_planes(_ptr(src.hi), _ptr(src.lo), z, rows, src_zstride, ld_src, c0, cols, shift, r_out,
                                              _ptr(dst.hi), _ptr(dst.lo), dst_zstride, ld_dst, _stream()), "pk_transpose_planes")


def layer_norm_bwd(x, gamma, dy, dx, accumulate, dgamma, dbeta, eps=1e-5):
    rows, d = x.numel() // x.shape[-1], x.shape[-1]
    _lib.check(_lib.lib().pk_layer_norm_bwd(_ptr(x), _ptr(gamma), _ptr(dy), eps, rows, d, _ptr(dx), 1 if accumulate else 0,
                                            _ptr(dgamma), _ptr(dbeta), _stream()), "pk_layer_norm_bwd")


def softmax_bwd(p, dp, keys, scale):
    ld = dp.shape[-1]
    rows = dp.numel() // ld
    ds = Split.empty(tuple(dp.shape), dp.device)
    _lib.check(_lib.lib().pk_softmax_bwd(_ptr(p.hi), _ptr(p.lo), _ptr(dp), rows, keys, ld, scale, _ptr(ds.hi), _ptr(ds.lo), _stream()),
               "pk_softmax_bwd")
    return ds


def colsum_(x, out):
    c = x.shape[-1]
    _lib.check(_lib.lib().pk_colsum(_ptr(x), x.numel() // c, c, _ptr(out), _stream()), "pk_colsum")


def colsum_split_(xs, cols, out):
    """out[c] += sum over all rows of (hi + lo)[..., c], c < cols, for a contiguous Split (..., ld)."""
    ld = xs.hi.shape[-1]
    assert xs.hi.is_contiguous() and cols <= ld
    _lib.check(_lib.lib().pk_colsum_split(_ptr(xs.hi), _ptr(xs.lo), xs.hi.numel() // ld, cols, ld, _ptr(out), _stream()), "pk_colsum_split")


def sum_slices(part, out):
    """out = part.sum(0) for fp32 part (S, ...) and contiguous out (split-K reduction; overwrites out)."""
    s = part.shape[0]
    assert part.is_contiguous() and out.is_contiguous() and part.numel() == s * out.numel()
    _lib.check(_lib.lib().pk_sum_slices(_ptr(part), s, out.numel(), _ptr(out), _stream()), "pk_sum_slices")
    return out


def relu_bwd(dy, y_split, want_f32=False):
    dx = torch.empty_like(dy) if want_f32 else None
    dxs = Split.empty(tuple(dy.shape), dy.device)
    _lib.check(_lib.lib().pk_relu_bwd(_ptr(dy), _ptr(y_split.hi), dy.numel(), _ptr(dx), _ptr(dxs.hi), _ptr(dxs.lo), _stream()), "pk_relu_bwd")
    return dx, dxs


def axpy_(a, x, y):
    _lib.check(_lib.lib().pk_axpy(float(a), _ptr(x), x.numel(), _ptr(y), _stream()), "pk_axpy")


def dropout(x, p, seed, site, step, out_f32=True, out_split=False, inplace=False, step_dev=None):
    """pk_dropout: x fp32 tensor or Split (any shape, contiguous) -> (y fp32 or None, y Split or None).  p == 0 is not a
    special case here (callers skip the call)."""
    is_split = isinstance(x, Split)
    ref = x.hi if is_split else x
    assert ref.is_contiguous()
    n = ref.numel()
    y = (x if (inplace and not is_split) else torch.empty(ref.shape, dtype=torch.float32, device=ref.device)) if out_f32 else None
    ys = (x if (inplace and is_split) else Split.empty(tuple(ref.shape), ref.device)) if out_split else None
    _lib.check(_lib.lib().pk_dropout(_ptr(None if is_split else x), _ptr(x.hi if is_split else None), _ptr(x.lo if is_split else None), n,
                                     float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(site), int(step), _ptr(step_dev), _ptr(y), _ptr(ys.hi if ys else None),
                                     _ptr(ys.lo if ys else None), _stream()), "pk_dropout")
    return y, ys
