"""Split-K weight gradients: dW = A^T-style NT matmuls whose reduction axis is the flattened (batch, time) axis (10^4 .. 10^6)
while the output is a handful of 128 x N tiles.  K is cut into S slices so that tiles x S fills the machine: S independent NT
matmuls (`pk_conv_gemm` batched over the slices) whose fp32 partial results are summed by `pk_sum_slices`."""
import torch

from .. import ops
from ..ops import Split


_ZERO_PLANES = {}


def zero_planes(role, shape, dev):
    """Persistent zero-initialised split planes for a transposed GEMM operand.  `transpose_planes` rewrites the same valid region
    on every use of a (role, shape) key and never touches the K padding, so the zeros are written once, not by a fill kernel per
    use.  `role` carries whatever fixes the valid region (batch, time, ...) and keeps operands that are alive together apart."""
    key = (role, tuple(shape), str(dev))
    buf = _ZERO_PLANES.get(key)
    if buf is None:
        buf = _ZERO_PLANES[key] = Split.zeros(tuple(shape), dev)
    return buf


def plan(batch, t, m, n, max_slices=128):
    """-> (Tp, S, ks, KKp): padded time, number of K slices, slice length (multiple of 64), padded reduction length S * ks."""
    tp = (t + 63) // 64 * 64
    kk = batch * tp
    tiles = ((m + 127) // 128) * ((n + 127) // 128)
    s = max(1, min(max_slices, kk // 512, -(-296 // tiles)))          # ~2 tiles per SM
    ks = ((kk + s - 1) // s + 63) // 64 * 64
    return tp, s, ks, s * ks


def nt_splitk(at, bt, m, n, s, ks, kkp, out=None):
    """at Split (m rows, kkp), bt Split (n rows, kkp) -> (m, n) fp32 = at . bt^T, reduced over the S slices."""
    dev = at.hi.device
    sa = dict(rows=m, cols=ks, ld=kkp, batch_stride=ks, batches=s, bmul=1, hmul=0, col0=0, colh=0)
    sb = dict(rows=n, cols=ks, ld=kkp, batch_stride=ks, batches=s, bmul=1, hmul=0, col0=0, colh=0)
    direct = out is not None and out.is_contiguous()
    if s == 1:
        y = out if direct else torch.empty(m, n, dtype=torch.float32, device=dev)
        ops.batched_matmul_nt(at, bt, batch=1, heads=1, m=m, n=n, k=ks, a_spec=sa, b_spec=sb, y_f32=y, y_batch_stride=0, y_head_stride=0, y_ld=n)
        if out is not None and not direct:
            out.copy_(y)
            return out
        return y
    part = torch.empty(s, m, n, dtype=torch.float32, device=dev)
    ops.batched_matmul_nt(at, bt, batch=s, heads=1, m=m, n=n, k=ks, a_spec=sa, b_spec=sb, y_f32=part, y_batch_stride=m * n, y_head_stride=0,
                          y_ld=n)
    y = out if direct else torch.empty(m, n, dtype=torch.float32, device=dev)
    ops.sum_slices(part, y)
    if out is not None and y is not out:
        out.copy_(y)
        return out
    return y
