"""Split-K weight gradients: dW = A^T-style NT matmuls whose reduction axis is the flattened (batch, time) axis (10^4 .. 10^6)
while the output is a handful of 128 x N tiles.  K is cut into S slices so that tiles x S fills the machine: S independent NT
matmuls (`pk_conv_gemm` batched over the slices) whose fp32 partial results are summed by `pk_sum_slices`."""
import torch

from .. import ops
from ..ops import Split


class ZeroPlanes:
    """Persistent zero-initialised split planes for the transposed GEMM operands of the weight gradients.

    `transpose_planes` rewrites the same valid region on every use of a (role, shape) key within one batch geometry and never
    touches the K padding, so the zeros are written once per geometry instead of by a fill kernel per use (~250 launches per
    FastSpeech2 step).  Buffers are filed under the current GEOMETRY (`begin(geom)`: the batch shape - it fixes every valid
    region, and a captured CUDA graph of that batch shape has the buffer addresses baked in).  At most `max_geoms` geometries
    are kept (LRU): evicting one frees its planes and calls `on_evict(geom)` so that the owner drops the graph captured for it.
    `role` keeps operands that are alive together apart."""

    def __init__(self, max_geoms=16, on_evict=None):
        self.max_geoms, self.on_evict = max_geoms, on_evict
        self._geoms = {}          # geom -> {(role, shape, device): Split}; insertion order = LRU order
        self._cur = None

    def begin(self, geom):
        """Make `geom` current (most recently used), evicting the least recently used geometries beyond the bound."""
        planes = self._geoms.pop(geom, None)
        self._geoms[geom] = planes if planes is not None else {}
        self._cur = geom
        while len(self._geoms) > self.max_geoms:
            old = next(iter(self._geoms))
            del self._geoms[old]
            if self.on_evict is not None:
                self.on_evict(old)

    touch = begin

    def get(self, role, shape, dev):
        if self._cur is None:
            self.begin(None)
        planes = self._geoms[self._cur]
        key = (role, tuple(shape), str(dev))
        buf = planes.get(key)
        if buf is None:
            buf = planes[key] = Split.zeros(tuple(shape), dev)
        return buf

    def __len__(self):
        return len(self._geoms)


_EVICT_HOOKS = []


def _default_evicted(geom):
    for hook in list(_EVICT_HOOKS):
        hook(geom)


_DEFAULT_PLANES = ZeroPlanes(max_geoms=16, on_evict=_default_evicted)   # one PWG step uses 2-3 geometries (sample rate, frame rate)


def on_default_evict(hook):
    """Register hook(geom) for evictions from the module-level cache (a training step that captures CUDA graphs over these planes
    drops its graphs there: their buffer addresses are baked in)."""
    _EVICT_HOOKS.append(hook)


def zero_planes(role, shape, dev, geom=None):
    """Module-level cache for callers without their own ZeroPlanes (the PWG training step: fixed batch geometry)."""
    if geom is not None and geom != _DEFAULT_PLANES._cur:
        _DEFAULT_PLANES.begin(geom)
    return _DEFAULT_PLANES.get(role, shape, dev)


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
