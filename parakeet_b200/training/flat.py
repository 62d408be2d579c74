"""Flat parameter / gradient buffers for the data-parallel training step.

The reference wraps the model in `paddle.DataParallel` (examples/fastspeech2/*/train.py:117-119), which all-reduces
gradients bucket by bucket.  Here every trainable tensor is a view into ONE flat fp32 buffer (and its gradient into a second
one), so the exchange step of the path is a single `all_reduce(SUM)` of the flat gradient and the optimiser is a single
kernel over the flat buffers; the 1/world scale of the DataParallel mean is applied by the optimiser.
Device-agnostic on purpose: the CPU tests run it over gloo (tests/test_dist_cpu.py).
"""
from collections import OrderedDict

import torch
import torch.distributed as dist


class FlatBuffers:
    def __init__(self, params: "OrderedDict[str, torch.Tensor]", names, device):
        """params: name -> tensor (replaced in place by views of the flat buffer for every name in `names`)."""
        sizes = [params[k].numel() for k in names]
        offs, tot = [], 0
        for s in sizes:
            offs.append(tot)
            tot += (s + 3) // 4 * 4                      # keep every view 16-byte aligned
        self.names, self.offsets, self.sizes, self.total = list(names), offs, sizes, tot
        self.flat = torch.zeros(tot, dtype=torch.float32, device=device)
        self.gflat = torch.zeros(tot, dtype=torch.float32, device=device)
        self.grads = {}
        for k, o, s in zip(names, offs, sizes):
            shape = params[k].shape
            self.flat[o:o + s].copy_(params[k].reshape(-1))
            params[k] = self.flat[o:o + s].view(shape)      # the model now reads the flat buffer
            self.grads[k] = self.gflat[o:o + s].view(shape)

    def all_reduce_grads(self, group=None):
        """The one exchange step of the path: SUM over ranks (the optimiser divides by the world size)."""
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.gflat, op=dist.ReduceOp.SUM, group=group)
