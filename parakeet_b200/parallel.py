"""Batch sharding for multi-GPU synthesis: utterances are independent, so inference shards as
contiguous slices of a length-sorted batch - one process per GPU, weights replicated, NO data-path collective.
The only communication is the host-side gather of results (variable-length tensors) through torch.distributed.

Mirrors what the reference gets from `DistributedBatchSampler` for training (examples/fastspeech2/train.py:101-105);
for inference the reference simply loops over utterances on one device (synthesize.py:96-104).
"""
from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_indices(lengths: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Indices of the utterances rank `rank` processes.  Utterances are sorted by length (longest first) and dealt in a
    boustrophedon (snake) order so that every rank receives the same count (+-1) and a similar total length."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    mine = []
    for pos, idx in enumerate(order):
        rnd, slot = divmod(pos, world_size)
        owner = slot if rnd % 2 == 0 else world_size - 1 - slot
        if owner == rank:
            mine.append(idx)
    return mine


def gather_variable(tensors: List[torch.Tensor], indices: List[int], total: int, group=None) -> List[torch.Tensor]:
    """All ranks receive the full list of per-utterance results (CPU tensors), ordered by the original utterance index.
    Uses all_gather_object: results are variable-length, small, and host-side (wav / mel files to be written)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    payload = [(int(i), t.detach().cpu()) for i, t in zip(indices, tensors)]
    if world == 1:
        gathered = [payload]
    else:
        gathered = [None] * world
        dist.all_gather_object(gathered, payload, group=group)
    out = [None] * total
    for part in gathered:
        for i, t in part:
            out[i] = t
    missing = [i for i, t in enumerate(out) if t is None]
    if missing:
        raise RuntimeError(f"utterances {missing} were not produced by any rank")
    return out


def pad_batch(seqs: List[torch.Tensor], pad_value=0):
    """(T_i, ...) tensors -> padded (B, Tmax, ...) + lengths (data/batch.py:170-189 `batch_sequences` semantics)."""
    lens = torch.tensor([s.shape[0] for s in seqs], dtype=torch.int64)
    out = torch.full((len(seqs), int(lens.max())) + tuple(seqs[0].shape[1:]), pad_value, dtype=seqs[0].dtype)
    for i, s in enumerate(seqs):
        out[i, :s.shape[0]] = s
    return out, lens
