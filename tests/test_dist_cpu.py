"""N > 1 path on CPU: world_size-2 gloo processes shard a ragged batch, 'synthesise' their slice and gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parakeet_b200.parallel import gather_variable, pad_batch, shard_indices


def test_shard_indices_partition_and_balance():
    lengths = [60, 140, 83, 71, 100, 97, 133, 65, 120]
    for world in (1, 2, 4, 8):
        parts = [shard_indices(lengths, world, r) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(lengths)))                       # a partition: every utterance exactly once
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
        if world == 2:
            tot = [sum(lengths[i] for i in p) for p in parts]
            assert abs(tot[0] - tot[1]) <= max(lengths)


def test_pad_batch():
    seqs = [torch.arange(3), torch.arange(5), torch.arange(1)]
    x, lens = pad_batch(seqs)
    assert x.tolist() == [[0, 1, 2, 0, 0], [0, 1, 2, 3, 4], [0, 0, 0, 0, 0]] and lens.tolist() == [3, 5, 1]


def _worker(rank, world, port, lengths, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_indices(lengths, world, rank)
        # stand-in for FastSpeech2 -> PWG on this rank's slice: a deterministic function of (index, length)
        results = [torch.full((lengths[i] * 3,), float(i)) for i in mine]
        full = gather_variable(results, mine, len(lengths))
        ok = all(full[i].shape[0] == lengths[i] * 3 and float(full[i][0]) == float(i) for i in range(len(lengths)))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, ok, float(t.item()), mine))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    lengths = [60, 140, 83, 71, 100, 97, 133]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in out)
    assert all(abs(tmax - 11.0) < 1e-9 for _, _, tmax, _ in out)
    assert sorted(i for _, _, _, mine in out for i in mine) == list(range(len(lengths)))
