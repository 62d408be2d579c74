"""The reference's dumped-feature dataset layout, collate function and per-rank sampler (parakeet_b200/data.py)."""
import json

import numpy as np
import torch

from parakeet_b200 import data


def _dump(tmp_path, lengths, n_mels=80, seed=0):
    rng = np.random.RandomState(seed)
    (tmp_path / "data_speech").mkdir(); (tmp_path / "data_pitch").mkdir(); (tmp_path / "data_energy").mkdir()
    records = []
    for i, t in enumerate(lengths):
        dur = rng.randint(1, 6, size=t)
        L = int(dur.sum())
        np.save(tmp_path / "data_speech" / f"u{i}_speech.npy", rng.randn(L, n_mels).astype(np.float32))
        np.save(tmp_path / "data_pitch" / f"u{i}_pitch.npy", rng.randn(t).astype(np.float32))
        np.save(tmp_path / "data_energy" / f"u{i}_energy.npy", rng.randn(t, 1).astype(np.float32))
        records.append({"utt_id": f"u{i}", "text": rng.randint(1, 60, size=t).tolist(), "text_lengths": t, "speech_lengths": L,
                        "durations": dur.tolist(), "speech": f"data_speech/u{i}_speech.npy", "pitch": f"data_pitch/u{i}_pitch.npy",
                        "energy": f"data_energy/u{i}_energy.npy"})
    with open(tmp_path / "metadata.jsonl", "wt") as f:
        for r in records:
            f.write(json.dumps(r) + "\n")
    return records


def test_metadata_table_and_collate_match_the_reference_layout(tmp_path):
    lengths = [5, 9, 3]
    recs = _dump(tmp_path, lengths)
    meta = data.read_metadata(tmp_path / "metadata.jsonl")
    assert [m["utt_id"] for m in meta] == ["u0", "u1", "u2"]
    table = data.FeatureTable(meta, root=str(tmp_path))
    assert len(table) == 3 and set(table[0]) == set(data.FS2_FIELDS)
    batch = data.fastspeech2_batch([table[i] for i in range(3)])
    Lmax = max(r["speech_lengths"] for r in recs)
    assert batch["text"].shape == (3, 9) and batch["text"].dtype == torch.int64
    assert batch["speech"].shape == (3, Lmax, 80) and batch["speech"].dtype == torch.float32
    assert batch["pitch"].shape == (3, 9, 1) and batch["energy"].shape == (3, 9, 1)
    assert batch["text_lengths"].tolist() == lengths and batch["speech_lengths"].tolist() == [r["speech_lengths"] for r in recs]
    for i, r in enumerate(recs):                                      # zero padding past each example, values untouched before it
        t, L = r["text_lengths"], r["speech_lengths"]
        assert batch["text"][i, :t].tolist() == r["text"] and batch["text"][i, t:].abs().sum() == 0
        assert batch["durations"][i].sum().item() == L
        assert np.array_equal(batch["speech"][i, :L].numpy(), np.load(tmp_path / r["speech"])) and batch["speech"][i, L:].abs().sum() == 0
        assert batch["pitch"][i, t:].abs().sum() == 0 and batch["energy"][i, t:].abs().sum() == 0
    # the oracle's loss / forward accept the batch as is (the same dict the updater passes to the model)
    from oracle import fastspeech2 as ofs
    p = ofs.synth_params(1)
    out = ofs.fs2_forward(p, None, batch["text"], batch["text_lengths"], batch["speech_lengths"], batch["durations"], batch["pitch"],
                          batch["energy"])
    assert out[1].shape == (3, Lmax, 80)


def test_distributed_batch_sampler_contract():
    n, bs = 53, 4
    for nranks in (1, 2, 4):
        per_rank = []
        for r in range(nranks):
            s = data.DistributedBatchSampler(n, bs, nranks, r, shuffle=True, drop_last=True)
            s.set_epoch(3)
            batches = list(s)
            assert all(len(b) == bs for b in batches) and len(batches) == len(s)
            per_rank.append([i for b in batches for i in b])
        assert len({len(p) for p in per_rank}) == 1                   # every rank runs the same number of steps
        flat = [i for p in per_rank for i in p]
        assert len(flat) - len(set(flat)) <= nranks                   # disjoint up to the wrap-around pad
        assert set(flat) <= set(range(n)) and len(set(flat)) >= n - nranks * bs
    a = data.DistributedBatchSampler(n, bs, 2, 0, shuffle=True); a.set_epoch(0)
    b = data.DistributedBatchSampler(n, bs, 2, 0, shuffle=True); b.set_epoch(1)
    assert list(a) != list(b)                                         # reshuffled per epoch
    c = data.DistributedBatchSampler(10, 4, 1, 0, shuffle=False, drop_last=False)
    assert list(c) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]] and len(c) == 3


def test_batch_sequences_against_vectors_produced_by_the_reference_code():
    """tests/golden/ref_executed.npz was written by scripts/make_golden_ref.py, which EXECUTES the reference's own
    parakeet/data/batch.py:170-189 (numpy only)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_executed.npz"))
    n = int(g["n"])
    for name in ("text", "speech", "pitch"):
        seqs = [g[f"{name}_in{i}"] for i in range(n)]
        got = data.batch_sequences(seqs)
        assert got.dtype == g[f"{name}_out"].dtype and np.array_equal(got, g[f"{name}_out"]), name
