"""Golden vectors produced by RUNNING the reference's own code - the few functions on the path that need numpy only
(everything else imports paddle, which cannot be installed here).  Loaded by file path so that `parakeet/__init__.py` (which
imports paddle) is never executed.  Run in the build container (needs /root/reference); the .npz travels with the repo.

    python scripts/make_golden_ref.py
"""
import importlib.util
import os

import numpy as np

REF = "/root/reference/parakeet"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_batch_sequences.npz")


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    batch = load(os.path.join(REF, "data", "batch.py"), "ref_batch")
    rng = np.random.RandomState(20260923)
    out = {}
    lengths = [5, 11, 3, 8]
    text = [rng.randint(1, 70, size=n).astype(np.int64) for n in lengths]
    speech = [rng.randn(3 * n, 7).astype(np.float32) for n in lengths]
    pitch = [rng.randn(n, 1).astype(np.float32) for n in lengths]
    for name, seqs in (("text", text), ("speech", speech), ("pitch", pitch)):
        for i, s in enumerate(seqs):
            out[f"{name}_in{i}"] = s
        out[f"{name}_out"] = batch.batch_sequences(seqs)              # parakeet/data/batch.py:170-189, executed
    out["n"] = np.asarray(len(lengths))
    np.savez(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k.endswith("_out")})


if __name__ == "__main__":
    main()
