"""CPU oracle for the Parakeet TTS hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

A torch-CPU / numpy fp32 restatement of the reference functions listed in
SURVEY.md section 8a. Each function cites the reference file:line it follows.

PARITY UNPINNED: the reference (PaddlePaddle/Parakeet) is pure Python on top
of PaddlePaddle, which is absent from this image and cannot be installed;
its hot-path tests print instead of assert and hold no golden vectors
(tests/unit/test_pwg.py:72-78, tests/unit/test_stft.py:39-42).  The oracle is
therefore anchored on (a) the reference's own call sites and shapes, (b) the
Paddle-2.1 semantics check-list in oracle/README.md and (c) independent
cross-checks available here (torch.stft / numpy.fft, torch SDPA, gather-vs-
0/1-matmul, flow forward-inverse identity); see tests/test_oracle_*.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (parakeet_b200) never
does: it fails loudly when its CUDA library is missing.
"""
