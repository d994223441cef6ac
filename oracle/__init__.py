"""CPU oracle for the speechcatcher streaming decode path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch fp32 on CPU + numpy fp64 where the
reference uses them) of the reference algorithm for the hot path named in
BASELINE.json: frontend -> contextual-block streaming encoder -> block-synchronous
beam search (decoder step + CTC prefix scorer + top-k) -> hypothesis bookkeeping.
Every function cites the reference file:line it follows (paths are relative to
the reference checkout, speechcatcher/...).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it, and only as the checker or the timed CPU
baseline -- never as part of the shipped CUDA path (`speechcatcher_b200/`),
which must fail loudly if its CUDA extension is missing.

Parity pinning: the reference's own tests hold no golden vectors for this path
(SURVEY.md 8(c)), so the oracle is pinned against outputs of the reference
itself, generated in the build container by `oracle/gen_golden.py` (which imports
the Python reference from /root/reference) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` replays them.
"""
