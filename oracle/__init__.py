"""CPU oracle for the APTAI hot path — TEST INFRASTRUCTURE ONLY.

Plain PyTorch-CPU fp32 (floating-point stages) and NumPy (integer / dynamic-programming stages) restatements of
the arithmetic the reference executes on its hot path.  Every function cites the reference file:line (or the
`HF:` line of transformers 5.5.0 modeling_wav2vec2.py, the un-vendored dependency that holds the encoder
arithmetic) it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
package.  Nothing under `aptai_b200/` imports it; the product path has no CPU fallback.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against
outputs of the reference's own classes run in the build container (`tests/golden/make_golden.py` imports
/root/reference + transformers 5.5.0 and commits small fixtures under tests/golden/), and the Viterbi rule against
`torchaudio.functional.forced_align`.
"""
