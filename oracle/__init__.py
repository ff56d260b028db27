"""CPU oracle for the entailment-cone hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference
(ankitdhall/learning_embeddings) for the one hot path this repository
accelerates.  It is the checker: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  The product
package `learning_embeddings_b200` never imports it and has no CPU fallback.

Parity status: PINNED.  The reference is a PyTorch program that imports and runs
in the build container, so every function here is checked against golden vectors
produced by the unmodified reference (`tests/golden/make_golden.py` ->
`tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`).  The reference
itself ships no tests, fixtures or known-answer vectors.

Modules
  cones.py    energies, hinges, loss assembly, row transforms, RSGD step, scoring,
              threshold sweep (torch on CPU, dtype-generic: run in float32 for
              like-for-like and float64 for ground truth) + closed-form gradients
  sampler.py  CPython `random.seed` / `random.choice` (MT19937) restated, and the
              negative-edge sampler on dense and compressed adjacency
  cone_oracle.c  the pair energy + gradient, RSGD row update and scoring restated
              in plain C (OpenMP) for full-size checks and the CPU baseline
"""
