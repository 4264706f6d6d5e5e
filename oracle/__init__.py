"""CPU oracle for the LiDiff denoising hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch-CPU / numpy, fp32 with an fp64 switch) of
the algorithm the reference executes on its hot path
(`lidiff/tools/diff_completion_pipeline.py:117-169`, `lidiff/models/minkunet.py`,
`lidiff/models/models.py:98-178`) *including* the third-party operators it calls
(MinkowskiEngine 0.5.4, pykeops 2.1.2 `argKmin`, diffusers 0.18.0
`DPMSolverMultistepScheduler`), none of which is vendored in `/root/reference` nor
installable in this environment (no network, no wheels).

PARITY UNPINNED: the reference ships no tests, golden vectors or expected outputs for
this path and its arithmetic lives in the un-vendored dependencies above, so this
oracle cannot be checked against the real libraries here.  What pins it instead:
  * closed-form known-answer values (timestep tables, DPM-Solver++ coefficients, the
    sinusoidal time embedding, round-half-even) in `tests/golden/known_answers.json`,
    produced by `tests/golden/make_known_answers.py`;
  * structural invariants of sparse convolution (dense-conv equivalence on a filled
    grid, transposed-map symmetry) checked in `tests/test_oracle.py`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this package; the product (`lidiff_b200/`) never
does, and fails loudly when its CUDA library is missing.
"""
