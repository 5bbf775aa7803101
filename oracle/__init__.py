"""CPU oracle for the edge+ESF-Net hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``egn_b200`` + ``libegn.so``) never does, and fails loudly without its CUDA
library.

Parity status: the reference repository has no tests, golden vectors or
known-answer values for this path, so parity is pinned
against *outputs of the reference itself run in the build container*:
``oracle/make_golden.py`` imports the unmodified reference modules from
``/root/reference`` (under the shims in ``oracle/ref_harness.py``), feeds them
the seeded synthetic checkpoints of ``oracle/synth.py`` and commits the results
under ``tests/golden/``.  ``oracle/graph.py`` (a torch-fp32 functional
restatement, every function citing the reference file:line it follows) is then
checked against those fixtures by ``tests/test_oracle_golden.py``.
"""
