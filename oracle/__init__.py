"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the qibo-b200 state-vector hot path.

A restatement (not a copy) of the algorithm the reference `NumpyBackend` runs for this path
(`/root/reference/src/qibo/backends/abstract.py`, `einsum_utils.py`, `npmatrices.py`).  Parity is
PINNED: `tests/golden/*.npz` were produced by importing the unmodified reference in the build
container (`tests/golden/make_golden.py`) and `tests/test_oracle_golden.py` checks every oracle
function against them; when the reference is importable (`baseline/_ref`), `tests/test_oracle_vs_ref.py`
also compares live.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import this package.  Nothing under `qibo_b200/` does -- the product path has no CPU fallback.
"""
