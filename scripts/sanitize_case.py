"""A few small programs through every new code path (permuting sweeps, out-of-place programs, INPUT_ZERO, the general and
the stage-only kernel), once each, for compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import oracle_run, rand_state, random_zoo  # noqa: E402
from qibo_b200 import circuits  # noqa: E402
from qibo_b200.engine import Engine, swaps_for_permutation  # noqa: E402

eng = Engine(0)
worst = 0.0
for dtype in ("complex128", "complex64"):
    for n in (14, 16):
        cases = {
            "qft+reversal": circuits.qft(n),
            "zoo+perm": random_zoo(n, 20, 1) + swaps_for_permutation(np.random.default_rng(2).permutation(n).tolist()),
            "reversal only": swaps_for_permutation(list(range(n - 1, -1, -1))),
            "variational": circuits.variational(n, 2, np.random.default_rng(7).random(4 * n) * 6.28),
        }
        for name, ops in cases.items():
            psi = rand_state(n, 3, dtype)
            ref = oracle_run(psi, ops, n)
            st = eng.upload(psi)
            eng.apply_program(st, n, ops)
            worst = max(worst, float(np.abs(st.numpy() - ref).max()))
            prog = eng.compile(n, dtype, ops)
            a, b = eng.upload(psi), eng.empty((1 << n,), dtype)
            eng.run_program(prog, a, alt=b)
            worst = max(worst, float(np.abs(a.numpy() - ref).max()))
            z = eng.uninitialised_state(n, dtype)
            eng.run_program(prog, z, input_zero=True)
            zref = oracle_run(np.eye(1, 1 << n, 0, dtype=dtype)[0], ops, n)
            worst = max(worst, float(np.abs(z.numpy() - zref).max()))
        gates_only = circuits.qft(n, with_swaps=False)
        cp = eng.compile_copying(n, dtype, gates_only)
        src, dst = eng.upload(psi), eng.empty((1 << n,), dtype)
        eng.run_copying(cp, src, dst)
        worst = max(worst, float(np.abs(dst.numpy() - oracle_run(psi, gates_only, n)).max()))
print("sanitize cases ok, worst error", worst, flush=True)
assert worst < 1e-5
