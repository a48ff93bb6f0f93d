"""complex64 tolerance probe: for the three tests that used a loosened bound, the error of (a) the CUDA path and (b) the
reference algorithm (oracle) in complex64, both against the oracle run in complex128 on the same input."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import oracle_run, rand_state, rand_unitary, random_zoo  # noqa: E402
from oracle import numpy_oracle as orc  # noqa: E402
from qibo_b200.engine import Engine  # noqa: E402
from qibo_b200.ops import Op  # noqa: E402

eng = Engine(0)
rows = []


def probe(tag, n, ops, psi64):
    exact = oracle_run(psi64.astype(np.complex128), ops, n)
    ref64 = oracle_run(psi64, ops, n)
    st = eng.upload(psi64)
    eng.apply_program(st, n, ops)
    k2 = st.numpy()
    st = eng.upload(psi64)
    for op in ops:
        eng.apply_op(st, n, op)
    k1 = st.numpy()
    rows.append({"case": tag, "oracle64_vs_exact": float(np.abs(ref64 - exact).max()), "k2_vs_exact": float(np.abs(k2 - exact).max()),
                 "k1_vs_exact": float(np.abs(k1 - exact).max()), "k2_vs_oracle64": float(np.abs(k2 - ref64).max()),
                 "k1_vs_oracle64": float(np.abs(k1 - ref64).max()), "max_abs_amp": float(np.abs(exact).max())})
    print(rows[-1], flush=True)


for seed in range(4):
    probe(f"zoo16_seed{seed}", 16, random_zoo(16, 60, seed), rand_state(16, seed, "complex64"))
rng = np.random.default_rng(6)
ops = [Op(rand_unitary(6, rng), (13, 2, 7, 0, 9, 4)), Op(orc.gate_matrix("H"), (3,)), Op(rand_unitary(6, rng), (1, 5, 3, 8, 12, 6), (10,))]
probe("six_qubit_dense", 15, ops, rand_state(15, 3, "complex64"))
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r2a_tol_probe.json"), "w"), indent=1)
