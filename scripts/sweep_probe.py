"""Per-sweep timings of QFT(n) (each planner sweep as its own timed program), with and without the tile arithmetic
(QB_SWEEP_SKIP_COMPUTE=1: the data movement of each tile shape alone), plus K8 permutation variants.

    python scripts/sweep_probe.py [n]        # prints one JSON line
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qibo_b200 import circuits  # noqa: E402
from qibo_b200.engine import Engine, plan_program  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dtype = sys.argv[2] if len(sys.argv) > 2 else "complex128"
B = 16 if dtype == "complex128" else 8
eng = Engine(0)
ops = circuits.qft(n, with_swaps=False)
_, sweep_of_op = plan_program(n, dtype, ops)
groups = {}
for i, s in enumerate(sweep_of_op):
    groups.setdefault(int(s), []).append(ops[i])
st = eng.basis_state(n, dtype)
out = {"n": n, "dtype": dtype, "bytes_per_sweep": 2 * B * 2.0**n}
for skip in ("0", "1"):
    os.environ["QB_SWEEP_SKIP_COMPUTE"] = skip
    rows = []
    for k in sorted(groups):
        best = 1e9
        for _ in range(4):
            s = eng.apply_program(st, n, groups[k], timed=True)
            best = min(best, s.elapsed_ms)
        rows.append((k, len(groups[k]), round(best, 3), round(2 * B * 2.0**n / best / 1e6, 1)))
    out["compute" if skip == "0" else "movement_only"] = rows
os.environ["QB_SWEEP_SKIP_COMPUTE"] = "0"
perm = [n - 1 - q for q in range(n)]
for low in ("6", "5", "4"):
    for ctas in ("0", "2", "4", "6"):
        os.environ["QB_PERM_LOW_BITS"] = low
        if ctas == "0":
            os.environ.pop("QB_PERM_CTAS_PER_SM", None)
        else:
            os.environ["QB_PERM_CTAS_PER_SM"] = ctas
        try:
            for _ in range(2):
                eng.permute_qubits(st, n, perm)
            ts = [eng.permute_qubits(st, n, perm, timed=True) for _ in range(4)]
            out[f"k8_low{low}_ctas{ctas}"] = round(min(ts), 3)
        except Exception as exc:  # noqa: BLE001
            out[f"k8_low{low}_ctas{ctas}"] = repr(exc)[:80]
print(json.dumps(out), flush=True)
