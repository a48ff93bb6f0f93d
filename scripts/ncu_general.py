"""BASELINE configs 3 / 4 through the general sweep kernel, for ncu captures and per-sweep CUDA-event timings.

    python scripts/ncu_general.py c3 30 [--per-sweep]      # RY/CZ ansatz, complex64
    python scripts/ncu_general.py c4 28 [--per-sweep]      # random circuit, complex128
    python scripts/ncu_general.py c3f 30                   # the ansatz as the reference fuser's 2-qubit blocks (dense 4x4)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qibo_b200 import circuits  # noqa: E402
from qibo_b200.engine import Engine, plan_program  # noqa: E402
from qibo_b200.ops import Op  # noqa: E402


def fused_pairs(n, layers, thetas):
    """What circuit.fuse(max_qubits=2) makes of the ansatz: RY,RY,CZ on a pair -> one dense 4x4 (real) block."""
    th = iter(thetas)
    ops = []
    ry = lambda t: circuits.matrix("RY", t)  # noqa: E731
    cz = circuits.matrix("CZ")
    for _ in range(layers):
        a = [ry(float(next(th))) for _ in range(n)]
        for i in range(0, n - 1, 2):
            ops.append(Op(cz @ np.kron(a[i], a[i + 1]), (i, i + 1)))
        b = [ry(float(next(th))) for _ in range(n)]
        ops.append(Op(b[0], (0,)))
        for i in range(1, n - 2, 2):
            ops.append(Op(cz @ np.kron(b[i], b[i + 1]), (i, i + 1)))
        ops.append(Op(b[n - 1], (n - 1,)))
        ops.append(circuits.op("CZ", (0, n - 1)))
    return ops


def main():
    case, n = sys.argv[1], int(sys.argv[2])
    per_sweep = "--per-sweep" in sys.argv
    layers = int(os.environ.get("LAYERS", 20))
    eng = Engine(0)
    if case.startswith("c3"):
        dtype, B = "complex64", 8
        thetas = 2 * np.pi * np.random.default_rng(7).random(2 * layers * n)
        ops = fused_pairs(n, layers, thetas) if case == "c3f" else circuits.variational(n, layers, thetas)
    else:
        dtype, B = "complex128", 16
        ops = circuits.random_circuit(n, int(os.environ.get("NGATES", 300)), seed=11)
    st = eng.basis_state(n, dtype)
    stats = eng.apply_program(st, n, ops, timed=True)
    stats = eng.apply_program(st, n, ops, timed=True)
    torch.cuda.synchronize()
    out = {"case": case, "n": n, "dtype": dtype, "gates": len(ops), "sweeps": stats.nsweeps, "passes": stats.ndense_passes,
           "ms": stats.elapsed_ms, "ms_per_sweep": stats.elapsed_ms / max(stats.nsweeps, 1),
           "GBps_per_sweep": 2 * B * 2.0**n * stats.nsweeps / (stats.elapsed_ms * 1e-3) / 1e9, "norm2": eng.norm2(st)}
    if per_sweep:
        _, sweep_of_op = plan_program(n, dtype, ops)
        groups = {}
        for i, s in enumerate(sweep_of_op):
            groups.setdefault(int(s), []).append(ops[i])
        times = []
        for k in sorted(groups):
            if k < 0:
                continue
            best = 1e9
            for _ in range(3):
                s = eng.apply_program(st, n, groups[k], timed=True)
                best = min(best, s.elapsed_ms)
            times.append((k, len(groups[k]), s.nsweeps, s.ndense_passes, round(best, 3)))
        out["per_sweep(index, ops, sweeps, passes, ms)"] = times
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
