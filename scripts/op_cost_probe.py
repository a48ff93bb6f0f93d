"""Cost model of the general sweep kernel: marginal time per micro-op kind and per pass (one sweep, n qubits), measured by
growing one sweep's program.  python scripts/op_cost_probe.py [n] [dtype]  -> one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qibo_b200 import circuits  # noqa: E402
from qibo_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dtype = sys.argv[2] if len(sys.argv) > 2 else "complex64"
eng = Engine(0)
st = eng.filled_state(n, 2.0 ** (-n / 2), dtype)
rng = np.random.default_rng(3)
q = lambda bit: n - 1 - bit  # noqa: E731
ry = lambda bit: circuits.op("RY", (q(bit),), float(rng.uniform(0.1, 3)))  # noqa: E731
rx = lambda bit: circuits.op("RX", (q(bit),), float(rng.uniform(0.1, 3)))  # noqa: E731
cz = lambda a, b: circuits.op("CZ", (q(a), q(b)))  # noqa: E731
cu1 = lambda a, b: circuits.op("CU1", (q(a), q(b)), float(rng.uniform(0.1, 3)))  # noqa: E731
cnot = lambda a, b: circuits.op("CNOT", (q(a), q(b)))  # noqa: E731
H = [20, 21, 22, 23, 24, 25, 26]  # high tile bits


def run(ops):
    best, s = 1e9, None
    for _ in range(4):
        s = eng.apply_program(st, n, ops, timed=True)
        best = min(best, s.elapsed_ms)
    return {"ms": round(best, 3), "sweeps": s.nsweeps, "passes": s.ndense_passes, "ops": len(ops)}


out = {"n": n, "dtype": dtype, "bytes_per_sweep": 2 * (16 if dtype == "complex128" else 8) * 2.0**n}
for k in (1, 5, 9, 17):
    out[f"ry_same_bit_x{k}"] = run([ry(23) for _ in range(k)])
    out[f"rx_same_bit_x{k}"] = run([rx(23) for _ in range(k)])
    out[f"ry_layer4_x{k}"] = run([ry(b) for _ in range(k) for b in H[:4]])
    out[f"ry_cz_regbits_x{k}"] = run([g for _ in range(k) for g in (ry(23), cz(23, 24))])
    out[f"ry_cz_lowbit_x{k}"] = run([g for _ in range(k) for g in (ry(23), cz(23, 2))])
    out[f"ry_cz_outside_x{k}"] = run([g for _ in range(k) for g in (ry(23), cz(23, 15))])
    out[f"ry_cu1_regbits_x{k}"] = run([g for _ in range(k) for g in (ry(23), cu1(23, 24))])
    out[f"ry_cnot_regbits_x{k}"] = run([g for _ in range(k) for g in (ry(23), cnot(23, 24))])
for k in (1, 2, 3):
    # k passes of one 4-gate layer each: RYs on 4k distinct tile bits (high bits first, then low ones)
    bits = (H + [0, 1, 2, 3, 4])[: 4 * k]
    out[f"passes_x{k}"] = run([ry(b) for b in bits])
os.environ["QB_SWEEP_SKIP_COMPUTE"] = "1"
out["movement_only"] = run([ry(b) for b in H[:4]])
print(json.dumps(out), flush=True)
