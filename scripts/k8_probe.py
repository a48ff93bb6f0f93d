import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from qibo_b200.engine import Engine
eng = Engine(0)
for n, dt in ((30, "complex128"), (31, "complex64")):
    st = eng.basis_state(n, dt)
    perm = [n - 1 - q for q in range(n)]
    for _ in range(2): eng.permute_qubits(st, n, perm)
    ts = [eng.permute_qubits(st, n, perm, timed=True) for _ in range(5)]
    B = 16 if dt == "complex128" else 8
    print(n, dt, "bit reversal ms", min(ts), "GB/s", 2 * B * 2.0**n / min(ts) / 1e6)
    rot = [(q + 7) % n for q in range(n)]
    ts = [eng.permute_qubits(st, n, rot, timed=True) for _ in range(3)]
    print(n, dt, "rotation ms", min(ts), "GB/s", 2 * B * 2.0**n / min(ts) / 1e6)
