"""The last (contiguous-tile, compute-heavy) sweep of QFT(n) alone, for an ncu capture:  python scripts/ncu_last_sweep.py 31"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qibo_b200 import circuits  # noqa: E402
from qibo_b200.engine import Engine, plan_program  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 31
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
eng = Engine(0)
ops = circuits.qft(n, with_swaps=False)
_, sweep_of_op = plan_program(n, "complex128", ops)
last = sorted(set(sweep_of_op))[which]
sub = [op for op, s in zip(ops, sweep_of_op) if s == last]
st = eng.basis_state(n, "complex128")
for _ in range(3):
    stats = eng.apply_program(st, n, sub, timed=True)
print(len(sub), stats.nsweeps, stats.elapsed_ms)
