"""Runs one named program a few times (for ncu):  python scripts/prof_case.py <case> [n] [dtype]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qibo_b200 import circuits  # noqa: E402
from qibo_b200.engine import Engine  # noqa: E402
from qibo_b200.ops import Op  # noqa: E402

case = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
dtype = sys.argv[3] if len(sys.argv) > 3 else "complex128"
eng = Engine(0)
st = eng.basis_state(n, dtype)
h = circuits.matrix("H")
if case == "h24":
    ops = [Op(h, (n - 1 - (3 + i % 3),)) for i in range(24)]
elif case == "h1":
    ops = [Op(h, (0,))]
elif case == "qft":
    ops = circuits.qft(n, with_swaps=False)
elif case == "qftswap":
    ops = circuits.qft(n)
elif case == "var":
    ops = circuits.variational(n, 2, np.random.default_rng(7).random(4 * n) * 6.28)
else:
    raise SystemExit("unknown case")
reps = int(os.environ.get("REPS", 2))
for _ in range(reps):
    stats = eng.apply_program(st, n, ops, timed=True)
    print(case, n, dtype, "sweeps", stats.nsweeps, "ms", stats.elapsed_ms)
torch.cuda.synchronize()
