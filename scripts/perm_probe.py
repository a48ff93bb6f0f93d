"""Permuting sweeps on the GPU: QFT(n) with its closing bit reversal riding on the last sweep against the two-launch form
(sweeps + K8), and the bare permutation through the sweep kernel against K8.

    python scripts/perm_probe.py [n] [dtype]      # prints one JSON line
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qibo_b200 import circuits  # noqa: E402
from qibo_b200.engine import Engine, resolve_spans, swaps_for_permutation  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dtype = sys.argv[2] if len(sys.argv) > 2 else "complex128"
eng = Engine(0)
st = eng.basis_state(n, dtype)
alt = eng.empty((1 << n,), dtype)
out = {"n": n, "dtype": dtype}


def timed_run(prog, reps=5):
    best = None
    for _ in range(reps):
        spans = []
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        stats = eng.run_program(prog, st, alt=alt, spans=spans)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        if best is None or ms < best[0]:
            best = (ms, {k: (round(v[0], 3), v[1]) for k, v in resolve_spans(spans).items()}, stats.nsweeps, stats.nperm, stats.nperm_fused)
    return {"ms": round(best[0], 3), "spans": best[1], "nsweeps": best[2], "nperm": best[3], "nperm_fused": best[4]}


cases = {"qft": circuits.qft(n), "bit_reversal_only": swaps_for_permutation([n - 1 - q for q in range(n)])}
k = 3
tail = circuits.qft(k, with_swaps=False)  # the stages of QFT(3) on the leading qubits
cases["sharded_tail"] = tail + swaps_for_permutation(list(range(k)) + [n - 1 - q for q in range(n - k)])
for name, ops in cases.items():
    for fused in (True, False):
        eng.fuse_permutations = fused
        prog = eng.compile(n, dtype, ops)
        timed_run(prog, 1)
        out[f"{name}_{'fused' if fused else 'two_launch'}"] = timed_run(prog)
        prog.close()
print(json.dumps(out), flush=True)
