"""BASELINE.json configs 3 and 4 on one B200 (timings for profiles/):  python scripts/configs_bench.py [--small]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qibo_b200 import circuits  # noqa: E402
from qibo_b200.engine import Engine  # noqa: E402

small = "--small" in sys.argv
eng = Engine(0)
out = {}


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


# ---- config 3: 32-qubit RY/CZ ansatz, depth 20, complex64
n, layers = (26, 4) if small else (32, 20)
thetas = 2 * np.pi * np.random.default_rng(7).random(2 * layers * n)
ops = circuits.variational(n, layers, thetas)
st = eng.basis_state(n, "complex64")
t, stats = timed(lambda: eng.apply_program(st, n, ops))
out["config3_variational"] = {"nqubits": n, "layers": layers, "dtype": "complex64", "gates": len(ops), "sweeps": stats.nsweeps,
                              "seconds": t, "gates_per_s": len(ops) / t, "GBs_per_sweep": stats.nsweeps * 2 * 8 * 2.0**n / t / 1e9,
                              "norm2": eng.norm2(st)}
print(out["config3_variational"], flush=True)
del st
torch.cuda.empty_cache()

# ---- config 4: 30-qubit random circuit, 1e6 shots, marginals
n, ngates, nshots = (24, 60, 10**5) if small else (30, 300, 10**6)
ops = circuits.random_circuit(n, ngates, seed=11)
st = eng.basis_state(n, "complex128")
t_circ, stats = timed(lambda: eng.apply_program(eng.basis_state(n, "complex128") if False else st, n, ops), reps=1)
res = {"nqubits": n, "gates": len(ops), "sweeps": stats.nsweeps, "circuit_seconds": t_circ, "gates_per_s": len(ops) / t_circ}
t, probs = timed(lambda: eng.probabilities(st, list(range(n)), n))
res["probabilities_all_seconds"] = t
res["probabilities_all_GBs"] = (16 + 8) * 2.0**n / t / 1e9
np.random.seed(1234)
u = np.random.random_sample(nshots)
t, shots = timed(lambda: eng.sample(probs, u), reps=1)
res["sample_seconds"] = t
res["shots_per_s"] = nshots / t
res["nshots"] = nshots
# search exactness against the CDF the device built (parallel scan at this size)
from qibo_b200 import _lib  # noqa: E402

cdf = eng.cdf(probs, _lib.QB_SCAN_PARALLEL)
chk = torch.searchsorted(cdf.tensor, torch.from_numpy(u).cuda(), right=True).cpu().numpy()
res["samples_equal_searchsorted_on_same_cdf"] = bool(np.array_equal(chk, shots))
for q in ([0], [0, 5, 7], list(range(10)), [1, 5, 2, 0]):
    t, m = timed(lambda: eng.probabilities(st, q, n))
    res[f"marginal_{len(q)}q_seconds"] = t
    res[f"marginal_{len(q)}q_sum"] = float(m.numpy().sum())
out["config4_sampling"] = res
print(res, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_bench.json"), "w"), indent=1)
