"""Times the sweep kernel of one library build (QB_LIB_PATH) on a few fixed programs, with correctness guards.

    QB_LIB_PATH=qibo_b200/lib/libqibo_b200_X.so python scripts/variant_bench.py [--n 30] [--tag X]

Prints one JSON line per case: per-sweep CUDA-event times (each sweep run as its own timed program).
"""

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qibo_b200 import circuits  # noqa: E402
from qibo_b200.engine import Engine, plan_program  # noqa: E402


def sweeps_of(n, dtype, ops):
    """Split ``ops`` at the planner's sweep boundaries."""
    _, sweep_of_op = plan_program(n, dtype, ops)
    groups = {}
    for i, s in enumerate(sweep_of_op):
        groups.setdefault(int(s), []).append(ops[i])
    return [groups[k] for k in sorted(groups) if k >= 0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=30)
    ap.add_argument("--tag", default="")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    n = args.n
    eng = Engine(0)
    out = {"tag": args.tag, "lib": os.environ.get("QB_LIB_PATH", "default"), "n": n}

    # ---- correctness guards (small, against torch.fft and norm) --------------------------------------------------
    m = 20
    rng = np.random.default_rng(1)
    psi = rng.normal(size=2**m) + 1j * rng.normal(size=2**m)
    psi /= np.linalg.norm(psi)
    for dt, tol in (("complex128", 1e-12), ("complex64", 1e-5)):
        st = eng.upload(psi.astype(dt))
        eng.apply_program(st, m, circuits.qft(m))
        ref = np.fft.ifft(psi.astype(dt).astype(np.complex128), norm="ortho")
        err = float(np.abs(st.numpy() - ref).max())
        out[f"qft{m}_{dt}_err"] = err
        assert err < tol, (dt, err)

    # ---- QFT(n) complex128 ------------------------------------------------------------------------------------------
    for dt in ("complex128",):
        B = 16 if dt == "complex128" else 8
        ops = circuits.qft(n, with_swaps=False)
        st = eng.basis_state(n, dt)
        for _ in range(2):
            eng.apply_program(st, n, ops)
        ts = []
        for _ in range(args.reps):
            st = eng.basis_state(n, dt)
            stats = eng.apply_program(st, n, ops, timed=True)
            ts.append(stats.elapsed_ms)
        amp = st.tensor[:4].cpu().numpy()
        assert np.abs(np.abs(amp) - 2.0 ** (-n / 2)).max() < 1e-12 * 2.0 ** (-n / 2) * 1e4, amp
        per = []
        for grp in sweeps_of(n, dt, ops):
            t = []
            for _ in range(args.reps):
                s = eng.apply_program(st, n, grp, timed=True)
                t.append(s.elapsed_ms)
            per.append({"gates": len(grp), "sweeps": s.nsweeps, "ms": min(t), "GBs": 2.0 * B * 2.0**n * s.nsweeps / (min(t) * 1e-3) / 1e9})
        out[f"qft_{dt}"] = {"ms": min(ts), "sweeps": stats.nsweeps, "per_sweep": per}

    # ---- variational complex64 (2 layers) ---------------------------------------------------------------------------
    dt = "complex64"
    nv = n + 1
    thetas = 2 * np.pi * np.random.default_rng(7).random(2 * 2 * nv)
    ops = circuits.variational(nv, 2, thetas)
    st = eng.basis_state(nv, dt)
    eng.apply_program(st, nv, ops)
    ts = []
    for _ in range(args.reps):
        st = eng.basis_state(nv, dt)
        stats = eng.apply_program(st, nv, ops, timed=True)
        ts.append(stats.elapsed_ms)
    nrm = eng.norm2(st)
    assert abs(nrm - 1.0) < 1e-3, nrm
    out["var_complex64"] = {"n": nv, "ms": min(ts), "sweeps": stats.nsweeps, "ms_per_sweep": min(ts) / stats.nsweeps,
                            "GBs_per_sweep": 2.0 * 8 * 2.0**nv * stats.nsweeps / (min(ts) * 1e-3) / 1e9, "norm2": nrm}

    # ---- random circuit complex128 -----------------------------------------------------------------------------------
    dt = "complex128"
    ops = circuits.random_circuit(n, 100, 11)
    st = eng.basis_state(n, dt)
    eng.apply_program(st, n, ops)
    ts = []
    for _ in range(args.reps):
        st = eng.basis_state(n, dt)
        stats = eng.apply_program(st, n, ops, timed=True)
        ts.append(stats.elapsed_ms)
    nrm = eng.norm2(st)
    assert abs(nrm - 1.0) < 1e-9, nrm
    out["random_complex128"] = {"ms": min(ts), "sweeps": stats.nsweeps, "ms_per_sweep": min(ts) / stats.nsweeps,
                                "GBs_per_sweep": 2.0 * 16 * 2.0**n * stats.nsweeps / (min(ts) * 1e-3) / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
