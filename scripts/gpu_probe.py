"""GPU probe: per-kernel timings (CUDA events) for the roofline work.  Writes gpurun_out/probe_<tag>.json.

    python scripts/gpu_probe.py [--n 30] [--tag r1a]
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
from qibo_b200.ops import Op  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=30)
    ap.add_argument("--tag", default="probe")
    ap.add_argument("--dtype", default="complex128")
    ap.add_argument("--skip-k1", action="store_true")
    args = ap.parse_args()
    n, dtype = args.n, args.dtype
    B = 16 if dtype == "complex128" else 8
    eng = Engine(0)
    out = {"n": n, "dtype": dtype, "gpu": torch.cuda.get_device_name(0)}
    st = eng.basis_state(n, dtype)
    full = 2.0 * B * 2.0**n

    # reference copy bandwidth on this box
    a = torch.empty(2**n, dtype=torch.complex128 if B == 16 else torch.complex64, device="cuda")
    med, mn = timeit(lambda: a.copy_(st.tensor))
    out["torch_copy_GBs"] = full / (mn * 1e-3) / 1e9
    del a

    h = circuits.matrix("H")
    rng = np.random.default_rng(0)
    u2 = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
    res = []
    if not args.skip_k1:
        for q in [0, 1, n // 2, n - 6, n - 4, n - 3, n - 2, n - 1]:
            op = Op(h, (q,))
            med, mn = timeit(lambda: eng.apply_op(st, n, op))
            res.append({"kernel": "k1_dense1", "qubit": q, "bit": n - 1 - q, "ms": mn, "GBs": full / (mn * 1e-3) / 1e9})
        for a_, b_ in [(0, 1), (0, n - 1), (n - 2, n - 1)]:
            op = Op(u2, (a_, b_))
            med, mn = timeit(lambda: eng.apply_op(st, n, op))
            res.append({"kernel": "k1_dense2", "qubits": [a_, b_], "ms": mn, "GBs": full / (mn * 1e-3) / 1e9})
        op = circuits.op("CU1", (3, 0), 0.3)
        med, mn = timeit(lambda: eng.apply_op(st, n, op))
        res.append({"kernel": "k1_phase(CU1)", "ms": mn, "GBs_alg(quarter)": full / 4 / (mn * 1e-3) / 1e9})
        op = circuits.op("SWAP", (0, n - 1))
        med, mn = timeit(lambda: eng.apply_op(st, n, op))
        res.append({"kernel": "k1_swap", "ms": mn, "GBs_alg(half)": full / 2 / (mn * 1e-3) / 1e9})
    out["k1"] = res

    # sweep kernel: one op per sweep, various run lengths (low bits) and target positions
    sw = []
    for low in ["4", "5", "6", "7"] if B == 16 else ["5", "6", "7", "8"]:
        os.environ["QB_SWEEP_LOW_BITS"] = low
        for q in [0, n // 2, n - 1]:
            ops = [Op(h, (q,))]
            med, mn = timeit(lambda: eng.apply_program(st, n, ops))
            sw.append({"low_bits": int(low), "ops": "H(%d)" % q, "ms": mn, "GBs": full / (mn * 1e-3) / 1e9})
        # 6 single-qubit gates on the highest qubits (as many high bits as the tile takes)
        ops = [Op(h, (q,)) for q in range(6)]
        med, mn = timeit(lambda: eng.apply_program(st, n, ops))
        st_, _ = plan_program(n, dtype, ops)
        sw.append({"low_bits": int(low), "ops": "H(0..5)", "sweeps": st_.nsweeps, "ms": mn, "GBs_per_sweep": st_.nsweeps * full / (mn * 1e-3) / 1e9})
    del os.environ["QB_SWEEP_LOW_BITS"]
    out["sweep_single"] = sw

    # gates per sweep: k single-qubit gates on bits inside the tile
    pp = []
    os.environ["QB_SWEEP_MAX_OPS"] = "96"
    ry = circuits.matrix("RY", 0.37)
    rx = circuits.matrix("RX", 0.37)
    for label, mat, bits in (("H mid bits 5..11", h, list(range(5, 12))), ("H low bits 0..2", h, [0, 1, 2]),
                             ("RX(complex) mid bits", rx, list(range(5, 12))), ("H bits 3..5", h, [3, 4, 5])):
        for k in [1, 3, 6, 12, 24]:
            ops = [Op(mat, (n - 1 - bits[i % len(bits)],)) for i in range(k)]
            st_, _ = plan_program(n, dtype, ops)
            med, mn = timeit(lambda: eng.apply_program(st, n, ops))
            pp.append({"what": label, "gates": k, "sweeps": st_.nsweeps, "passes": st_.ndense_passes, "ms": mn,
                       "GBs_per_sweep": st_.nsweeps * full / (mn * 1e-3) / 1e9})
    for k in [1, 2, 4, 8]:
        ops = [circuits.op("CU1", (n - 1 - i, 0), 0.1 * (i + 1)) for i in range(3)] * 1
        ops = []
        for j in range(k):  # k fans (distinct control), each with 10 phases
            ops += [circuits.op("CU1", (n - 1 - i, j), 0.1 * (i + 1)) for i in range(10)]
        st_, _ = plan_program(n, dtype, ops)
        med, mn = timeit(lambda: eng.apply_program(st, n, ops))
        pp.append({"fans": k, "sweeps": st_.nsweeps, "ms": mn, "GBs_per_sweep": st_.nsweeps * full / (mn * 1e-3) / 1e9})
    del os.environ["QB_SWEEP_MAX_OPS"]
    out["sweep_passes"] = pp

    # whole circuits
    circ = []
    for mp in ["12", "24", "96"]:
        os.environ["QB_SWEEP_MAX_OPS"] = mp
        ops = circuits.qft(n)
        st_, _ = plan_program(n, dtype, ops)
        eng.basis_state(n, dtype)
        med, mn = timeit(lambda: eng.apply_program(st, n, ops), reps=3, warm=1)
        circ.append({"circuit": f"QFT({n})", "max_ops": int(mp), "passes": st_.ndense_passes, "gates": len(ops), "sweeps": st_.nsweeps, "ms": mn,
                     "gates_per_s": len(ops) / (mn * 1e-3), "GBs_per_sweep": st_.nsweeps * full / (mn * 1e-3) / 1e9})
    del os.environ["QB_SWEEP_MAX_OPS"]
    ops = circuits.qft(n)
    med, mn = timeit(lambda: eng.apply_program(st, n, ops, fuse=False), reps=2, warm=1)
    circ.append({"circuit": f"QFT({n}) one sweep per gate", "gates": len(ops), "ms": mn, "gates_per_s": len(ops) / (mn * 1e-3),
                 "GBs_per_sweep": len(ops) * full / (mn * 1e-3) / 1e9})
    if not args.skip_k1:
        def k1_all():
            for op in ops:
                eng.apply_op(st, n, op)
        med, mn = timeit(k1_all, reps=2, warm=1)
        circ.append({"circuit": f"QFT({n}) K1 gate-by-gate", "gates": len(ops), "ms": mn, "gates_per_s": len(ops) / (mn * 1e-3)})
    out["circuits"] = circ

    # measurement kernels
    meas = []
    for qubits in ([0], [0, 5, 7], list(range(10)), [1, 5, 2, 0], list(range(n))):
        med, mn = timeit(lambda: eng.probabilities(st, qubits, n), reps=3, warm=1)
        meas.append({"kernel": "k3_probs", "m": len(qubits), "ms": mn, "GBs_read": B * 2.0**n / (mn * 1e-3) / 1e9})
    out["measure"] = meas

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"probe_{args.tag}.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
