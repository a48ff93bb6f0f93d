"""One-GPU reproduction of what rank R of W does in a sharded QFT(n): every local segment of the plan runs through
Engine.apply_program on a 2^(n-g) shard (exchanges skipped: they only move data), optionally checked against the
gate-by-gate K1 kernels on a second copy.   python scripts/dist_local_repro.py N W RANK [--check] [--c64] [--block]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qibo_b200 import circuits, distributed as D  # noqa: E402
from qibo_b200.engine import Engine  # noqa: E402


def main():
    n, W, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    check = "--check" in sys.argv
    dtype = "complex64" if "--c64" in sys.argv else "complex128"
    g = W.bit_length() - 1
    nlocal = n - g
    eng = Engine(0)
    plan = D.Plan(n, g, circuits.qft(n)) if "--block" in sys.argv else D.choose_layout(n, g, circuits.qft(n))
    print("global qubits", plan.global_qubits, "exchanges", plan.nexchanges, flush=True)
    gen = torch.Generator(device="cuda").manual_seed(1)
    tdt = torch.float64 if dtype == "complex128" else torch.float32
    a = eng.empty((1 << nlocal,), dtype)
    torch.view_as_real(a.tensor).normal_(generator=gen).mul_(2.0 ** (-(nlocal + 1) / 2))
    b = eng.copy(a) if check else None
    print(f"n={n} W={W} rank={r} nlocal={nlocal} dtype={dtype} tdt={tdt}", flush=True)
    for si, seg in enumerate(plan.segments):
        if seg.kind != "local":
            continue
        local = [o for o in (D.specialise(p, nlocal, r) for p in seg.ops) if o is not None]
        if not local:
            continue
        t0 = time.time()
        st = eng.apply_program(a, nlocal, local, timed=True)
        eng.synchronize()
        msg = f"seg {si}: {len(local)} ops, {st.nsweeps} sweeps, {st.elapsed_ms:.2f} ms (wall {time.time() - t0:.2f} s)"
        if check:
            for op in local:
                eng.apply_op(b, nlocal, op)
            eng.synchronize()
            err = (a.tensor - b.tensor).abs().max().item()
            msg += f" max|sweep - gatewise| = {err:.3e}"
        print(msg, flush=True)
    print("done", flush=True)


if __name__ == "__main__":
    main()
