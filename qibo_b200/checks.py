"""Device-side self-checks that need no oracle: closed forms evaluated on the GPU next to the state, chunk by chunk, so
they cover EVERY amplitude at the benchmark sizes (2^30 .. 2^32) where no CPU restatement finishes.  Used by the
``-m gpu`` tests and by ``bench.py``'s ``verify`` block; plain torch arithmetic in int64 / float64 (checker code, not
the product path)."""

import math

import torch


def qft_basis_state_error(shard: torch.Tensor, nqubits: int, x: int, canonical_index=None, chunk: int = 1 << 24) -> float:
    """max_k |amp[k] - exp(2 pi i x k / 2^n) / sqrt(2^n)| * sqrt(2^n)  over all amplitudes of ``shard``.

    QFT|x> in Qibo's convention (models/qft.py:47-58: H + CU1 ladder + the closing SWAPs) is the inverse DFT of the basis
    state: amplitude k = exp(+2 pi i x k / 2^n) / sqrt(2^n), k the canonical index (qubit 0 = most significant bit).  Every
    H and every CU1 of the circuit acts non-trivially on a generic |x>, so a wrong fan table, a wrong per-tile factor or a
    misplaced tile shows up in the phase of some amplitude.  ``canonical_index``: maps the int64 tensor of shard-local
    indices to canonical indices (sharded runs); default: the identity.  The phase x*k mod 2^n is exact in int64 (wrapping
    multiplication keeps the low 64 bits), the angle is then one float64 rounding away from exact."""
    n = nqubits
    total = shard.numel()
    scale = 2.0 ** (n / 2.0)
    mask = (1 << n) - 1
    worst = 0.0
    dev = shard.device
    real_dtype = torch.float64
    xs = x & mask
    # two's-complement: keep x as a signed int64 multiplier
    xmul = xs if xs < (1 << 63) else xs - (1 << 64)
    for start in range(0, total, chunk):
        stop = min(total, start + chunk)
        loc = torch.arange(start, stop, device=dev, dtype=torch.int64)
        k = loc if canonical_index is None else canonical_index(loc)
        phase = (k * xmul) & mask
        angle = phase.to(real_dtype) * (2.0 * math.pi / float(1 << n))
        got = shard[start:stop].to(torch.complex128) * scale
        err = torch.maximum((got.real - torch.cos(angle)).abs().max(), (got.imag - torch.sin(angle)).abs().max())
        worst = max(worst, float(err.item()))
        del loc, k, phase, angle, got
    return worst


def generic_basis_state(nqubits: int) -> int:
    """A fixed basis state with a generic bit pattern (both values on neighbouring qubits, lowest and highest bit set)."""
    pattern = "1011001110001011010111001010011011010011"
    return int("1" + pattern[: nqubits - 2] + "1", 2) if nqubits >= 2 else 1
