"""The device-side closed-form checker (qibo_b200/checks.py) against the oracle, on CPU tensors: it is what the -m gpu
tests and bench.py rely on at 2^30 .. 2^32 amplitudes, so it has to be right about Qibo's QFT convention itself."""

import numpy as np
import pytest
import torch

from helpers import ops_from_named, oracle_run
from oracle import numpy_oracle as orc
from qibo_b200.checks import generic_basis_state, qft_basis_state_error


@pytest.mark.parametrize("n", [2, 5, 9, 12])
def test_closed_form_matches_oracle_qft(n):
    x = generic_basis_state(n)
    assert 0 < x < 2**n and x & 1 and x >> (n - 1)
    psi = np.zeros(2**n, dtype=np.complex128)
    psi[x] = 1
    out = orc.run_ops(psi, orc.qft_ops(n), n)
    t = torch.from_numpy(out)
    assert qft_basis_state_error(t, n, x, chunk=1 << 7) < 1e-12
    # a single wrong phase anywhere is seen
    bad = out.copy()
    bad[3 % 2**n] *= np.exp(1e-6j)
    assert qft_basis_state_error(torch.from_numpy(bad), n, x) > 5e-7
    # sharded form: two halves with a canonical-index map
    half = 2 ** (n - 1)
    for r in range(2):
        e = qft_basis_state_error(t[r * half : (r + 1) * half], n, x, canonical_index=lambda loc, r=r: loc + r * half)
        assert e < 1e-12


def test_closed_form_phase_arithmetic_is_exact_at_32_qubits():
    """x * k overflows int64 at n = 32; the wrapped product still has the right low n bits."""
    n = 32
    x = generic_basis_state(n)
    ks = torch.tensor([0, 1, 2**31 + 12345, 2**32 - 1, 3_000_000_019], dtype=torch.int64)
    mask = (1 << n) - 1
    got = (ks * x) & mask
    want = [(int(k) * x) % (1 << n) for k in ks.tolist()]
    assert got.tolist() == want
