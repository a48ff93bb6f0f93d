"""Measurement of a state sharded over global qubits (SURVEY.md 8e "reduction-type collectives").

Layout: the canonical one ``ShardedProgram`` leaves behind -- rank ``r`` of ``W = 2^g`` holds the ``2^(n-g)`` amplitudes
whose ``g`` leading qubits (qubit 0 = most significant bit, tests/test_gates_gates.py:39-43) spell ``r``.

* probabilities / marginals (abstract.py:2734-2758): every rank reduces its shard with K3 over the measured LOCAL
  qubits; measured GLOBAL qubits select where a rank's partial lands in the caller-ordered result; one all-reduce(sum).
  All qubits in ascending order: the result stays sharded like the state (no collective at all).
* shot sampling (abstract.py:2774-2781): per-rank mass -> all-gather -> exclusive scan over ranks; every rank resolves
  the uniforms that fall into its interval with K4 (CDF + search) on its own bins; one all-reduce(sum) of int64.
* collapse (abstract.py:3279-3304): ranks whose global bits contradict the outcome zero their shard, the others project
  locally with K5; the norm is one scalar all-reduce.

The per-shard arithmetic goes through a small provider object: ``EngineLocal`` (the CUDA kernels, the product path) or,
in the world-size-2 gloo tests on CPU, a NumPy restatement supplied by the test-suite.
"""

from typing import Sequence

import numpy as np
import torch
import torch.distributed as dist

from qibo_b200.ops import Op


class EngineLocal:
    """Shard-local pieces on the GPU: K3 probabilities, K4 CDF/search, K5 collapse, K1 scaling."""

    def __init__(self, engine):
        self.engine = engine

    def probabilities(self, shard, qubits, nlocal):  # -> torch tensor (real)
        return self.engine.probabilities(shard, list(qubits), nlocal).tensor

    def cdf(self, probs):  # normalised CDF (float64) of a real tensor
        from qibo_b200 import _lib
        from qibo_b200.array import DeviceArray

        arr = DeviceArray(probs.contiguous())
        mode = _lib.QB_SCAN_EXACT if arr.size <= (1 << 22) else _lib.QB_SCAN_PARALLEL
        return self.engine.cdf(arr, mode).tensor

    def search(self, cdf, uniforms):  # #{k : cdf[k] <= u}, int64
        from qibo_b200.array import DeviceArray

        return self.engine.sample_cdf(DeviceArray(cdf), DeviceArray(uniforms.contiguous())).tensor

    def collapse(self, shard, nlocal, qubits, outcome):
        self.engine.collapse(shard, nlocal, list(qubits), outcome, normalize=False)

    def zero(self, shard):
        shard.tensor.zero_()

    def norm2(self, shard):
        return self.engine.norm2(shard)

    def vdot(self, a, b, nlocal):
        return self.engine.vdot(a, b, nlocal)

    def scale(self, shard, nlocal, factor):
        self.engine.apply_op(shard, nlocal, Op(np.array([factor, factor]), (0,), is_diagonal=True))

    def device(self, shard):
        return shard.tensor.device


class ShardMeasure:
    def __init__(self, nqubits: int, local, group=None):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.g = int(round(np.log2(self.world)))
        if 1 << self.g != self.world:
            raise ValueError("the number of ranks must be a power of two")
        self.n, self.nlocal = nqubits, nqubits - self.g
        self.local, self.group = local, group

    def _all_reduce(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def _split(self, qubits: Sequence[int]):
        qubits = [int(q) for q in qubits]
        if len(set(qubits)) != len(qubits) or any(q < 0 or q >= self.n for q in qubits):
            raise ValueError(f"bad measured qubits {qubits}")
        return qubits, [q - self.g for q in qubits if q >= self.g]

    # ---- scalars ------------------------------------------------------------------------------------------------------
    def global_norm2(self, shard) -> float:
        """<psi|psi> over all shards (callbacks.Norm, callbacks.py:161-174)."""
        t = torch.tensor([self.local.norm2(shard)], dtype=torch.float64, device=self.local.device(shard))
        return float(self._all_reduce(t).item())

    def global_vdot(self, a, b) -> complex:
        """<a|b> over all shards (Backend.overlap_statevector, abstract.py:2180-2190)."""
        v = complex(self.local.vdot(a, b, self.nlocal))
        t = torch.tensor([v.real, v.imag], dtype=torch.float64, device=self.local.device(b))
        self._all_reduce(t)
        return complex(float(t[0].item()), float(t[1].item()))

    def gather(self, shard) -> np.ndarray:
        """The full state in canonical order as a host array on every rank (block layout; small registers only)."""
        tensor = shard.tensor.clone()
        if self.world == 1:
            return tensor.cpu().numpy()
        parts = [torch.empty_like(tensor) for _ in range(self.world)]
        dist.all_gather(parts, tensor, group=self.group)
        return torch.cat(parts).cpu().numpy()

    # ---- probabilities ----------------------------------------------------------------------------------------------
    def probabilities(self, shard, qubits: Sequence[int]):
        """-> (tensor, sharded).  ``sharded`` is True only for ``qubits == range(n)``: the tensor is then this rank's
        slice of the 2^n probabilities; otherwise the full caller-ordered marginal, replicated on every rank."""
        qubits, lq = self._split(qubits)
        part = self.local.probabilities(shard, lq, self.nlocal)
        if self.world == 1:
            return part, False
        if qubits == list(range(self.n)):
            return part, True
        m, ml = len(qubits), len(lq)
        if m > 30:
            raise NotImplementedError("a replicated marginal over more than 30 qubits; measure all qubits in order for a sharded result")
        if ml == m:  # only local qubits: plain sum over ranks
            return self._all_reduce(part.clone()), False
        # caller-ordered index = global measured bits (fixed by the rank) interleaved with the local measured bits
        base, j = 0, torch.arange(1 << ml, device=part.device, dtype=torch.int64)
        idx = torch.zeros_like(j)
        seen_local = 0
        for pos, q in enumerate(qubits):
            shift = m - 1 - pos
            if q < self.g:
                base |= ((self.rank >> (self.g - 1 - q)) & 1) << shift
            else:
                idx |= ((j >> (ml - 1 - seen_local)) & 1) << shift
                seen_local += 1
        out = torch.zeros(1 << m, device=part.device, dtype=part.dtype)
        out[idx + base] = part
        return self._all_reduce(out), False

    # ---- sampling -----------------------------------------------------------------------------------------------------
    def sample(self, probs, uniforms, sharded: bool):
        """Inverse-CDF samples (int64, replicated) for host ``uniforms``; ``probs`` as returned by ``probabilities``."""
        u = torch.as_tensor(np.ascontiguousarray(uniforms, dtype=np.float64)).to(probs.device)
        if not sharded or self.world == 1:
            return self.local.search(self.local.cdf(probs), u)
        mass = probs.sum(dtype=torch.float64).reshape(1)
        masses = [torch.zeros_like(mass) for _ in range(self.world)]
        dist.all_gather(masses, mass, group=self.group)
        masses = torch.cat(masses)
        edges = torch.cat([torch.zeros(1, dtype=torch.float64, device=mass.device), torch.cumsum(masses, 0)])
        total = edges[-1]
        lo, hi = edges[self.rank], edges[self.rank + 1]
        x = u * total  # position on the unnormalised global CDF
        mine = (x >= lo) if self.rank == self.world - 1 else (x >= lo) & (x < hi)
        local_u = torch.clamp((x - lo) / torch.clamp(mass[0], min=1e-300), 0.0, np.nextafter(1.0, 0.0))
        nbins = probs.numel()
        idx = self.local.search(self.local.cdf(probs), local_u)
        idx = torch.clamp(idx, max=nbins - 1) + self.rank * nbins
        out = torch.where(mine, idx, torch.zeros_like(idx))
        return self._all_reduce(out)

    # ---- collapse -----------------------------------------------------------------------------------------------------
    def collapse(self, shard, qubits: Sequence[int], outcome: int, normalize: bool = True):
        """Project the sharded state onto ``outcome`` (decimal over ``qubits``, qubits[0] = MSB; sorted, as
        gates/measurements.py:199 passes them) in place, and renormalise globally."""
        qubits, lq = self._split(qubits)
        m = len(qubits)
        match, local_outcome = True, 0
        for pos, q in enumerate(qubits):
            bit = (int(outcome) >> (m - 1 - pos)) & 1
            if q < self.g:
                match &= ((self.rank >> (self.g - 1 - q)) & 1) == bit
            else:
                local_outcome = (local_outcome << 1) | bit
        if not match:
            self.local.zero(shard)
        elif lq:
            self.local.collapse(shard, self.nlocal, lq, local_outcome)
        if normalize:
            nrm = torch.tensor([self.local.norm2(shard)], dtype=torch.float64, device=self.local.device(shard))
            self._all_reduce(nrm)
            self.local.scale(shard, self.nlocal, 1.0 / float(np.sqrt(nrm.item())))
        return shard
