"""Global-qubit distributed execution: one process per GPU, ``torch.distributed`` for the plumbing.

Re-implements the scheme of ``qibo.models.distcircuit`` (DistributedQubits / DistributedQueues,
models/distcircuit.py:9-329) for an executor the reference never shipped
(``Backend.execute_distributed_circuit`` raises NotImplementedError, backends/abstract.py:2638-2647).

Layout.  W = 2^g ranks; g logical qubits are GLOBAL -- their values spell the rank -- and the other qubits index the
shard in ascending order (piece layout ``sorted(global) + local``, distcircuit.py:33-36).  Which qubits are global is
the planner's choice, as in the reference (DistributedQubits, distcircuit.py:287-298): the leading ones (block layout,
the default), the trailing ones (cyclic layout, what ``_DistributedQFT`` asks for, models/qft.py:66), any set, or
"auto" = whichever needs the fewest exchanges.  Instead of relabelling the caller's gate objects in place
(distcircuit.py:236-244) the planner keeps a logical-qubit -> physical-bit map:

  * a gate needs a qubit LOCAL only if its matrix mixes that qubit's 0/1 subspaces; control qubits and every
    qubit of a diagonal gate may stay global -- on each rank the gate is specialised to the rank's bit values
    (the reference does the same for global controls, distcircuit.py:316-327);
  * when a mixing target sits on a global bit it is exchanged with a high local bit whose qubit is needed
    furthest in the future and, among equals, belongs on that global bit at the end; once a segment is cut, the
    other global qubits with mixing gates ahead come in too while finished qubits can leave for them, so exchanges
    sit back to back (pairwise half-shard exchange, contiguous chunks of >= 2^20 amplitudes; a run of them on the
    leading local bits is ONE all-to-all kernel over NVLink peer memory);
  * uncontrolled SWAP gates are applied as relabelling; the layout is restored at the end
    (the reference appends reverse swaps for the same purpose, distcircuit.py:267).

The local gate runs go through the same sweep kernels as the single-GPU path (``Engine.apply_program``).
The planner and the specialisation are pure host code: ``tests/test_distributed_cpu.py`` drives them with a
NumPy shard executor over a world_size-2/4 gloo group.
"""

import math
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from qibo_b200.ops import Op


def world_size():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_world_size()
    return 1


def rank():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank()
    return 0


# ---------------------------------------------------------------------------------------------- planning
@dataclass
class PhysOp:
    """An Op with its qubits resolved to physical bit positions at planning time."""

    data: np.ndarray
    tbits: Tuple[int, ...]  # targets[0] first (MSB of the matrix index)
    cbits: Tuple[int, ...]
    is_diagonal: bool


@dataclass
class Segment:
    kind: str  # "local" | "exchange"
    ops: Optional[List[PhysOp]] = None
    gbit: int = -1  # exchange: global physical bit
    lbit: int = -1  # exchange: local physical bit


def exchange_runs(segments):
    """Consecutive exchange segments on pairwise distinct bits, as lists of (gbit, lbit): they commute, so a run can be
    carried out as ONE all-to-all (ShardedProgram.run) -- [("local", ops) | ("exchange", [(gbit, lbit), ...])]."""
    out = []
    for seg in segments:
        if seg.kind == "local":
            out.append(("local", seg.ops))
            continue
        if out and out[-1][0] == "exchange":
            used = {b for pr in out[-1][1] for b in pr}
            if seg.gbit not in used and seg.lbit not in used:
                out[-1][1].append((seg.gbit, seg.lbit))
                continue
        out.append(("exchange", [(seg.gbit, seg.lbit)]))
    return out


def split_trailing_permutation(phys_ops, nlocal: int, k: int, min_swaps: int = 3):
    """EXPERIMENTAL (opt-in, QB_A2A_FUSE_PERM=1; not measured yet -- DESIGN.md section 8, step 1).  For the local segment
    that follows a run of k exchanges on the k leading local bits: if it is `gates G` + a closing run of plain SWAPs, the
    SWAPs only move bits BELOW the k leading local bits and G only touches those k bits (or global ones), then permutation
    and G commute and the permutation can ride on the all-to-all (every chunk is written permuted).
    -> (G as PhysOps, dest_of_qubit over the nlocal - k lower local qubits) or None.  Decided on the PLAN's ops, not on a
    rank's specialised ones, so that every rank takes the same decision."""
    lo = nlocal - k
    nswaps = 0
    for p in reversed(phys_ops):
        if p.is_diagonal or p.cbits or len(p.tbits) != 2 or not np.array_equal(p.data, _SWAP):
            break
        nswaps += 1
    if nswaps < min_swaps:
        return None
    gates, swaps = list(phys_ops[: len(phys_ops) - nswaps]), phys_ops[len(phys_ops) - nswaps :]
    if any(b < lo for p in gates for b in tuple(p.tbits) + tuple(p.cbits)):
        return None
    if any(b >= lo for p in swaps for b in p.tbits):
        return None
    dest = list(range(lo))  # sub-qubit s of a chunk <-> bit lo - 1 - s
    for p in swaps:
        a, b = lo - 1 - p.tbits[0], lo - 1 - p.tbits[1]
        dest = [b if d == a else a if d == b else d for d in dest]
    if dest == list(range(lo)):
        return None
    return gates, dest


def alltoall_entries(rank_: int, nlocal: int, pairs):
    """The chunk swaps of rank ``rank_`` for a run of exchanges whose local bits are the k leading ones: chunk t (the k
    leading bits of the shard index) trades places with chunk t' of rank r', where every (gbit, lbit) pair swaps one bit
    of r with one bit of t.  -> [(peer rank, my offset, peer offset, begin, end)] in amplitudes, or None when the local
    bits are not the leading ones.  The two ranks of a pair split the chunk: the lower rank moves the first half."""
    k = len(pairs)
    lo = nlocal - k
    if sorted(l for _, l in pairs) != list(range(lo, nlocal)) or len({g for g, _ in pairs}) != k:
        return None
    csz = 1 << lo
    out = []
    for t in range(1 << k):
        r2, t2 = rank_, t
        for gbit, lbit in pairs:
            j, c = gbit - nlocal, lbit - lo
            rb, tb = (rank_ >> j) & 1, (t >> c) & 1
            r2 = (r2 & ~(1 << j)) | (tb << j)
            t2 = (t2 & ~(1 << c)) | (rb << c)
        if r2 == rank_:
            continue
        begin, end = (0, csz // 2) if rank_ < r2 else (csz // 2, csz)
        out.append((r2, t * csz, t2 * csz, begin, end))
    return out


def alltoall_push_entries(rank_: int, nlocal: int, pairs):
    """Out-of-place form: EVERY chunk t of this rank (the one that stays included) is copied to chunk t' of rank r''s second
    buffer -> [(destination rank, my offset, its offset, 0, chunk size)], or None (see ``alltoall_entries``)."""
    k = len(pairs)
    lo = nlocal - k
    if sorted(l for _, l in pairs) != list(range(lo, nlocal)) or len({g for g, _ in pairs}) != k:
        return None
    csz = 1 << lo
    out = []
    for t in range(1 << k):
        r2, t2 = rank_, t
        for gbit, lbit in pairs:
            j, c = gbit - nlocal, lbit - lo
            rb, tb = (rank_ >> j) & 1, (t >> c) & 1
            r2 = (r2 & ~(1 << j)) | (tb << j)
            t2 = (t2 & ~(1 << c)) | (rb << c)
        out.append((r2, t * csz, t2 * csz, 0, csz))
    return out


def mixing_targets(op: Op) -> List[int]:
    """Targets whose 0/1 subspaces the matrix mixes -- the only qubits that must be local."""
    if op.is_diagonal:
        return []
    k = len(op.targets)
    dim = 1 << k
    m = op.data
    rows, cols = np.nonzero(m)
    diff = np.bitwise_xor(rows, cols)
    mixed = 0
    for d in np.unique(diff):
        mixed |= int(d)
    return [op.targets[i] for i in range(k) if (mixed >> (k - 1 - i)) & 1]


def is_plain_swap(op: Op) -> bool:
    if op.is_diagonal or len(op.targets) != 2 or op.controls:
        return False
    sw = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
    return bool(np.array_equal(op.data, sw))


_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def block_layout(nqubits: int, nglobal: int) -> Tuple[int, ...]:
    """Global qubits = the ``g`` leading qubits: rank r holds a contiguous block of the canonical state."""
    return tuple(range(nglobal))


def cyclic_layout(nqubits: int, nglobal: int) -> Tuple[int, ...]:
    """Global qubits = the ``g`` trailing qubits (what ``_DistributedQFT`` asks for, models/qft.py:66): rank r holds the
    amplitudes whose canonical index is congruent to r modulo 2^g."""
    return tuple(range(nqubits - nglobal, nqubits))


class Plan:
    """Host-side plan: identical on every rank (pure function of the op list, n, g and the layout).

    ``global_qubits[j]`` is the logical qubit that sits on the j-th most significant rank bit (physical bit n-1-j)
    before the first and after the last gate; the local qubits follow in ascending order (the first local qubit is the
    most significant bit of the shard index).  The reference picks its global qubits per circuit as well
    (DistributedQubits, distcircuit.py:287-298; ``_DistributedQFT`` makes them the trailing ones): with the trailing
    qubits global a QFT needs one exchange per global qubit, because its closing SWAPs put every early, finished qubit
    exactly where a late one has to leave from."""

    def __init__(self, nqubits: int, nglobal: int, ops: Sequence[Op], relabel_swaps: bool = True, high_window: int = 8,
                 global_qubits: Optional[Sequence[int]] = None, batch_exchanges: bool = True,
                 final_global_qubits: Optional[Sequence[int]] = None):
        self.n, self.g = nqubits, nglobal
        self.nlocal = nqubits - nglobal
        if self.nlocal < 1:
            raise ValueError("need at least one local qubit per rank")
        n, nlocal = self.n, self.nlocal
        gq = tuple(int(q) for q in (block_layout(n, nglobal) if global_qubits is None else global_qubits))
        if len(gq) != nglobal or len(set(gq)) != nglobal or any(q < 0 or q >= n for q in gq):
            raise ValueError(f"global_qubits must name {nglobal} distinct qubits, got {gq}")
        self.global_qubits = gq
        self.local_qubits = tuple(q for q in range(n) if q not in gq)
        self.segments: List[Segment] = []
        self.nexchanges = 0
        home = [0] * n  # logical qubit -> physical bit of the layout
        for j, q in enumerate(gq):
            home[q] = n - 1 - j
        for k, q in enumerate(self.local_qubits):
            home[q] = nlocal - 1 - k
        self.home = tuple(home)
        pos = list(home)
        # ``final_global_qubits``: leave the state in ANOTHER layout (e.g. run a QFT from the cyclic layout, which needs one
        # exchange per global qubit, and hand the result to the sharded measurement in the block layout it works on)
        fq = gq if final_global_qubits is None else tuple(int(q) for q in final_global_qubits)
        if len(fq) != nglobal or len(set(fq)) != nglobal or any(q < 0 or q >= n for q in fq):
            raise ValueError(f"final_global_qubits must name {nglobal} distinct qubits, got {fq}")
        self.final_global_qubits = fq
        self.final_local_qubits = tuple(q for q in range(n) if q not in fq)
        if fq != gq:
            home = [0] * n
            for j, q in enumerate(fq):
                home[q] = n - 1 - j
            for k, q in enumerate(self.final_local_qubits):
                home[q] = nlocal - 1 - k
        self.final_home = tuple(home)
        # next dense use of every logical qubit, for the furthest-in-future eviction rule
        needs = [mixing_targets(op) for op in ops]
        swaps = [relabel_swaps and is_plain_swap(op) for op in ops]
        next_use = [[] for _ in range(n)]
        for i, (nd, sw) in enumerate(zip(needs, swaps)):
            if sw:
                continue
            for q in nd:
                next_use[q].append(i)
        # label_at_end[i][q]: the label the amplitudes called q before op i carry after the last relabelling SWAP -- the
        # bit they belong on at the end is home[that label].  Walked backwards; only stored where it changes.
        end_label = list(range(n))
        label_changes = {}
        for i in range(len(ops) - 1, -1, -1):
            if swaps[i]:
                a, b = ops[i].targets
                end_label[a], end_label[b] = end_label[b], end_label[a]
                label_changes[i] = list(end_label)
        change_points = sorted(label_changes)

        def end_home(q, i):
            """Physical bit where the amplitudes named q before op i must end up."""
            import bisect

            k = bisect.bisect_left(change_points, i)
            lab = label_changes[change_points[k]][q] if k < len(change_points) else q
            return home[lab]

        ptr = [0] * n
        cur: List[PhysOp] = []
        window = [b for b in range(nlocal - 1, max(nlocal - 1 - high_window, -1), -1)]
        never = len(ops) + 1

        def use_after(q, i):
            while ptr[q] < len(next_use[q]) and next_use[q][ptr[q]] < i:
                ptr[q] += 1
            return next_use[q][ptr[q]] if ptr[q] < len(next_use[q]) else never

        def flush():
            if cur:
                self.segments.append(Segment("local", ops=list(cur)))
                cur.clear()

        def exchange(gbit, lbit):
            flush()
            self.segments.append(Segment("exchange", gbit=gbit, lbit=lbit))
            self.nexchanges += 1
            qa, qb = pos.index(gbit), pos.index(lbit)
            pos[qa], pos[qb] = lbit, gbit

        def evictee(gbit, i, busy, only_finished):
            """Window bit whose qubit leaves for ``gbit``: the one needed furthest in the future; among equals the one
            that belongs on ``gbit`` at the end (saves the exchange that would bring it there later)."""
            best, best_key = None, None
            for b in window:
                e = pos.index(b)
                if e in busy:
                    continue
                use = use_after(e, i)
                if only_finished and use != never:
                    continue
                key = (use, 1 if end_home(e, i) == gbit else 0)
                if best_key is None or key > best_key:
                    best, best_key = b, key
            return best

        for i, op in enumerate(ops):
            if swaps[i]:
                a, b = op.targets
                pos[a], pos[b] = pos[b], pos[a]  # the label "a" now names the amplitudes that were "b"
                continue
            if len(needs[i]) > nlocal:
                raise ValueError("gate has more mixing targets than there are local qubits")
            exchanged = False
            for q in needs[i]:
                if pos[q] < nlocal:
                    continue
                best = evictee(pos[q], i, needs[i], False)
                if best is None:
                    raise ValueError("no local qubit available to exchange with")
                exchange(pos[q], best)
                exchanged = True
            if exchanged and batch_exchanges:
                # the local segment has been cut anyway: bring in every other global qubit that still has a mixing gate
                # ahead, as long as a FINISHED qubit (no mixing gate ahead: it never has to come back) can leave for it.
                # Same number of exchanges, but the gates between them end up in one local segment (one sweep).
                ahead = sorted((use_after(q, i), q) for q in range(n) if pos[q] >= nlocal)
                for use, q in ahead:
                    if use == never:
                        continue
                    best = evictee(pos[q], i, needs[i], True)
                    if best is None:
                        break
                    exchange(pos[q], best)
            cur.append(PhysOp(op.data, tuple(pos[q] for q in op.targets), tuple(pos[q] for q in op.controls), op.is_diagonal))

        # ---- back to the layout: logical qubit q on physical bit home[q]
        def local_swap(b1, b2):
            cur.append(PhysOp(_SWAP, (b1, b2), (), False))
            qa, qb = pos.index(b1), pos.index(b2)
            pos[qa], pos[qb] = b2, b1

        for gbit in range(n - 1, nlocal - 1, -1):  # fix the global bits first
            want = home.index(gbit)  # logical qubit that belongs here
            if pos[want] == gbit:
                continue
            if pos[want] >= nlocal:  # it sits on another global bit: route through a high local bit
                exchange(pos[want], window[0])
            b = pos[want]
            if b not in window:  # bring it to a high local bit first (a local sweep)
                free = next(w for w in window if pos.index(w) != want)
                local_swap(b, free)
                b = free
            exchange(gbit, b)
        for q in range(n):  # then the local permutation, as SWAP gates for the sweep planner
            target = home[q]
            if target < nlocal and pos[q] != target:
                local_swap(pos[q], target)
        flush()
        assert pos == list(home)


def choose_layout(nqubits: int, nglobal: int, ops: Sequence[Op], relabel_swaps: bool = True,
                  final_global_qubits: Optional[Sequence[int]] = None) -> Plan:
    """The plan with the fewest exchanges among the block layout and the two orders of the cyclic one (ties: block) as
    the INITIAL layout; the final one is the same unless ``final_global_qubits`` fixes it."""
    cyc = cyclic_layout(nqubits, nglobal)
    best = None
    for gq in (block_layout(nqubits, nglobal), cyc, cyc[::-1]):
        plan = Plan(nqubits, nglobal, ops, relabel_swaps=relabel_swaps, global_qubits=gq, final_global_qubits=final_global_qubits)
        if best is None or plan.nexchanges < best.nexchanges:
            best = plan
    return best


def specialise(p: PhysOp, nlocal: int, rank_: int) -> Optional[Op]:
    """PhysOp -> the Op this rank applies to its shard (or None if a global control is 0 here)."""
    tb, cb = list(p.tbits), []
    for c in p.cbits:
        if c >= nlocal:
            if not (rank_ >> (c - nlocal)) & 1:
                return None
        else:
            cb.append(c)
    data = p.data
    k = len(tb)
    glob = [i for i in range(k) if tb[i] >= nlocal]
    if glob:
        keep = np.arange(1 << k)
        for i in glob:
            v = (rank_ >> (tb[i] - nlocal)) & 1
            keep = keep[((keep >> (k - 1 - i)) & 1) == v]
        data = data[keep] if p.is_diagonal else data[np.ix_(keep, keep)]
        tb = [b for b in tb if b < nlocal]
        if p.is_diagonal:
            if np.all(data == 1):
                return None
        elif np.array_equal(data, np.eye(len(keep))):
            return None
    to_q = lambda b: nlocal - 1 - b  # noqa: E731
    return Op(np.ascontiguousarray(data), tuple(to_q(b) for b in tb), tuple(to_q(b) for b in cb), is_diagonal=p.is_diagonal)


# ---------------------------------------------------------------------------------------------- execution
def exchange_half(state: torch.Tensor, nlocal: int, gbit: int, lbit: int, staging: List[torch.Tensor], group=None):
    """Swap global physical bit ``gbit`` with local bit ``lbit``: this rank (global bit value b) sends the half of
    its shard with local bit == 1-b to rank ^ (1 << j) and receives the partner's half with local bit == b into the
    same place.  Contiguous chunks, double-buffered staging; NCCL send/recv over NVLink (gloo on CPU)."""
    import torch.distributed as dist

    r = dist.get_rank(group)
    j = gbit - nlocal
    peer = r ^ (1 << j)
    b = (r >> j) & 1
    view = state.view(-1, 2, 1 << lbit)[:, 1 - b, :]  # (nchunks, 2^lbit), rows are contiguous
    nrows, row = view.shape
    cap = staging[0].numel()
    pending = None
    pieces = []
    for i in range(nrows):
        for s in range(0, row, cap):
            pieces.append(view[i, s : s + min(cap, row - s)])
    for k, piece in enumerate(pieces):
        buf = staging[k % 2][: piece.numel()]
        ops = [dist.P2POp(dist.isend, piece, peer, group), dist.P2POp(dist.irecv, buf, peer, group)]
        if r > peer:
            ops.reverse()
        reqs = dist.batch_isend_irecv(ops)
        if pending is not None:
            pending[0].copy_(pending[1])
        for q in reqs:
            q.wait()
        pending = (piece, buf)
    if pending is not None:
        pending[0].copy_(pending[1])
    return 2 * state.element_size() * (state.numel() // 2)  # bytes sent + received by this rank


class PeerShard:
    """This rank's shard in a CUDA-IPC exportable buffer, with every other rank's shard mapped into this process
    (NVLink peer memory).  Lets the exchange run as ONE kernel per rank -- loads and stores on the partner's
    memory -- instead of NCCL send/recv through staging buffers."""

    def __init__(self, engine, nlocal: int, dtype):
        import torch.distributed as dist

        self.engine = engine
        self.rank = dist.get_rank()
        # two exported buffers: the out-of-place permutation kernel (K8) writes into the other one and the shard moves
        # there -- on every rank at the same point of the plan (SWAP runs are never specialised away), so each rank knows
        # which of a peer's two mappings is current without asking
        self.array = engine.malloc_exportable((1 << nlocal,), dtype)
        self.alt = engine.malloc_exportable((1 << nlocal,), dtype)
        self._base = (self.array.data_ptr(), self.alt.data_ptr())
        handles = [None] * dist.get_world_size()
        dist.all_gather_object(handles, (engine.ipc_handle(self.array), engine.ipc_handle(self.alt)))
        self._peer_ptrs = {}
        for r, (h0, h1) in enumerate(handles):
            if r != dist.get_rank():
                self._peer_ptrs[r] = (engine.ipc_open(h0), engine.ipc_open(h1))
        self.flag = torch.zeros(1, device=self.array.tensor.device)

    @property
    def current(self) -> int:
        return self._base.index(self.array.data_ptr())

    @property
    def alt_ptrs(self):
        """rank -> device pointer of that rank's SECOND (not current) buffer in this process, this rank included."""
        other = 1 - self.current
        out = {r: p[other] for r, p in self._peer_ptrs.items()}
        out[self.rank] = self.alt.data_ptr()
        return out

    def flip(self):
        """Continue in the second buffer (every rank does so at the same point of the plan)."""
        self.array.tensor, self.alt.tensor = self.alt.tensor, self.array.tensor
        self.array._owner, self.alt._owner = self.alt._owner, self.array._owner

    @property
    def peer_ptr(self):
        """rank -> device pointer of that rank's CURRENT buffer in this process."""
        cur = self.current
        return {r: p[cur] for r, p in self._peer_ptrs.items()}

    def fence(self):
        """Stream-ordered barrier across ranks: kernels enqueued after it start only when every rank's stream got here."""
        import torch.distributed as dist

        dist.all_reduce(self.flag)

    @property
    def tensor(self):
        return self.array.tensor


class LocalPeerShard(PeerShard):
    """One of W shards that all live on ONE GPU (SingleDeviceGroup): the "peer" pointers are the sibling shards' buffers in
    the same address space, so the real exchange kernels (k7_alltoall_push / k7_alltoall_p2p / k7_swap_half_p2p) and the
    per-rank plans run end to end without NVLink -- the fences are no-ops because the ranks take turns on one stream."""

    def __init__(self, engine, nlocal: int, dtype, rank_: int):
        self.engine, self.rank = engine, rank_
        self.array = engine.empty((1 << nlocal,), dtype)
        self.alt = engine.empty((1 << nlocal,), dtype)
        self.array._owner = self.alt._owner = None  # (plain torch allocations: nothing to hand over when the buffers flip)
        self._base = (self.array.data_ptr(), self.alt.data_ptr())
        self._peer_ptrs = {}

    @staticmethod
    def link(shards):
        for s in shards:
            s._peer_ptrs = {t.rank: t._base for t in shards if t is not s}

    def fence(self):
        pass


class ShardedProgram:
    """A gate queue planned once for (n, world size) and specialised for this rank; ``run`` applies it to a shard."""

    def __init__(self, engine, nqubits: int, dtype, ops: Sequence[Op], relabel_swaps: bool = True, apply=None,
                 staging_elems: int = 1 << 26, global_qubits=None, final_global_qubits=None, world_: Optional[int] = None,
                 rank_: Optional[int] = None):
        """``global_qubits``: None = the leading qubits (block layout: rank r holds state[r * 2^nlocal : (r+1) * 2^nlocal]),
        a sequence of qubits (Plan), or "auto" = whichever of the block / cyclic layouts needs the fewest exchanges.
        ``final_global_qubits``: the layout the state is left in (default: the one it came in); ``shard_of`` / ``scatter``
        / ``basis_state`` / ``locate`` speak the initial layout, ``gather`` / ``canonical_index`` the final one."""
        self.engine = engine
        # (world_, rank_): play rank `rank_` of `world_` without a process group (SingleDeviceGroup, tests)
        self.world = world_size() if world_ is None else int(world_)
        self.rank = rank() if rank_ is None else int(rank_)
        self.g = int(round(math.log2(self.world)))
        if 1 << self.g != self.world:
            raise ValueError("the number of ranks must be a power of two")
        self.n, self.dtype = nqubits, np.dtype(dtype)
        self.nlocal = nqubits - self.g
        if isinstance(global_qubits, str):
            if global_qubits != "auto":
                raise ValueError(f"unknown layout {global_qubits!r}")
            self.plan = choose_layout(nqubits, self.g, ops, relabel_swaps=relabel_swaps, final_global_qubits=final_global_qubits)
        else:
            self.plan = Plan(nqubits, self.g, ops, relabel_swaps=relabel_swaps, global_qubits=global_qubits,
                             final_global_qubits=final_global_qubits)
        self.global_qubits, self.local_qubits = self.plan.global_qubits, self.plan.local_qubits
        self.final_global_qubits, self.final_local_qubits = self.plan.final_global_qubits, self.plan.final_local_qubits
        self.fuse_perm = os.environ.get("QB_A2A_FUSE_PERM", "0") not in ("", "0")  # experimental, see split_trailing_permutation
        # chunk-pipelined exchange (DMA copies over peer memory overlapped with the local sweeps before them)
        self.pipeline = os.environ.get("QB_NO_PIPELINE", "") in ("", "0")
        # the chunk of a pipelined exchange that stays on its rank: swept out of place into the second buffer (its last
        # sweep writes there) instead of being copied first and swept afterwards
        self.local_out_of_place = os.environ.get("QB_NO_LOCAL_OUT_OF_PLACE", "") in ("", "0")
        self.copy_streams = int(os.environ.get("QB_COPY_STREAMS", "4"))
        self._copy_stream = None
        self.segments = []
        runs = exchange_runs(self.plan.segments)
        # OPT-IN (QB_SINK_LOW_QUBITS=5): measured on 2 B200s, QFT(33): 161.8 ms with the five lowest stages sunk behind the
        # exchange, 142.3 ms without -- the exchange phase does not get shorter with fewer stages in its chunk sweeps (it is
        # three passes over the shard next to the DMA traffic, 93 ms either way) and the last sweep, now eight stages with the
        # arrived qubits' ascending-order fans, leaves the stage-only kernel (23 -> 45 ms)
        if self.pipeline:
            runs = sink_behind_exchanges(runs, self.nlocal, int(os.environ.get("QB_SINK_LOW_QUBITS", "0")),
                                         int(os.environ.get("QB_SINK_MAX_BITS", "3")))
        for i, (kind, payload) in enumerate(runs):
            if kind == "local":
                prev = self.segments[-1] if self.segments else None
                if prev is not None and prev[0] == "exchange" and prev[2] is not None:
                    payload = prev[3]  # its closing permutation rides on the exchange before it
                local = [o for o in (specialise(p, self.nlocal, self.rank) for p in payload) if o is not None]
                self.segments.append(("local", local))
            else:
                pairs, sub_dest, gates = list(payload), None, None
                if self.fuse_perm and i + 1 < len(runs) and runs[i + 1][0] == "local" and alltoall_push_entries(0, self.nlocal, pairs):
                    split = split_trailing_permutation(runs[i + 1][1], self.nlocal, len(pairs))
                    if split is not None:
                        gates, sub_dest = split
                piped = None
                if self.pipeline and sub_dest is None and len(pairs) <= 3 and alltoall_push_entries(0, self.nlocal, pairs):
                    # the tail of the local segment in front of the exchange that leaves the k leading local qubits alone
                    # runs chunk by chunk, each chunk leaving as soon as it is done (_pipelined_exchange)
                    piped = PipedOps([], [])
                    if self.segments and self.segments[-1][0] == "local":
                        head, tail = split_for_pipeline(self.nlocal, self.dtype, self.segments[-1][1], len(pairs))
                        if tail is not None:
                            self.segments[-1] = ("local", head)
                            piped = tail
                self.segments.append(("exchange", pairs, sub_dest, gates, piped))
        # runs of >= alltoall_min exchanges go through the all-to-all kernel (peer-memory shards only)
        self.alltoall = os.environ.get("QB_NO_ALLTOALL", "") in ("", "0")
        self.alltoall_min = int(os.environ.get("QB_ALLTOALL_MIN", "1"))
        # out-of-place all-to-all (remote stores only, then flip buffers): 676 vs 653 GB/s per direction on 2 GPUs
        self.alltoall_push = os.environ.get("QB_ALLTOALL_PUSH", "1") not in ("", "0")
        self._apply = apply  # test hook: NumPy shard executor
        self._programs = {}  # id(segment) -> engine.CompiledProgram
        self._staging = None
        self._staging_elems = staging_elems
        self.ngates = len(ops)

    # ---- layout: canonical index <-> (rank, index in the shard) ------------------------------------------
    def locate(self, index: int):
        """-> (rank, shard index) of the amplitude with canonical index ``index`` (qubit 0 = most significant bit)."""
        bit = lambda q: (index >> (self.n - 1 - q)) & 1  # noqa: E731
        r = 0
        for q in self.global_qubits:
            r = (r << 1) | bit(q)
        loc = 0
        for q in self.local_qubits:
            loc = (loc << 1) | bit(q)
        return r, loc

    def canonical_index(self, rank_: int, loc: int) -> int:
        """Canonical index of amplitude ``loc`` of rank ``rank_`` AFTER the program (the final layout; the inverse of
        ``locate`` when the layout does not change)."""
        index = 0
        for j, q in enumerate(self.final_global_qubits):
            index |= ((rank_ >> (self.g - 1 - j)) & 1) << (self.n - 1 - q)
        nl = len(self.final_local_qubits)
        for k, q in enumerate(self.final_local_qubits):
            index |= ((loc >> (nl - 1 - k)) & 1) << (self.n - 1 - q)
        return index

    def _axes(self):
        return list(self.global_qubits) + list(self.local_qubits)

    # ---- shard constructors --------------------------------------------------------------------
    def peer_shard(self, index: Optional[int] = 0):
        """A shard in peer-mapped memory (enables the single-kernel NVLink exchange), initialised to |index>."""
        ps = PeerShard(self.engine, self.nlocal, self.dtype)
        ps.tensor.zero_()
        if index is not None:
            r, loc = self.locate(index)
            if r == self.rank:
                ps.tensor[loc] = 1
        return ps

    def basis_state(self, index: int = 0):
        """|index> of the full register: one amplitude on the rank that ``locate`` names."""
        r, loc = self.locate(index)
        st = self.engine.basis_state(self.nlocal, self.dtype, 0)
        if r != self.rank:
            st.tensor.zero_()
        elif loc:
            st.tensor.zero_()
            st.tensor[loc] = 1
        return st

    def shard_of(self, full: np.ndarray) -> np.ndarray:
        """This rank's shard of a full host state in canonical order (small n)."""
        if self.global_qubits == tuple(range(self.g)):
            lo = self.rank << self.nlocal
            return np.ascontiguousarray(full[lo : lo + (1 << self.nlocal)])
        t = np.asarray(full).reshape((2,) * self.n).transpose(self._axes()).reshape(self.world, 1 << self.nlocal)
        return np.ascontiguousarray(t[self.rank])

    def scatter(self, full: np.ndarray):
        return self.engine.upload(self.shard_of(full).astype(self.dtype))

    # ---- execution ---------------------------------------------------------------------------------
    def _stage(self, tensor):
        if self._staging is None or self._staging[0].device != tensor.device or self._staging[0].dtype != tensor.dtype:
            n = min(self._staging_elems, tensor.numel() // 2)
            self._staging = [torch.empty(n, dtype=tensor.dtype, device=tensor.device) for _ in range(2)]
        return self._staging

    def run(self, state, timed: bool = True, compiled: bool = True):
        """Apply the program to this rank's shard (DeviceArray, or a torch tensor with the test hook).  ``compiled``: the
        local segments run as device-resident compiled programs (planned on first use); False sends the host gate
        matrices through qb_apply_program on every call.  ``timed``: CUDA events are recorded around every segment
        WITHOUT synchronising the host (the step loop never waits for the GPU); ``RunStats.resolve()`` reads them after
        the caller's own synchronisation -- the properties of the returned stats do that on first access."""
        out = RunStats()
        for index in range(len(self.segments)):
            self.run_segment(index, state, out, timed=timed, compiled=compiled)
        return out

    def run_segment(self, index: int, state, out: "RunStats", timed: bool = False, compiled: bool = True, phase: Optional[int] = None):
        """One segment of the plan on this rank's shard.  Every rank runs segment ``index`` before any rank runs
        ``index + 1`` (the fences inside the exchanges enforce it across processes; SingleDeviceGroup does it by taking
        turns).  ``phase``: an exchange segment whose transport cannot pipeline first applies the deferred ops to the whole
        shard (phase 0) and then exchanges (phase 1); None = both.  (Ranks that take turns must all finish phase 0 before
        the first one pulls a sibling's data in phase 1.)"""
        peer = state if isinstance(state, PeerShard) else None
        if peer is not None:
            state = peer.array
        seg = self.segments[index]
        spans = out.spans if timed else None
        tensor = state.tensor if hasattr(state, "tensor") else state  # (a permutation re-points the DeviceArray)
        if seg[0] == "local":
            if seg[1] and phase in (None, 0):
                self._run_local(seg, seg[1], state, peer, out, spans, compiled)
            return
        pairs = seg[1]
        piped = seg[4] if len(seg) > 4 else None
        can_pipeline = piped is not None and self._can_pipeline(peer, pairs)
        if phase in (None, 0) and piped is not None and not can_pipeline and piped.nlocal_ops:
            # transport without chunk pipelining (NCCL staging, in-place kernels, CPU test hook): the deferred ops run on
            # the whole shard
            self._run_local(piped.full_key, piped.nlocal_ops, state, peer, out, None, compiled)
        if phase == 0:
            return
        ev = _span_begin(spans, tensor)
        if can_pipeline:
            if phase not in (None, 1):
                return
            self._pipelined_exchange(state, peer, pairs, piped, out, compiled)
        elif phase is not None and self._is_pairwise(peer, pairs, seg[2]):
            # ranks taking turns: one pairwise exchange per phase (both partners must finish pair k before pair k + 1)
            if phase - 1 < len(pairs):
                self._exchange_run(state, peer, tensor, [pairs[phase - 1]], out, timed=timed, compiled=compiled)
            if phase != len(pairs):
                return
        elif phase in (None, 1):
            self._exchange_run(state, peer, tensor, pairs, out, sub_dest=seg[2], timed=timed, compiled=compiled)
        else:
            return
        out.nexchanges += len(pairs)
        _span_end(spans, ev, "exchange", 1)

    def _is_pairwise(self, peer, pairs, sub_dest) -> bool:
        """True when ``_exchange_run`` would carry the run out as one half-shard swap per pair."""
        if sub_dest is not None:
            return False
        use_a2a = peer is not None and self.alltoall and len(pairs) >= self.alltoall_min
        if use_a2a and self.alltoall_push and len(pairs) <= 3 and alltoall_push_entries(self.rank, self.nlocal, pairs) is not None:
            return False
        return not (use_a2a and alltoall_entries(self.rank, self.nlocal, pairs) is not None)

    def segment_phases(self, index: int) -> int:
        """Phases a group of ranks that take turns must step through for segment ``index`` (run_segment's ``phase``)."""
        seg = self.segments[index]
        return 1 if seg[0] == "local" else 1 + max(1, len(seg[1]))

    def _can_pipeline(self, peer, pairs) -> bool:
        return (peer is not None and self._apply is None and self.alltoall and self.alltoall_push and self.pipeline
                and alltoall_push_entries(self.rank, self.nlocal, pairs) is not None)

    def _run_local(self, key, ops, state, peer, out, spans, compiled):
        tensor = state.tensor if hasattr(state, "tensor") else state
        if self._apply is not None:
            self._apply(tensor, self.nlocal, ops)
            return
        alt = peer.alt if peer is not None else None
        if compiled:
            st = self.engine.run_program(self._compiled(key, ops), state, alt=alt, spans=spans)
        else:
            st = self.engine.apply_program(state, self.nlocal, ops, alt=alt, spans=spans)
        out.nsweeps += st.nsweeps
        out.nperm += getattr(st, "nperm", 0)

    def _compiled(self, key, ops=None, nqubits=None):
        """The local segment compiled for this rank's engine, on first use (qb_program_create): later runs of the plan
        launch kernels only."""
        k = id(key)
        prog = self._programs.get(k)
        if prog is None:
            prog = self._programs[k] = self.engine.compile(self.nlocal if nqubits is None else nqubits, self.dtype,
                                                           key[1] if ops is None else ops)
        return prog

    # ---- the exchange overlapped with the local sweeps before it ----------------------------------------------------
    def _pipelined_exchange(self, state, peer, pairs, piped: "PipedOps", out, compiled) -> bool:
        """A run of exchanges on the k leading local bits, out of place, chunk by chunk (chunk = the k leading bits of the
        shard index, which is also what the all-to-all moves): the gates that precede the exchange and touch none of the
        k leading local qubits (``piped.ops``, already re-indexed to the chunk's nlocal - k qubits) run on one chunk at a
        time on the compute stream; as soon as a chunk is done it leaves for the destination rank's second buffer on the
        copy stream -- a DMA copy over NVLink peer memory that occupies no SM -- while the next chunk is being swept.
        Remote chunks are swept and sent in order of rank XOR distance (every rank then receives from one sender at a
        time); the chunk that stays on this rank is copied first, unswept, while the copy engine is still idle, and swept
        in the second buffer at the end (measured on 2 GPUs, QFT(33): 212.9 ms without the pipeline, 177.5 ms with it and
        the local chunk copied last).  The ranks then continue in their second buffers."""
        entries = alltoall_push_entries(self.rank, self.nlocal, pairs)
        k = len(pairs)
        chunk_n = self.nlocal - k
        sub_bits = getattr(piped, "sub_bits", 0)
        if sub_bits and any(hi - lo != (1 << chunk_n) for _, _, _, lo, hi in entries):
            raise RuntimeError("internal: pieces of a chunk need whole-chunk all-to-all entries")
        sub_n = chunk_n - sub_bits  # qubits of one piece: what piped.ops are indexed for
        npieces, piece_elems = 1 << sub_bits, 1 << sub_n
        elem = state.tensor.element_size()
        eng = self.engine
        main = torch.cuda.current_stream(state.tensor.device)
        if self._copy_stream is None:
            # several copy streams: one DMA copy per stream keeps more than one copy engine busy (8 GPUs, measured:
            # a single stream moved the chunks at 460 GB/s per direction, the push kernel reaches 680)
            self._copy_stream = [torch.cuda.Stream(device=state.tensor.device) for _ in range(max(1, self.copy_streams))]
        copies = self._copy_stream
        prog = None
        copying = None  # the same ops with the last sweep writing into the second buffer: the chunk that stays needs no copy
        if piped.ops:
            prog = self._compiled(piped, piped.ops, nqubits=sub_n) if compiled else None
            if compiled and self.local_out_of_place:
                key = ("copying", id(piped))
                if key not in self._programs:
                    self._programs[key] = eng.compile_copying(sub_n, self.dtype, piped.ops)
                copying = self._programs[key]
        dst = peer.alt_ptrs
        src0 = state.data_ptr()
        peer.fence()  # every rank is done reading what is now its second buffer
        order = sorted(entries, key=lambda e: ((e[0] ^ self.rank) == 0, e[0] ^ self.rank))
        from qibo_b200.array import DeviceArray

        def sweep_chunk(view):
            if not piped.ops:
                return
            st = eng.run_program(prog, view) if prog is not None else eng.apply_program(view, sub_n, piped.ops)
            out.nsweeps += st.nsweeps
            out.nchunk_sweeps += st.nsweeps

        # the chunk that stays on this rank moves FIRST, unswept, while the copy engine has nothing else to do (the remote
        # chunks are still being swept); its gates then run on the copy, in the second buffer, at the end
        def dma(dst_ptr, src_ptr, count, after):
            """One chunk as len(copies) DMA copies, each on its own stream, all after event ``after``."""
            part = -(-count // len(copies))
            for k_, stream in enumerate(copies):
                lo_, hi_ = k_ * part, min(count, (k_ + 1) * part)
                if lo_ >= hi_:
                    break
                stream.wait_event(after)
                eng.memcpy_async(dst_ptr + lo_ * elem, src_ptr + lo_ * elem, (hi_ - lo_) * elem, stream.cuda_stream)

        start = torch.cuda.Event()
        start.record(main)
        stays = None
        for r2, a, b, lo, hi in order:
            if r2 == self.rank:
                stays_at, stays_from = b, a
                stays = []
                if copying is not None:
                    continue  # swept straight into the second buffer at the end
                dma(dst[r2] + b * elem, src0 + a * elem, hi - lo, start)
                for stream in copies:
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    stays.append(ev)
        for r2, a, b, lo, hi in order:
            if r2 == self.rank:
                continue
            for piece in range(npieces):  # piece by piece: swept on the compute stream, then on its way
                off = piece * piece_elems
                sweep_chunk(DeviceArray(state.tensor[a + off : a + off + piece_elems]))
                done = torch.cuda.Event()
                done.record(main)
                dma(dst[r2] + (b + off) * elem, src0 + (a + off) * elem, piece_elems, done)
            out.exchange_bytes += 2 * elem * (hi - lo)
        if stays is not None:
            for ev in stays:
                main.wait_event(ev)
            for piece in range(npieces):
                off = stays_at + piece * piece_elems
                if copying is not None:
                    src_off = stays_from + piece * piece_elems
                    st = eng.run_copying(copying, DeviceArray(state.tensor[src_off : src_off + piece_elems]),
                                         DeviceArray(peer.alt.tensor[off : off + piece_elems]))
                    out.nsweeps += st.nsweeps
                    out.nchunk_sweeps += st.nsweeps
                else:
                    sweep_chunk(DeviceArray(peer.alt.tensor[off : off + piece_elems]))
        for stream in copies:
            landed = torch.cuda.Event()
            landed.record(stream)
            main.wait_event(landed)
        peer.fence()  # every rank's chunks have landed
        peer.flip()
        out.nexchange_launches += len(order)
        out.pipelined += 1
        return True

    def _exchange_run(self, state, peer, tensor, pairs, out, sub_dest=None, timed=False, compiled=True):
        """One run of exchanges [(gbit, lbit), ...] on pairwise distinct bits.  ``sub_dest`` (experimental): a permutation
        of the lower local qubits that the plan moved in front of the gates after the exchange -- written by the
        all-to-all itself when it runs out of place, else applied right after the exchange."""
        use_a2a = peer is not None and self.alltoall and len(pairs) >= self.alltoall_min
        elem = tensor.element_size()
        if sub_dest is not None:
            k = len(pairs)
            entries = alltoall_push_entries(self.rank, self.nlocal, pairs) if (use_a2a and self.alltoall_push and k <= 3) else None
            if entries is not None:
                # one K8 launch per chunk, straight into the destination rank's second buffer; peers in order of r ^ r'
                # so that every rank receives from one sender at a time
                dst, src0, lo = peer.alt_ptrs, state.data_ptr(), self.nlocal - k
                peer.fence()
                for r2, a, b, _, _ in sorted(entries, key=lambda e: e[0] ^ self.rank):
                    self.engine.permute_raw(src0 + a * elem, dst[r2] + b * elem, lo, self.dtype, sub_dest)
                peer.fence()
                peer.flip()
                out.exchange_bytes += 2 * elem * sum(hi - lo_ for r2, _, _, lo_, hi in entries if r2 != self.rank)
                out.nexchange_launches += len(entries)
                return
            # other transports: exchange as usual, then the permutation as a local segment of its own
            self._exchange_run(state, peer, tensor, pairs, out)
            full = list(range(k)) + [k + d for d in sub_dest]
            from qibo_b200.engine import swaps_for_permutation

            swaps = swaps_for_permutation(full)
            if self._apply is not None:
                self._apply(state.tensor if hasattr(state, "tensor") else state, self.nlocal, swaps)
            else:
                st = self.engine.apply_program(state, self.nlocal, swaps, alt=peer.alt if peer is not None else None)
                out.nsweeps += st.nsweeps
                out.nperm += getattr(st, "nperm", 0)
            return
        if use_a2a and self.alltoall_push and len(pairs) <= 3:
            entries = alltoall_push_entries(self.rank, self.nlocal, pairs)
            if entries is not None:
                # out of place: local loads, remote STORES into the destination ranks' second buffers, then every rank
                # continues in its second buffer (the same flip the K8 permutation does)
                dst = peer.alt_ptrs
                peer.fence()
                self.engine.alltoall_p2p(state, [(dst[r2], a, b, lo, hi) for r2, a, b, lo, hi in entries], push=True)
                peer.fence()
                peer.flip()
                out.exchange_bytes += 2 * elem * sum(hi - lo for r2, _, _, lo, hi in entries if r2 != self.rank)
                out.nexchange_launches += 1
                return
        entries = alltoall_entries(self.rank, self.nlocal, pairs) if use_a2a else None
        if entries is not None:
            # ONE kernel for the whole run of exchanges: an all-to-all of contiguous chunks over peer memory
            ptrs = peer.peer_ptr
            peer.fence()
            self.engine.alltoall_p2p(state, [(ptrs[r2], a, b, lo, hi) for r2, a, b, lo, hi in entries])
            peer.fence()
            out.exchange_bytes += 4 * elem * sum(hi - lo for _, _, _, lo, hi in entries)
            out.nexchange_launches += 1
            return
        for gbit, lbit in pairs:
            if peer is not None:
                j = gbit - self.nlocal
                b = (self.rank >> j) & 1
                peer.fence()
                self.engine.swap_half_p2p(state, peer.peer_ptr[self.rank ^ (1 << j)], self.nlocal, self.nlocal - 1 - lbit, b, b, 2)
                peer.fence()
                out.exchange_bytes += elem * tensor.numel()
            else:
                out.exchange_bytes += exchange_half(tensor, self.nlocal, gbit, lbit, self._stage(tensor))
            out.nexchange_launches += 1

    def assemble(self, concatenated: np.ndarray) -> np.ndarray:
        """All shards in rank order -> the full state in canonical order (the FINAL layout of the program)."""
        if self.final_global_qubits == tuple(range(self.g)):
            return concatenated
        inv = np.argsort(list(self.final_global_qubits) + list(self.final_local_qubits))
        return np.ascontiguousarray(concatenated.reshape((2,) * self.n).transpose(inv)).reshape(-1)

    def gather(self, state) -> np.ndarray:
        """Full state in canonical order on every rank (small n only)."""
        import torch.distributed as dist

        tensor = state.tensor if hasattr(state, "tensor") else state
        tensor = tensor.clone()
        parts = [torch.empty_like(tensor) for _ in range(self.world)]
        dist.all_gather(parts, tensor)
        return self.assemble(torch.cat(parts).cpu().numpy())


_SWAP4 = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def sink_behind_exchanges(runs, nlocal: int, low_bits: int, max_mixed_after: int):
    """Plan-level rewrite, before the ops are specialised per rank: for every run of exchanges between two local segments,
    the END of the segment in front -- from the first gate that mixes one of the ``low_bits`` lowest state bits on, all of
    whose mixing gates act on those bits only -- moves behind the exchange when the segment behind mixes at most
    ``max_mixed_after`` qubits (a sharded QFT: the stages on the arrived qubits + the closing permutation: a nearly empty
    sweep whose tile spans the lowest bits anyway).  A gate moved past an exchange is the same gate with the exchanged bits
    relabelled (global bit <-> local bit), so a CU1 between a low qubit and a qubit that was global in front of the
    exchange is a CU1 on two LOCAL bits behind it -- it must not be specialised for the old rank bits.  Five stages less in
    the chunk sweeps of the pipelined exchange (compute-bound 8-stage sweeps, what the exchange phase takes)."""
    if low_bits <= 0:
        return runs
    runs = [(k, list(p)) for k, p in runs]

    def is_swap(o):
        return not o.cbits and len(o.tbits) == 2 and not o.is_diagonal and np.shape(o.data) == (4, 4) and np.array_equal(o.data, _SWAP4)

    for i in range(1, len(runs) - 1):
        if runs[i][0] != "exchange" or runs[i - 1][0] != "local" or runs[i + 1][0] != "local":
            continue
        before, pairs, after = runs[i - 1][1], runs[i][1], runs[i + 1][1]
        mixed_after = {b for o in after if not _is_diagonal_op(o) and not is_swap(o) for b in tuple(o.tbits) + tuple(o.cbits)}
        if len(mixed_after) > max_mixed_after:
            continue
        exchanged = {b for pr in pairs for b in pr}
        start = len(before)
        for j in range(len(before) - 1, -1, -1):
            o = before[j]
            # (a mixing gate must stay on local bits: none of the exchanged ones, whatever the size of the register)
            if not _is_diagonal_op(o) and any(b >= low_bits or b in exchanged for b in tuple(o.tbits) + tuple(o.cbits)):
                break
            start = j
        while start < len(before) and _is_diagonal_op(before[start]):
            start += 1  # leading diagonal gates stay with the gate in front of them
        if start >= len(before):
            continue
        relabel = {}
        for g, l in pairs:
            relabel[g], relabel[l] = l, g
        moved = [PhysOp(o.data, tuple(relabel.get(b, b) for b in o.tbits), tuple(relabel.get(b, b) for b in o.cbits), o.is_diagonal)
                 for o in before[start:]]
        runs[i - 1] = ("local", before[:start])
        runs[i + 1] = ("local", moved + after)
    return runs


class PipedOps:
    """The tail of a local segment that rides on the exchange after it: ``ops`` act on the nlocal - k - sub_bits lower local
    qubits of one piece of a chunk (re-indexed), ``nlocal_ops`` are the same gates in shard numbering (transports without
    pipelining).  ``sub_bits``: the tail touches none of the ``sub_bits`` local qubits below the k leading ones either, so
    every chunk is swept and sent in 2^sub_bits pieces -- the first DMA copy starts after one piece has been swept, not one
    chunk (QFT(33) on 2 GPUs: the ONE remote chunk is half a shard; in one piece its 49 ms copy could only start when all
    of its sweeps were done)."""

    def __init__(self, ops, nlocal_ops, sub_bits=0):
        self.ops, self.nlocal_ops, self.sub_bits = ops, nlocal_ops, sub_bits
        self.full_key = ("full", nlocal_ops)

    def __getitem__(self, i):  # (_compiled keys a segment by identity and reads its ops from [1])
        return (None, self.ops)[i]


def _is_diagonal_op(o: Op) -> bool:
    if o.is_diagonal:
        return True
    d = np.asarray(o.data)
    return d.ndim == 2 and not np.count_nonzero(d - np.diag(np.diagonal(d)))


def split_for_pipeline(nlocal: int, dtype, ops: Sequence[Op], k: int):
    """Split a rank's local segment in front of an exchange on the k leading local bits into (head, tail): ``tail`` = the
    gates that touch none of the k leading local qubits and sit, in the sweep planner's own schedule, in sweeps after the
    last sweep that does -- they can run chunk by chunk.  Both keep program order; the ops that moved commute with what
    they moved past (they belong to different sweeps of a valid schedule).  -> (head ops, PipedOps or None)."""
    from qibo_b200.engine import is_plain_swap as engine_plain_swap
    from qibo_b200.engine import plan_program

    if not ops or k < 1 or nlocal - k < 4:
        return list(ops), None
    if any(engine_plain_swap(o) for o in ops):
        return list(ops), None  # SWAP runs become out-of-place permutations of the whole shard
    touch = [any(q < k for q in tuple(o.targets) + tuple(o.controls)) for o in ops]
    _, sweep_of_op = plan_program(nlocal, dtype, ops)
    last = max((s for s, t in zip(sweep_of_op, touch) if t and s >= 0), default=-1)
    head = [o for o, s in zip(ops, sweep_of_op) if s <= last]
    tail = [o for o, s in zip(ops, sweep_of_op) if s > last]
    if not tail:
        return list(ops), None
    # pieces: as many further leading qubits as the tail leaves alone, down to pieces of 2^QB_PIPE_PIECE_QUBITS amplitudes
    # (default 28: 4 GiB of complex128 -- the sweeps of a piece still fill the GPU many times over)
    first_touched = min((q for o in tail for q in tuple(o.targets) + tuple(o.controls)), default=nlocal)
    piece = int(os.environ.get("QB_PIPE_PIECE_QUBITS", "28"))
    sub = max(0, min(first_touched - k, (nlocal - k) - piece))
    if os.environ.get("QB_NO_PIPE_PIECES", "0") not in ("", "0") or nlocal - k - sub < 4:
        sub = 0
    ks = k + sub
    shifted = [Op(o.data, tuple(q - ks for q in o.targets), tuple(q - ks for q in o.controls), is_diagonal=o.is_diagonal) for o in tail]
    return head, PipedOps(shifted, tail, sub)


def _span_begin(spans, tensor):
    if spans is None or not getattr(tensor, "is_cuda", False):
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


def _span_end(spans, e0, kind, count):
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    spans.append((kind, e0, e1, count))


class RunStats:
    """Counters of one ``ShardedProgram.run`` plus the CUDA-event spans recorded around its segments.  The spans are read
    lazily (``resolve``): nothing in the step loop waits for the GPU."""

    def __init__(self):
        self.nsweeps = 0
        self.nchunk_sweeps = 0  # of which: launches on one chunk of the shard (pipelined exchange)
        self.nperm = 0
        self.nexchanges = 0
        self.nexchange_launches = 0  # kernels / DMA copies (or NCCL rounds): a run of exchanges done as one all-to-all counts once
        self.exchange_bytes = 0
        self.pipelined = 0
        self.spans = []  # (kind, event0, event1, launches): "sweep" | "perm" | "exchange"
        self._ms = None

    def resolve(self):
        if self._ms is None:
            ms = {"sweep": 0.0, "perm": 0.0, "exchange": 0.0}
            for kind, e0, e1, _ in self.spans:
                e1.synchronize()
                ms[kind] += e0.elapsed_time(e1)
            self._ms = ms
        return self._ms

    @property
    def elapsed_ms(self):  # CUDA-event time of the local sweeps incl. the K8 launches
        m = self.resolve()
        return m["sweep"] + m["perm"]

    @property
    def perm_ms(self):
        return self.resolve()["perm"]

    @property
    def exchange_ms(self):  # a pipelined exchange includes the chunk sweeps it overlaps with
        return self.resolve()["exchange"]


class SingleDeviceGroup:
    """All W = 2^g shards of an n-qubit register on ONE GPU, every "rank" with its own plan, specialised ops and compiled
    programs, taking turns segment by segment on one stream.  The exchange kernels get the sibling shards' buffers as
    peer pointers, so the whole multi-GPU data path -- Plan, specialise, K7 all-to-all (out of place / in place), pairwise
    half swaps, the chunk-pipelined DMA exchange, buffer flips -- runs and can be checked on a single-GPU box."""

    def __init__(self, engine, nqubits: int, world: int, dtype, ops: Sequence[Op], **kw):
        self.world, self.n = world, nqubits
        self.programs = [ShardedProgram(engine, nqubits, dtype, ops, world_=world, rank_=r, **kw) for r in range(world)]
        self.nlocal = self.programs[0].nlocal
        self.shards = [LocalPeerShard(engine, self.nlocal, dtype, r) for r in range(world)]
        LocalPeerShard.link(self.shards)

    def configure(self, **flags):
        for p in self.programs:
            for k, v in flags.items():
                if not hasattr(p, k):
                    raise AttributeError(k)
                setattr(p, k, v)

    def scatter(self, full: np.ndarray):
        for p, s in zip(self.programs, self.shards):
            s.tensor.copy_(torch.from_numpy(p.shard_of(full).astype(p.dtype)))

    def run(self, compiled: bool = True):
        stats = [RunStats() for _ in self.programs]
        nseg = len(self.programs[0].segments)
        assert all(len(p.segments) == nseg for p in self.programs)
        for index in range(nseg):
            for phase in range(max(p.segment_phases(index) for p in self.programs)):
                for p, s, st in zip(self.programs, self.shards, stats):
                    p.run_segment(index, s, st, compiled=compiled, phase=phase)
        return stats

    def gather(self) -> np.ndarray:
        torch.cuda.synchronize()
        return self.programs[0].assemble(torch.cat([s.tensor for s in self.shards]).cpu().numpy())


def api_layout(nqubits: int, gather_max: int, has_initial_state: bool, world: int, has_special: bool = False) -> dict:
    """Layout arguments of the ShardedProgram behind ``execute_distributed_circuit``: a state that is gathered may use
    any layout (the one with the fewest exchanges); a larger register goes to the sharded measurement, which works on the
    block layout -- it starts in the cheapest layout when it starts from |0...0> (which looks the same in all of them) and
    is handed over in the block layout; a user-supplied initial state is scattered in the block layout.  Circuits with
    special gates (callbacks, collapsing measurements) keep the block layout between their gate runs: the sharded
    reductions of dist_measure.py work on it."""
    if has_special:
        return {}
    if nqubits <= gather_max:
        return dict(global_qubits="auto")
    if not has_initial_state:
        return dict(global_qubits="auto", final_global_qubits=block_layout(nqubits, int(round(math.log2(world)))))
    return {}


class ShardedState:
    """Handle on a state that is too large to replicate (more than QB_GATHER_MAX_QUBITS qubits): every rank holds its
    2^(n-g) amplitudes of the block layout in ``tensor`` (rank r = the g leading qubits).  What ``QuantumState`` offers
    on a full state (result.py:31-163) is answered by reductions across the ranks instead."""

    def __init__(self, backend, shard, nqubits: int):
        from qibo_b200.dist_measure import EngineLocal, ShardMeasure

        self.backend, self.shard, self.nqubits = backend, shard, nqubits
        self.measure = ShardMeasure(nqubits, EngineLocal(backend.engine_gpu))
        self.rank, self.world = self.measure.rank, self.measure.world

    @property
    def tensor(self):
        return self.shard.tensor

    def probabilities(self, qubits=None):
        """Marginal over ``qubits`` in the caller's order, replicated on every rank (DeviceArray); all qubits in order
        (the default): this rank's slice of the 2^n probabilities."""
        from qibo_b200.array import DeviceArray

        qubits = list(range(self.nqubits)) if qubits is None else [int(q) for q in qubits]
        probs, _ = self.measure.probabilities(self.shard, qubits)
        return DeviceArray(probs)

    def samples(self, nshots: int, qubits=None, binary: bool = True):
        """Shots over ``qubits`` (default: all) drawn from the global legacy RNG, identical on every rank (which must
        share the seed, ``Backend.set_seed``)."""
        qubits = list(range(self.nqubits)) if qubits is None else [int(q) for q in qubits]
        probs, sharded = self.measure.probabilities(self.shard, qubits)
        uniforms = np.random.random_sample(int(nshots))
        out = self.measure.sample(probs, uniforms, sharded).cpu().numpy()
        return self.backend.samples_to_binary(out, len(qubits)) if binary else out

    def norm(self) -> float:
        return float(np.sqrt(self.measure.global_norm2(self.shard)))

    def state(self, numpy: bool = False):
        from qibo.config import raise_error

        raise_error(NotImplementedError, f"a {self.nqubits}-qubit state is not replicated on every rank; use .tensor (this rank's "
                    "shard), .probabilities(qubits) or .samples(nshots)")

    def __repr__(self):
        return f"ShardedState(nqubits={self.nqubits}, rank={self.rank}/{self.world}, shard={tuple(self.shard.tensor.shape)})"


def _apply_special(backend, gate, shard, n, measure, gather_max):
    """Gates that must see the whole state (distcircuit.py:278-284): collapsing measurements and callbacks, on the sharded
    state in the block layout.  M(collapse=True) -> marginal (all-reduce), one shot from the global RNG (same seed on
    every rank), projection + global renormalisation (gates/measurements.py:189-205).  Callbacks: Norm and Overlap are
    reductions over the ranks; every other callback receives the gathered state while it fits."""
    from qibo.config import raise_error

    from qibo_b200.array import DeviceArray

    name = gate.__class__.__name__
    if name == "M":
        gate.result.backend = backend
        if not gate.collapse:
            return
        qubits = sorted(gate.target_qubits)
        probs, _ = measure.probabilities(shard, qubits)
        shot = gate.result.add_shot(DeviceArray(probs), backend=backend)
        measure.collapse(shard, qubits, int(np.asarray(shot).ravel()[0]))
        return
    if name == "CallbackGate":
        cb = gate.callback
        cb.nqubits = n
        kind = cb.__class__.__name__
        if kind == "Norm":
            cb.append(float(np.sqrt(measure.global_norm2(shard))))
            return
        if kind == "Overlap":
            target = backend.to_numpy(cb.state) if isinstance(cb.state, DeviceArray) else np.asarray(cb.state)
            nl = measure.nlocal
            mine = backend.engine_gpu.upload(np.ascontiguousarray(target[measure.rank << nl : (measure.rank + 1) << nl]).astype(shard.dtype))
            cb.append(measure.global_vdot(mine, shard))  # <target|state>, complex, as overlap_statevector returns it (abstract.py:2180-2190)
            return
        if n > gather_max:
            raise_error(NotImplementedError, f"callback {kind} needs the full state, which is not gathered above {gather_max} qubits")
        full = backend.engine_gpu.upload(measure.gather(shard))
        gate.apply(backend, full, n)
        return
    raise_error(NotImplementedError, f"{name} is not supported inside distributed circuits")


def execute_circuit(backend, circuit, initial_state=None, nshots=None, return_state: bool = False):
    """``Backend.execute_distributed_circuit`` under torchrun: every rank calls it with the same circuit (and the same
    RNG seed).  The queue is cut at the special gates (distcircuit.py:278-284: callbacks and collapsing measurements see
    the full state); each run of plain gates is one ShardedProgram.  ``return_state``: the gathered final state itself
    (what the per-shot loop of execute_circuit_repeated asks for, abstract.py:2579-2582)."""
    from qibo.config import raise_error
    from qibo.result import CircuitResult, MeasurementOutcomes, QuantumState

    from qibo_b200.array import DeviceArray
    from qibo_b200.dist_measure import EngineLocal, ShardMeasure

    n = circuit.nqubits
    runs, cur = [], []  # [("ops", [Op]) | ("special", gate)]
    for gate in circuit.queue:
        if backend._is_plain(gate):
            cur.extend(backend._gate_ops(gate, n))
            continue
        if gate.__class__.__name__ == "M" and not gate.collapse:
            gate.result.backend = backend
            continue
        if cur:
            runs.append(("ops", cur))
            cur = []
        runs.append(("special", gate))
    if cur:
        runs.append(("ops", cur))
    has_special = any(kind == "special" for kind, _ in runs)
    gather_max = int(os.environ.get("QB_GATHER_MAX_QUBITS", 30))
    eng = backend.engine_gpu
    layout = api_layout(n, gather_max, initial_state is not None, world_size(), has_special)
    measure = ShardMeasure(n, EngineLocal(eng))
    shard, prog = None, None
    first = True
    for kind, payload in runs or [("ops", [])]:
        if kind == "ops":
            prog = ShardedProgram(eng, n, backend._cdtype, payload, **layout)
            if first:
                if initial_state is None:
                    shard = prog.basis_state(0)
                else:
                    host = backend.to_numpy(initial_state) if isinstance(initial_state, DeviceArray) else np.asarray(initial_state)
                    shard = prog.scatter(host)
            prog.run(shard, timed=False)
        else:
            if first:
                prog = ShardedProgram(eng, n, backend._cdtype, [], **layout)
                if initial_state is None:
                    shard = prog.basis_state(0)
                else:
                    host = backend.to_numpy(initial_state) if isinstance(initial_state, DeviceArray) else np.asarray(initial_state)
                    shard = prog.scatter(host)
            _apply_special(backend, payload, shard, n, measure, gather_max)
        first = False
    if return_state:
        if n > gather_max:
            raise_error(NotImplementedError, f"per-shot re-execution gathers the state: not above {gather_max} qubits")
        return eng.upload(prog.gather(shard))
    if n > gather_max:
        # too large to replicate: measurement outcomes come from the sharded state (dist_measure.py) -- marginal over the
        # measured qubits (all-reduce), then inverse-CDF sampling with the global legacy RNG as sample_shots does
        # (abstract.py:2774-2781); without measurements the caller gets a handle on the sharded state.
        if not circuit.measurements:
            circuit._final_state = ShardedState(backend, shard, n)
            return circuit._final_state
        qubits = []
        for m in circuit.measurements:  # order of the joint measurement gate: the qubits as they were added (result.py:444-458)
            qubits.extend(q for q in m.target_qubits if q not in qubits)
        probs, sharded = measure.probabilities(shard, qubits)
        nshots = 1000 if nshots is None else nshots
        uniforms = np.random.random_sample(nshots)  # every rank must hold the same seed (Backend.set_seed)
        samples = measure.sample(probs, uniforms, sharded).cpu().numpy()
        binary = backend.samples_to_binary(samples, len(qubits))
        circuit._final_state = MeasurementOutcomes(circuit.measurements, backend=backend, samples=binary, nshots=nshots)
        return circuit._final_state
    full = eng.upload(prog.gather(shard))
    if circuit.measurements:
        circuit._final_state = CircuitResult(full, circuit.measurements, backend=backend, nshots=1000 if nshots is None else nshots)
    else:
        circuit._final_state = QuantumState(full, backend=backend)
    return circuit._final_state
