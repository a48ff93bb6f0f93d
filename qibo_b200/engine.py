"""Host-side engine over the C ABI: the qibo-independent half of the backend.

Owns one library context per CUDA device and exposes the hot path on :class:`DeviceArray` states.
torch is used for what the task statement calls plumbing: device allocation (the caching allocator)
and the current stream.  All arithmetic happens inside libqibo_b200.so.
"""

import ctypes
from collections import Counter
from typing import Optional, Sequence

import numpy as np
import torch

from qibo_b200 import _lib
from qibo_b200.array import DeviceArray, torch_dtype
from qibo_b200.ops import Op, is_wide, pack_ops

_DT = {np.dtype("complex64"): _lib.QB_C64, np.dtype("complex128"): _lib.QB_C128}
_RT = {np.dtype("float32"): _lib.QB_F32, np.dtype("float64"): _lib.QB_F64}
import os

# NumPy's exact-scan contract (sequential float64 cumsum, bit-identical to np.random.choice) is kept up to this many
# bins; beyond it the parallel scan is used unless the caller opts in (Engine.exact_scan_max_bins, env
# QB_EXACT_SCAN_MAX_BINS, B200Backend.exact_sampling): the sequential scan costs ~4 ns per bin (seconds at 2^30 bins)
EXACT_SCAN_MAX_BINS = int(os.environ.get("QB_EXACT_SCAN_MAX_BINS", 1 << 22))


def _int_array(values):
    arr = (ctypes.c_int * max(len(values), 1))(*[int(v) for v in values])
    return arr


class Engine:
    """One context = one GPU + one stream.  Raises if there is no CUDA device (no CPU fallback)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.QiboB200Error(
                "qibo_b200 needs a CUDA device: the hot path exists only as sm_100a kernels (no CPU fallback)"
            )
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)  # make sure the primary context exists before the library attaches
        # torch's current stream; its default stream reports handle 0, for which the CUDA runtime's explicit
        # name is cudaStreamLegacy (0x1) -- NULL would ask the library for a private non-blocking stream,
        # unordered with torch's copies and clones of the same buffers.
        stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        handle = ctypes.c_void_p()
        _lib.check(self.lib.qb_create(self.device_index, ctypes.c_void_p(stream), ctypes.byref(handle)))
        self.handle = handle
        self._stream = stream
        self._freeze_imports()
        self.last_stats = None
        self.permute_swap_runs = True
        # a run of SWAPs behind a run of gates rides on the LAST sweep of those gates (qb_apply_program_permuted): the sweep
        # writes its tiles permuted, out of place -- no separate K8 pass over the state
        self.fuse_permutations = os.environ.get("QB_NO_FUSE_PERM", "0") in ("", "0")
        self.exact_scan_max_bins = EXACT_SCAN_MAX_BINS

    def __reduce__(self):
        # (pickled with the backend that owns it -- joblib's process pools in qibo/parallel.py, MeasurementResult symbols:
        # a library handle means nothing in another process: the copy opens its own context on the same device index)
        return (Engine, (self.device.index,))

    def bind_current_stream(self):
        """Follow torch's current stream (``with torch.cuda.stream(s):``): the library's kernels and torch's copies /
        allocations of the same buffers must be ordered on ONE stream.  Called at the top of every entry point; a no-op
        unless the current stream changed since the last call."""
        stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        if stream != self._stream:
            _lib.check(self.lib.qb_set_stream(self.handle, ctypes.c_void_p(stream)))
            self._stream = stream

    def close(self):
        if getattr(self, "handle", None):
            self.lib.qb_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- memory / state construction (K6) ---------------------------------------------------------
    def empty(self, shape, dtype) -> DeviceArray:
        return DeviceArray(self._alloc(shape, torch_dtype(dtype)))

    _gc_frozen = False

    @classmethod
    def _freeze_imports(cls, again: bool = False):
        """Once per process, when the first engine is created (i.e. right after the imports, before any circuit of this
        backend exists): move the long-lived module objects (qibo, sympy, torch: ~3e5 tracked objects) out of the cyclic
        collector's scans, so that the collections ``_alloc`` triggers under memory pressure cost < 1 ms instead of
        ~170 ms.  Objects created later are collected as usual.  QB_NO_GC_FREEZE=1 leaves the collector alone.  ``again``:
        the backend calls this once more when it is constructed -- a process that made an Engine BEFORE importing qibo
        (bench.py does) would otherwise scan qibo's and sympy's import-time objects on every collection (measured: the
        plugin-level QFT(32) step 167 ms instead of 138 ms)."""
        if (cls._gc_frozen and not again) or os.environ.get("QB_NO_GC_FREEZE", "") not in ("", "0"):
            return
        import gc

        gc.collect()
        gc.freeze()
        cls._gc_frozen = True

    def _alloc(self, shape, tdtype) -> torch.Tensor:
        """torch.empty with one precaution for states that fill a large part of the device: the reference keeps a finished
        Circuit (and through ``_final_state`` its 2^n amplitudes) alive in a reference cycle (Circuit -> M gate ->
        MeasurementResult -> Circuit), so the buffer of the PREVIOUS execution is only returned by Python's cyclic
        collector.  Before a large allocation that would not fit, collect."""
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=tdtype).element_size()
        if nbytes >= (1 << 30):
            # the allocator's own counters first: when its cache can serve the request nothing else is asked.  (cudaMemGetInfo
            # costs 0.5-1.5 ms as a rule but was seen to take 30-100 ms at random in the plugin-level QFT(32) loop --
            # profiles/r2s_e2e_profile2.txt --, where the second 64 GiB buffer of a step used to take this branch because of
            # a safety margin on top of an exactly fitting cached block.)
            cached = torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
            if cached < nbytes:
                free, _ = torch.cuda.mem_get_info(self.device)
                if free < nbytes + (1 << 28):
                    self._reclaim()
        try:
            return torch.empty(shape, dtype=tdtype, device=self.device)
        except torch.cuda.OutOfMemoryError:
            self._reclaim()
            torch.cuda.empty_cache()
            return torch.empty(shape, dtype=tdtype, device=self.device)

    @staticmethod
    def _reclaim():
        import gc

        gc.collect()

    def reclaim_before(self, nbytes: int):
        """Called by the backend before it builds a state of ``nbytes``: for states that fill a large part of the device,
        collect the PREVIOUS execution's cyclic garbage now (< 2 ms once the imports are frozen) -- its state buffer goes
        back to the allocator's cache and its compiled program is destroyed while the stream is idle -- instead of in the
        middle of this execution, where whether the buffers fit depended on when Python's collector last ran (measured:
        plugin-level QFT(32) steps of 145 or 215 ms at random)."""
        if nbytes >= (1 << 30):
            self._reclaim()

    def basis_state(self, nqubits: int, dtype="complex128", index: int = 0) -> DeviceArray:
        """zero_state (abstract.py:2243-2273) generalised to any basis index."""
        out = self.empty((1 << nqubits,), dtype)
        _lib.check(self.lib.qb_state_set_basis(self.handle, out.data_ptr(), nqubits, _DT[np.dtype(dtype)], index))
        return out

    def uninitialised_state(self, nqubits: int, dtype="complex128") -> DeviceArray:
        """A state buffer with NOTHING written to it, for ``run_program(..., input_zero=True)`` / ``apply_program(...,
        input_zero=True)``: the first sweep of the program makes the |0...0> tiles itself."""
        return self.empty((1 << nqubits,), dtype)

    def _write_zero_state(self, state: DeviceArray, nqubits: int):
        _lib.check(self.lib.qb_state_set_basis(self.handle, state.data_ptr(), nqubits, _DT[np.dtype(state.dtype)], 0))

    def filled_state(self, nqubits: int, value: complex, dtype="complex128") -> DeviceArray:
        """plus_state (abstract.py:2199-2221): every amplitude equal to ``value``."""
        out = self.empty((1 << nqubits,), dtype)
        value = complex(value)
        _lib.check(self.lib.qb_state_fill(self.handle, out.data_ptr(), nqubits, _DT[np.dtype(dtype)], value.real, value.imag))
        return out

    def upload(self, array, dtype=None) -> DeviceArray:
        """Host (or torch) array -> DeviceArray (Backend.cast for a state)."""
        if isinstance(array, DeviceArray):
            return array if dtype is None or np.dtype(dtype) == array.dtype else self.cast(array, dtype)
        if isinstance(array, torch.Tensor):
            t = array.to(self.device)
            return DeviceArray(t if dtype is None else t.to(torch_dtype(dtype)))
        host = np.ascontiguousarray(array) if dtype is None else np.ascontiguousarray(array, dtype=dtype)
        return DeviceArray(torch.from_numpy(host).to(self.device))

    def cast(self, array: DeviceArray, dtype) -> DeviceArray:
        dtype = np.dtype(dtype)
        if dtype == array.dtype:
            return array
        if array.dtype in _DT and dtype in _DT:
            out = self.empty(array.shape, dtype)
            _lib.check(
                self.lib.qb_state_cast(self.handle, out.data_ptr(), _DT[dtype], array.data_ptr(), _DT[array.dtype], array.size)
            )
            return out
        return DeviceArray(array.tensor.to(torch_dtype(dtype)))

    def copy(self, array: DeviceArray) -> DeviceArray:
        return DeviceArray(array.tensor.clone())

    def synchronize(self):
        _lib.check(self.lib.qb_sync(self.handle))

    def norm2(self, state: DeviceArray) -> float:
        n = int(state.size).bit_length() - 1
        out = ctypes.c_double()
        _lib.check(self.lib.qb_state_norm2(self.handle, state.data_ptr(), n, _DT[state.dtype], ctypes.byref(out)))
        return out.value

    # ---- K1: one gate ---------------------------------------------------------------------------
    def apply_op(self, state: DeviceArray, nqubits: int, op: Op) -> DeviceArray:
        self.bind_current_stream()
        if is_wide(op):
            return self.apply_wide(state, nqubits, op)
        if len(op.targets) > 5 and not op.is_diagonal:
            # the one-gate kernels hold 2^k amplitudes per thread (k <= 5); 6-qubit blocks go through the sweep kernel
            self._apply_sweeps(state, nqubits, [op], True, False)
            return state
        fn = self.lib.qb_apply_diagonal if op.is_diagonal else self.lib.qb_apply_matrix
        _lib.check(
            fn(
                self.handle, state.data_ptr(), nqubits, _DT[state.dtype], op.data.ctypes.data,
                len(op.targets), _int_array(op.targets), len(op.controls), _int_array(op.controls),
            )
        )
        return state

    def apply_wide(self, state: DeviceArray, nqubits: int, op: Op) -> DeviceArray:
        """A block on k > 6 targets (the reference takes a Unitary of any width, gates/gates.py:2774): 2^k complex MACs
        per amplitude make it GEMM-shaped, so the targets (then the controls) are brought to the leading qubits with one
        K8 permutation sweep, the matrix multiplies the state viewed as a (2^k, 2^(n-k)) matrix -- a plain library GEMM
        (cuBLAS through torch.matmul) on the slice where all controls are 1 -- and a second K8 sweep restores the order."""
        k, c = len(op.targets), len(op.controls)
        mat = np.diag(op.data) if op.is_diagonal else op.data
        if c == 0 and np.array_equal(mat, np.eye(1 << k)):
            return state  # gates.I(*range(7)) and friends
        lead = list(op.targets) + list(op.controls)
        if len(set(lead)) != k + c or any(q < 0 or q >= nqubits for q in lead):
            raise ValueError("bad target / control qubits")
        rest = [q for q in range(nqubits) if q not in lead]
        order = lead + rest  # order[p] = the qubit that moves to position p
        dest = [0] * nqubits
        for p, q in enumerate(order):
            dest[q] = p
        moved = dest != list(range(nqubits))
        if moved:
            self.permute_qubits(state, nqubits, dest)
        u = torch.from_numpy(np.ascontiguousarray(mat)).to(self.device).to(state.tensor.dtype)
        view = state.tensor.view(1 << k, 1 << c, -1)
        if c == 0:
            state.tensor = torch.matmul(u, view[:, 0, :]).reshape(-1)
        else:
            view[:, -1, :] = torch.matmul(u, view[:, -1, :])
        if moved:
            self.permute_qubits(state, nqubits, order)  # position p goes back to qubit order[p]
        return state

    # ---- K2: a gate queue -------------------------------------------------------------------------
    def _apply_sweeps(self, state: DeviceArray, nqubits: int, ops: Sequence[Op], fuse: bool, timed: bool, extra_flags: int = 0):
        self.bind_current_stream()
        stats = _lib.QbProgramStats()
        if len(ops) == 0:
            if extra_flags & _lib.QB_PROGRAM_INPUT_ZERO:
                self._write_zero_state(state, nqubits)
            return stats
        arr, keep = pack_ops(ops)
        flags = (0 if fuse else _lib.QB_PROGRAM_NO_FUSE) | (_lib.QB_PROGRAM_TIME if timed else 0) | extra_flags
        _lib.check(
            self.lib.qb_apply_program(
                self.handle, state.data_ptr(), nqubits, _DT[state.dtype], arr, len(ops), flags, ctypes.byref(stats)
            )
        )
        del keep
        return stats

    def permute_qubits(self, state: DeviceArray, nqubits: int, dest_of_qubit: Sequence[int], timed: bool = False, alt: Optional[DeviceArray] = None,
                       spans: Optional[list] = None):
        """K8: out-of-place qubit permutation in one sweep; the DeviceArray is re-pointed at the result buffer.  ``alt``:
        a second buffer of the same shape to permute into -- the two DeviceArrays then trade buffers (shards that other
        ranks have mapped through CUDA IPC ping-pong between two exported buffers this way); an IPC-exported buffer
        without ``alt`` gets its result copied back in place.  Returns elapsed ms or None."""
        self.bind_current_stream()
        scratch = self._alloc(tuple(state.tensor.shape), state.tensor.dtype) if alt is None else alt.tensor
        if timed or spans is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(
            self.lib.qb_permute_qubits(
                self.handle, state.data_ptr(), scratch.data_ptr(), nqubits, _DT[state.dtype], _int_array(dest_of_qubit)
            )
        )
        self._adopt_result(state, scratch, alt)
        if spans is not None:
            e1.record()
            spans.append(("perm", e0, e1, 1))
        elif timed:
            e1.record()
            e1.synchronize()
            return e0.elapsed_time(e1)
        return None

    @staticmethod
    def _adopt_result(state: DeviceArray, scratch: torch.Tensor, alt: Optional[DeviceArray]):
        """After an out-of-place launch wrote its result to ``scratch`` (= ``alt.tensor`` when given): re-point ``state``."""
        if alt is not None:
            state.tensor, alt.tensor = alt.tensor, state.tensor
            so, ao = getattr(state, "_owner", None), getattr(alt, "_owner", None)
            state._owner, alt._owner = ao, so
        elif getattr(state, "_owner", None) is not None:
            state.tensor.copy_(scratch)
        else:
            state.tensor = scratch

    def _apply_sweeps_permuted(self, state: DeviceArray, nqubits: int, ops: Sequence[Op], dest_of_qubit: Sequence[int], fuse: bool,
                               timed: bool, alt: Optional[DeviceArray], extra_flags: int = 0):
        """``ops`` then the qubit permutation, the permutation riding on the last sweep (qb_apply_program_permuted with
        QB_PROGRAM_PERM_FUSED_ONLY: NotImplementedError -- nothing launched -- when it cannot; OutOfMemoryError when no
        second buffer fits).  The DeviceArray is re-pointed at the result buffer as in ``permute_qubits``."""
        self.bind_current_stream()
        scratch = self._alloc(tuple(state.tensor.shape), state.tensor.dtype) if alt is None else alt.tensor
        stats = _lib.QbProgramStats()
        arr, keep = pack_ops(ops)
        flags = (0 if fuse else _lib.QB_PROGRAM_NO_FUSE) | (_lib.QB_PROGRAM_TIME if timed else 0) | _lib.QB_PROGRAM_PERM_FUSED_ONLY | extra_flags
        _lib.check(
            self.lib.qb_apply_program_permuted(
                self.handle, state.data_ptr(), scratch.data_ptr(), nqubits, _DT[state.dtype], arr, len(ops), _int_array(dest_of_qubit),
                flags, ctypes.byref(stats)
            )
        )
        del keep
        self._adopt_result(state, scratch, alt)
        return stats

    def permute_raw(self, src_ptr: int, dst_ptr: int, nqubits: int, dtype, dest_of_qubit: Sequence[int]):
        """K8 on raw device pointers (a chunk of a shard as source, possibly a peer-mapped buffer as destination)."""
        _lib.check(
            self.lib.qb_permute_qubits(
                self.handle, ctypes.c_void_p(src_ptr), ctypes.c_void_p(dst_ptr), nqubits, _DT[np.dtype(dtype)], _int_array(dest_of_qubit)
            )
        )

    def apply_program(self, state: DeviceArray, nqubits: int, ops: Sequence[Op], fuse: bool = True, timed: bool = False,
                      alt: Optional[DeviceArray] = None, spans: Optional[list] = None, input_zero: bool = False):
        """Apply ``ops`` in order, several gates per HBM sweep.  Runs of >= MIN_SWAP_RUN uncontrolled SWAP gates (the
        bit reversal ending a QFT) become ONE out-of-place permutation sweep when a scratch buffer fits in memory.
        Returns the planner/timing statistics.  ``timed``: CUDA-event time per segment, read back at once (the host waits
        for the GPU); ``spans``: a list that receives (kind, event0, event1, launches) per segment instead -- nothing
        waits, the caller reads the events after its own synchronisation."""
        total = _lib.QbProgramStats()
        total.nops = len(ops)
        total.perm_ms, total.nperm = 0.0, 0  # K8 launches inside this program (reported apart from the sweep kernel)
        segments = split_segments(ops, nqubits, fuse and self.permute_swap_runs, fuse and self.fuse_permutations)
        total.nperm_fused = 0
        # ``input_zero``: ``state`` is uninitialised memory standing for |0...0>; the first sweep program makes it
        # (QB_PROGRAM_INPUT_ZERO), anything else that comes first needs it written
        zero_flag = _lib.QB_PROGRAM_INPUT_ZERO if input_zero else 0
        if zero_flag and (not segments or segments[0][0] not in ("ops", "opsperm")):
            self._write_zero_state(state, nqubits)
            zero_flag = 0
        while segments:
            kind, payload = segments.pop(0)
            if kind == "wide":
                self.apply_wide(state, nqubits, payload)
                total.nsweeps += 3
                total.bytes_moved += 6.0 * state.nbytes
                continue
            if kind == "opsperm":
                gops, dest = payload
                e0 = _record_event() if spans is not None else None
                try:
                    st = self._apply_sweeps_permuted(state, nqubits, gops, dest, fuse, timed, alt, zero_flag)
                except (torch.cuda.OutOfMemoryError, NotImplementedError):
                    segments[:0] = ([("ops", gops)] if gops else []) + [("perm", dest)]  # the two-launch form
                    if zero_flag and not gops:
                        self._write_zero_state(state, nqubits)
                        zero_flag = 0
                    continue
                zero_flag = 0
                if e0 is not None:
                    spans.append(("sweep", e0, _record_event(), st.nsweeps))
                total.nperm_fused += 1
                kind = "done"
            if kind == "perm":
                try:
                    ms = self.permute_qubits(state, nqubits, payload, timed=timed, alt=alt, spans=spans)
                except torch.cuda.OutOfMemoryError:
                    ms = None
                    payload = swaps_for_permutation(payload)
                    kind = "ops"
                else:
                    total.nsweeps += 1
                    total.ndense_passes += 1
                    total.bytes_moved += 2.0 * state.nbytes
                    total.elapsed_ms += ms or 0.0
                    total.perm_ms += ms or 0.0
                    total.nperm += 1
                    continue
            if kind != "done":
                e0 = _record_event() if spans is not None else None
                st = self._apply_sweeps(state, nqubits, payload, fuse, timed, zero_flag)
                zero_flag = 0
                if e0 is not None:
                    spans.append(("sweep", e0, _record_event(), st.nsweeps))
            total.nsweeps += st.nsweeps
            total.ndense_passes += st.ndense_passes
            total.ndiag_ops += st.ndiag_ops
            total.nstage_sweeps += st.nstage_sweeps
            total.bytes_moved += st.bytes_moved
            total.elapsed_ms += st.elapsed_ms
        self.last_stats = total
        return total

    def plan(self, nqubits: int, dtype, ops: Sequence[Op], fuse: bool = True):
        return plan_program(nqubits, dtype, ops, fuse)

    # ---- compiled programs: plan once, launch many times ------------------------------------------------
    def compile(self, nqubits: int, dtype, ops: Sequence[Op], fuse: bool = True) -> "CompiledProgram":
        """The queue ``apply_program`` would run, planned once and kept resident on the device (qb_program_create):
        ``run_program`` then costs kernel launches only -- no canonicalisation, planning or program upload per call."""
        return CompiledProgram(self, nqubits, dtype, ops, fuse)

    # ---- out-of-place programs: the result lands in ANOTHER buffer (the identity "permutation" rides on the last sweep)
    def compile_copying(self, nqubits: int, dtype, ops: Sequence[Op]):
        """``ops`` compiled so that the LAST sweep writes its tiles to another buffer (qb_program_create_permuted with the
        identity permutation): a program that has to end up elsewhere anyway -- the chunk of a pipelined exchange that stays
        on its rank -- saves the copy.  None when the queue is not one sweep program or the planner cannot fuse."""
        segments = split_segments(ops, nqubits, True, False)
        if len(segments) != 1 or segments[0][0] != "ops" or not self.fuse_permutations:
            return None
        arr, keep = pack_ops(ops)
        handle, st = ctypes.c_void_p(), _lib.QbProgramStats()
        try:
            _lib.check(self.lib.qb_program_create_permuted(
                self.handle, nqubits, _DT[np.dtype(dtype)], arr, len(ops), _int_array(list(range(nqubits))),
                _lib.QB_PROGRAM_PERM_FUSED_ONLY, ctypes.byref(handle), ctypes.byref(st)))
        except NotImplementedError:
            return None
        del keep
        return _CopyingProgram(self, handle, st.nsweeps)

    def run_copying(self, prog: "_CopyingProgram", src: DeviceArray, dst: DeviceArray):
        """Apply a ``compile_copying`` program to ``src``; the result is in ``dst`` (``src`` holds an intermediate state)."""
        self.bind_current_stream()
        st = _lib.QbProgramStats()
        _lib.check(self.lib.qb_program_run_permuted(self.handle, prog.handle, src.data_ptr(), dst.data_ptr(), 0, ctypes.byref(st)))
        return st

    def run_program(self, prog: "CompiledProgram", state: DeviceArray, timed: bool = False, alt: Optional[DeviceArray] = None,
                    spans: Optional[list] = None, input_zero: bool = False):
        """Apply a compiled program to ``state`` (same nqubits / dtype / device as it was compiled for).  ``timed`` /
        ``spans`` as in ``apply_program``."""
        self.bind_current_stream()
        if np.dtype(state.dtype) != prog.dtype or state.size != (1 << prog.nqubits):
            raise ValueError(f"program compiled for {prog.nqubits} qubits of {prog.dtype}, got a state of {state.size} x {state.dtype}")
        total = _lib.QbProgramStats()
        total.nops = prog.nops
        total.perm_ms, total.nperm = 0.0, 0
        flags = _lib.QB_PROGRAM_TIME if timed else 0
        total.nperm_fused = 0
        # ``input_zero``: as in ``apply_program`` -- the first segment's first sweep makes |0...0> when it is a sweep program
        zero_flag = _lib.QB_PROGRAM_INPUT_ZERO if input_zero else 0
        if zero_flag and (not prog.segments or prog.segments[0][0] not in ("prog", "progperm")):
            self._write_zero_state(state, prog.nqubits)
            zero_flag = 0
        for index, (kind, payload) in enumerate(prog.segments):
            if kind == "wide":
                self.apply_wide(state, prog.nqubits, payload)
                total.nsweeps += 3
                total.bytes_moved += 6.0 * state.nbytes
                continue
            if kind == "progperm":
                st = _lib.QbProgramStats()
                try:
                    scratch = self._alloc(tuple(state.tensor.shape), state.tensor.dtype) if alt is None else alt.tensor
                except torch.cuda.OutOfMemoryError:  # no second buffer: the gates in place, the permutation as SWAP gates
                    gops, dest = prog.fused_source[index]
                    st = self._apply_sweeps(state, prog.nqubits, list(gops) + swaps_for_permutation(dest), True, timed, zero_flag)
                else:
                    e0 = _record_event() if spans is not None else None
                    _lib.check(self.lib.qb_program_run_permuted(self.handle, payload, state.data_ptr(), scratch.data_ptr(), flags | zero_flag,
                                                                ctypes.byref(st)))
                    self._adopt_result(state, scratch, alt)
                    if e0 is not None:
                        spans.append(("sweep", e0, _record_event(), st.nsweeps))
                    total.nperm_fused += 1
                zero_flag = 0
                total.nsweeps += st.nsweeps
                total.ndense_passes += st.ndense_passes
                total.ndiag_ops += st.ndiag_ops
                total.nstage_sweeps += st.nstage_sweeps
                total.bytes_moved += st.bytes_moved
                total.elapsed_ms += st.elapsed_ms
                continue
            if kind == "perm":
                try:
                    ms = self.permute_qubits(state, prog.nqubits, payload, timed=timed, alt=alt, spans=spans)
                except torch.cuda.OutOfMemoryError:  # no room for the scratch buffer: in place, as SWAP gates
                    st = self._apply_sweeps(state, prog.nqubits, swaps_for_permutation(payload), True, timed)
                    total.nsweeps += st.nsweeps
                    total.bytes_moved += st.bytes_moved
                    total.elapsed_ms += st.elapsed_ms
                    continue
                total.nsweeps += 1
                total.ndense_passes += 1
                total.bytes_moved += 2.0 * state.nbytes
                total.elapsed_ms += ms or 0.0
                total.perm_ms += ms or 0.0
                total.nperm += 1
                continue
            st = _lib.QbProgramStats()
            e0 = _record_event() if spans is not None else None
            _lib.check(self.lib.qb_program_run(self.handle, payload, state.data_ptr(), flags | zero_flag, ctypes.byref(st)))
            zero_flag = 0
            if e0 is not None:
                spans.append(("sweep", e0, _record_event(), st.nsweeps))
            total.nsweeps += st.nsweeps
            total.ndense_passes += st.ndense_passes
            total.ndiag_ops += st.ndiag_ops
            total.nstage_sweeps += st.nstage_sweeps
            total.bytes_moved += st.bytes_moved
            total.elapsed_ms += st.elapsed_ms
        self.last_stats = total
        return total

    # ---- K3: probabilities --------------------------------------------------------------------------
    def probabilities(self, state: DeviceArray, qubits: Sequence[int], nqubits: int) -> DeviceArray:
        self.bind_current_stream()
        rdtype = np.dtype("float64") if state.dtype == np.dtype("complex128") else np.dtype("float32")
        out = self.empty((1 << len(qubits),), rdtype)
        _lib.check(
            self.lib.qb_probabilities(
                self.handle, state.data_ptr(), nqubits, _DT[state.dtype], _int_array(qubits), len(qubits), out.data_ptr()
            )
        )
        return out

    def probabilities_dm(self, rho: DeviceArray, qubits: Sequence[int], nqubits: int) -> DeviceArray:
        """calculate_probabilities(density_matrix=True), abstract.py:2741-2749, on a device-resident rho."""
        self.bind_current_stream()
        rdtype = np.dtype("float64") if rho.dtype == np.dtype("complex128") else np.dtype("float32")
        out = self.empty((1 << len(qubits),), rdtype)
        _lib.check(
            self.lib.qb_probabilities_dm(
                self.handle, rho.data_ptr(), nqubits, _DT[rho.dtype], _int_array(qubits), len(qubits), out.data_ptr()
            )
        )
        return out

    def collapse_dm(self, rho: DeviceArray, nqubits: int, qubits: Sequence[int], outcome: int, normalize: bool = True):
        """_collapse_density_matrix, abstract.py:3249-3277, in place on a device-resident rho."""
        self.bind_current_stream()
        _lib.check(
            self.lib.qb_collapse_dm(
                self.handle, rho.data_ptr(), nqubits, _DT[rho.dtype], _int_array(qubits), len(qubits), int(outcome), 1 if normalize else 0
            )
        )
        return rho

    # ---- K4: sampling ---------------------------------------------------------------------------------
    def sample(self, probs: DeviceArray, uniforms: np.ndarray, mode: Optional[int] = None, return_total: bool = False):
        """Inverse-CDF sampling: ``searchsorted(cumsum(p) / sum, u, side="right")`` as np.random.choice does."""
        self.bind_current_stream()
        nbins = probs.size
        if mode is None:
            mode = _lib.QB_SCAN_EXACT if nbins <= self.exact_scan_max_bins else _lib.QB_SCAN_PARALLEL
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        out = np.empty(uniforms.shape[0], dtype=np.int64)
        total = ctypes.c_double()
        _lib.check(
            self.lib.qb_sample(
                self.handle, probs.data_ptr(), _RT[probs.dtype], nbins, uniforms.ctypes.data, uniforms.shape[0],
                out.ctypes.data, mode, ctypes.byref(total),
            )
        )
        return (out, total.value) if return_total else out

    def cdf(self, probs: DeviceArray, mode: Optional[int] = None) -> DeviceArray:
        self.bind_current_stream()
        if mode is None:
            mode = _lib.QB_SCAN_EXACT if probs.size <= self.exact_scan_max_bins else _lib.QB_SCAN_PARALLEL
        out = self.empty((probs.size,), np.float64)
        _lib.check(self.lib.qb_cdf(self.handle, probs.data_ptr(), _RT[probs.dtype], probs.size, out.data_ptr(), mode))
        return out

    def sample_cdf(self, cdf: DeviceArray, uniforms: DeviceArray) -> DeviceArray:
        out = self.empty((uniforms.size,), np.int64)
        _lib.check(self.lib.qb_sample_cdf(self.handle, cdf.data_ptr(), cdf.size, uniforms.data_ptr(), uniforms.size, out.data_ptr()))
        return out

    # ---- K9: expectation values ------------------------------------------------------------------------------
    def expval_pauli(self, state: DeviceArray, nqubits: int, paulis: str, qubits: Sequence[int]) -> complex:
        """<state| P |state> for the Pauli string ``paulis`` (one of I, X, Y, Z per entry of ``qubits``)."""
        out = (ctypes.c_double * 2)()
        _lib.check(
            self.lib.qb_expval_pauli(
                self.handle, state.data_ptr(), nqubits, _DT[state.dtype], paulis.encode(), _int_array(qubits), len(qubits), out
            )
        )
        return complex(out[0], out[1])

    def vdot(self, a: DeviceArray, b: DeviceArray, nqubits: int) -> complex:
        """<a|b> (conjugate on the first argument, as numpy.vdot)."""
        if a.dtype != b.dtype:
            raise ValueError("vdot needs two states of the same dtype")
        out = (ctypes.c_double * 2)()
        _lib.check(self.lib.qb_state_vdot(self.handle, a.data_ptr(), b.data_ptr(), nqubits, _DT[a.dtype], out))
        return complex(out[0], out[1])

    # ---- K5: collapse -----------------------------------------------------------------------------------
    def collapse(self, state: DeviceArray, nqubits: int, qubits: Sequence[int], outcome: int, normalize: bool = True):
        self.bind_current_stream()
        _lib.check(
            self.lib.qb_collapse(
                self.handle, state.data_ptr(), nqubits, _DT[state.dtype], _int_array(qubits), len(qubits), int(outcome),
                1 if normalize else 0,
            )
        )
        return state

    # ---- K7 plumbing: IPC-exportable buffers and peer mappings -------------------------------------------------
    def malloc_exportable(self, shape, dtype):
        """A buffer from qb_malloc (plain cudaMalloc, exportable through CUDA IPC) viewed as a DeviceArray."""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = count * dtype.itemsize
        ptr = ctypes.c_void_p()
        _lib.check(self.lib.qb_malloc(self.handle, nbytes, ctypes.byref(ptr)))
        owner = _RawCuda(self, ptr.value, nbytes)
        tensor = torch.as_tensor(owner, device=self.device).view(torch_dtype(dtype)).reshape(shape)
        arr = DeviceArray(tensor)
        arr._owner = owner
        return arr

    def ipc_handle(self, array: DeviceArray) -> bytes:
        buf = ctypes.create_string_buffer(64)
        _lib.check(self.lib.qb_ipc_get_handle(self.handle, array.data_ptr(), buf))
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        ptr = ctypes.c_void_p()
        _lib.check(self.lib.qb_ipc_open_handle(self.handle, ctypes.create_string_buffer(handle, 64), ctypes.byref(ptr)))
        return ptr.value

    def swap_half_p2p(self, state: DeviceArray, peer_ptr: int, nqubits: int, local_qubit: int, my_bit: int, part: int, nparts: int):
        _lib.check(
            self.lib.qb_swap_half_p2p(
                self.handle, state.data_ptr(), ctypes.c_void_p(peer_ptr), nqubits, _DT[state.dtype], local_qubit, my_bit, part, nparts
            )
        )

    def alltoall_p2p(self, state: DeviceArray, entries, push: bool = False):
        """K7b: ``entries`` = [(buffer pointer, my_offset, its_offset, begin, end)] in amplitudes
        (distributed.alltoall_entries: in-place chunk swaps; distributed.alltoall_push_entries with ``push``: copies
        into the destination ranks' second buffers)."""
        k = len(entries)
        ptrs = (ctypes.c_void_p * k)(*[e[0] for e in entries])
        cols = [(ctypes.c_uint64 * k)(*[int(e[c]) for e in entries]) for c in (1, 2, 3, 4)]
        fn = self.lib.qb_alltoall_push_p2p if push else self.lib.qb_alltoall_p2p
        _lib.check(fn(self.handle, state.data_ptr(), _DT[state.dtype], k, ptrs, *cols))

    def memcpy_async(self, dst_ptr: int, src_ptr: int, nbytes: int, stream: int = 0):
        """Device-to-device DMA copy (also across peer-mapped buffers) on ``stream`` (a cudaStream_t; 0 = the engine's)."""
        _lib.check(self.lib.qb_memcpy_async(self.handle, ctypes.c_void_p(dst_ptr), ctypes.c_void_p(src_ptr), int(nbytes), 2,
                                            ctypes.c_void_p(stream or None)))

    def mem_info(self):
        free, total = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(self.lib.qb_mem_info(self.handle, ctypes.byref(free), ctypes.byref(total)))
        return free.value, total.value


def _record_event():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def resolve_spans(spans):
    """-> {"sweep": (ms, launches), "perm": (ms, launches), ...} of a span list (waits for the last event of each)."""
    out = {}
    for kind, e0, e1, count in spans:
        e1.synchronize()
        ms, n = out.get(kind, (0.0, 0))
        out[kind] = (ms + e0.elapsed_time(e1), n + count)
    return out


class _RawCuda:
    """Owner of a qb_malloc'ed buffer, exposed to torch through __cuda_array_interface__."""

    def __init__(self, engine, ptr, nbytes):
        self.engine, self.ptr = engine, ptr
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}

    def __del__(self):  # pragma: no cover
        try:
            if self.engine.handle:
                self.engine.lib.qb_free(self.engine.handle, ctypes.c_void_p(self.ptr))
        except Exception:
            pass


def _dropped_program():
    return None


class _CopyingProgram:
    """Handle of an out-of-place sweep program (Engine.compile_copying); freed with the object, pickles to None."""

    def __init__(self, engine, handle, nsweeps):
        self.engine, self.handle, self.nsweeps = engine, handle, nsweeps

    def __reduce__(self):
        return (_dropped_program, ())

    def close(self):
        if self.handle and getattr(self.engine, "handle", None):
            self.engine.lib.qb_program_destroy(self.engine.handle, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CompiledProgram:
    """Segments of a gate queue compiled for one Engine: ("prog", qb_program handle) for runs of gates, ("perm", dest) for
    runs of plain SWAPs (K8).  Frees its device programs with the object.  Pickles to None: the handles are device-resident
    programs of THIS process (a Circuit that carries its compiled program is compiled again where it is unpickled)."""

    def __reduce__(self):
        return (_dropped_program, ())

    def __init__(self, engine: "Engine", nqubits: int, dtype, ops: Sequence[Op], fuse: bool = True):
        self.engine, self.nqubits, self.dtype, self.nops = engine, nqubits, np.dtype(dtype), len(ops)
        self.segments = []
        self.nsweeps = 0
        segments = split_segments(ops, nqubits, fuse and engine.permute_swap_runs, fuse and engine.fuse_permutations)
        flags = 0 if fuse else _lib.QB_PROGRAM_NO_FUSE
        self.fused_source = {}  # segment index of a "progperm" -> (its ops, dest_of_qubit): the no-second-buffer fallback
        # op index in ``ops`` -> (segment index, index inside that segment's program), or (-1, -1) for ops that became a
        # permutation / a wide block (set_params addresses ops by their position in the compiled list)
        self.op_segment = [(-1, -1)] * len(ops)
        position = {id(op): i for i, op in enumerate(ops)}
        try:
            while segments:
                kind, payload = segments.pop(0)
                if kind == "wide":
                    self.segments.append(("wide", payload))
                    self.nsweeps += 3
                    continue
                if kind == "opsperm":
                    gops, dest = payload
                    arr, keep = pack_ops(gops)
                    handle, st = ctypes.c_void_p(), _lib.QbProgramStats()
                    try:
                        _lib.check(engine.lib.qb_program_create_permuted(
                            engine.handle, nqubits, _DT[self.dtype], arr, len(gops), _int_array(dest),
                            flags | _lib.QB_PROGRAM_PERM_FUSED_ONLY, ctypes.byref(handle), ctypes.byref(st)))
                    except NotImplementedError:
                        segments[:0] = ([("ops", gops)] if gops else []) + [("perm", dest)]
                        continue
                    del keep
                    for local, op in enumerate(gops):
                        self.op_segment[position[id(op)]] = (len(self.segments), local)
                    self.fused_source[len(self.segments)] = (list(gops), list(dest))
                    self.segments.append(("progperm", handle))
                    self.nsweeps += st.nsweeps
                    continue
                if kind == "perm":
                    self.segments.append(("perm", list(payload)))
                    self.nsweeps += 1
                    continue
                if not payload:
                    continue
                for local, op in enumerate(payload):
                    self.op_segment[position[id(op)]] = (len(self.segments), local)
                arr, keep = pack_ops(payload)
                handle, st = ctypes.c_void_p(), _lib.QbProgramStats()
                _lib.check(engine.lib.qb_program_create(engine.handle, nqubits, _DT[self.dtype], arr, len(payload), flags,
                                                        ctypes.byref(handle), ctypes.byref(st)))
                del keep
                self.segments.append(("prog", handle))
                self.nsweeps += st.nsweeps
        except Exception:
            self.close()
            raise

    # numpy mirror of qb_param_update (include/qibo_b200.h): records are filled column-wise, no per-gate Python work
    PARAM_DTYPE = np.dtype([("op_index", "<i4"), ("family", "<i4"), ("conjugate", "<i4"), ("reserved", "<i4"),
                            ("theta", "<f8", (3,)), ("matrix", "<u8")])

    def param_records(self, indices, families, conjugates):
        """Records for ``set_param_records``: one per op (index in the list the program was compiled from).  -> (records
        with op_index already translated to the segment-local index, segment of every record)."""
        assert self.PARAM_DTYPE.itemsize == ctypes.sizeof(_lib.QbParamUpdate)
        rec = np.zeros(len(indices), dtype=self.PARAM_DTYPE)
        seg = np.empty(len(indices), dtype=np.int64)
        for k, index in enumerate(indices):
            seg[k], rec["op_index"][k] = self.op_segment[index]
        if (seg < 0).any():
            raise NotImplementedError("an op inside a SWAP run / wide block has no parameter slot: compile again")
        rec["family"], rec["conjugate"] = families, conjugates
        return rec, seg

    def set_param_records(self, rec, seg):
        """qb_program_set_params on prepared records (``param_records``; the caller has filled ``theta`` / ``matrix``)."""
        for s_ in np.unique(seg):
            part = np.ascontiguousarray(rec[seg == s_])
            _lib.check(self.engine.lib.qb_program_set_params(
                self.engine.handle, self.segments[int(s_)][1], part.ctypes.data_as(ctypes.POINTER(_lib.QbParamUpdate)), len(part)))

    def set_params(self, updates):
        """New matrices for some ops of the program: ``updates`` = [(op index in the list the program was compiled from,
        family, thetas, matrix or None, conjugate)], family one of _lib.QB_GATE_*.  The schedule is kept; the sweep
        programs are emitted again and re-uploaded in stream order."""
        if not updates:
            return
        rec, seg = self.param_records([u[0] for u in updates], [u[1] for u in updates], [1 if u[4] else 0 for u in updates])
        keep = []
        for k, (_, _, thetas, matrix, _) in enumerate(updates):
            th = list(thetas)[:3]
            rec["theta"][k, : len(th)] = th
            if matrix is not None:
                m = np.ascontiguousarray(matrix, dtype=np.complex128)
                keep.append(m)
                rec["matrix"][k] = m.ctypes.data
        self.set_param_records(rec, seg)
        del keep

    def close(self):
        for kind, payload in self.segments:
            if kind in ("prog", "progperm") and payload and getattr(self.engine, "handle", None):
                self.engine.lib.qb_program_destroy(self.engine.handle, payload)
        self.segments = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def plan_program(nqubits: int, dtype, ops: Sequence[Op], fuse: bool = True):
    """Host-only planner call (no GPU needed): -> (stats, sweep index of every op)."""
    lib = _lib.load()
    stats = _lib.QbProgramStats()
    arr, keep = pack_ops(ops)
    sweep_of_op = (ctypes.c_int32 * max(len(ops), 1))()
    flags = 0 if fuse else _lib.QB_PROGRAM_NO_FUSE
    _lib.check(lib.qb_plan_program(nqubits, _DT[np.dtype(dtype)], arr, len(ops), flags, ctypes.byref(stats), sweep_of_op))
    del keep
    return stats, list(sweep_of_op)[: len(ops)]


MIN_SWAP_RUN = 3
_SWAP_MATRIX = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def is_plain_swap(op: Op) -> bool:
    if op.is_diagonal or len(op.targets) != 2 or op.controls:
        return False
    d = op.data
    # (two entries decide for every gate but the SWAP-like ones: no 16-element comparison per CU1 of a QFT)
    return d[0, 0] == 1 and d[1, 2] == 1 and np.array_equal(d, _SWAP_MATRIX)


def split_swap_runs(ops: Sequence[Op], nqubits: int):
    """-> [("ops", [...]) | ("perm", dest_of_qubit)]: maximal runs of >= MIN_SWAP_RUN plain SWAPs become permutations."""
    out, cur, run = [], [], []

    def close_run():
        if len(run) >= MIN_SWAP_RUN:
            if cur:
                out.append(("ops", list(cur)))
                cur.clear()
            dest = list(range(nqubits))
            for op in run:
                a, b = op.targets
                dest = [b if d == a else a if d == b else d for d in dest]
            if dest != list(range(nqubits)):
                out.append(("perm", dest))
        else:
            cur.extend(run)
        run.clear()

    for op in ops:
        if is_plain_swap(op):
            run.append(op)
        else:
            close_run()
            cur.append(op)
    close_run()
    if cur:
        out.append(("ops", cur))
    return out


def split_segments(ops: Sequence[Op], nqubits: int, swap_runs: bool = True, fuse_perm: bool = False):
    """-> [("ops", [...]) | ("perm", dest_of_qubit) | ("wide", op)]: the queue cut at blocks too wide for a sweep tile pass
    (Engine.apply_wide) and, between them, at runs of plain SWAPs (K8).  ``fuse_perm``: a permutation, with the run of
    gates in front of it if any, becomes ("opsperm", (ops, dest_of_qubit)) -- one program whose last sweep permutes."""
    out = _split_segments(ops, nqubits, swap_runs)
    if not fuse_perm:
        return out
    merged = []
    for kind, payload in out:
        if kind == "perm":
            if merged and merged[-1][0] == "ops":
                merged[-1] = ("opsperm", (merged[-1][1], payload))
            else:
                merged.append(("opsperm", ([], payload)))
        else:
            merged.append((kind, payload))
    return merged


def _split_segments(ops: Sequence[Op], nqubits: int, swap_runs: bool = True):
    out, cur = [], []

    def flush():
        if cur:
            out.extend(split_swap_runs(cur, nqubits) if swap_runs else [("ops", list(cur))])
            cur.clear()

    for op in ops:
        if is_wide(op):
            flush()
            out.append(("wide", op))
        else:
            cur.append(op)
    flush()
    return out


def swaps_for_permutation(dest_of_qubit: Sequence[int]):
    """SWAP gates realising the permutation (fallback when no scratch buffer fits)."""
    dest = list(dest_of_qubit)
    ops = []
    # data at slot q must go to slot dest[q]: cycle decomposition into transpositions
    cur = list(range(len(dest)))  # cur[s] = original slot of the data now sitting in slot s
    for s in range(len(dest)):
        want = dest.index(s)  # original slot whose data belongs in s
        at = cur.index(want)
        if at != s:
            ops.append(Op(_SWAP_MATRIX, (s, at)))
            cur[s], cur[at] = cur[at], cur[s]
    return ops


def frequencies_from_samples(samples: np.ndarray) -> Counter:
    """calculate_frequencies (abstract.py:2727-2732) on the host-resident sample vector (S3)."""
    res, counts = np.unique(samples, return_counts=True)
    return Counter(dict(zip(res.tolist(), counts.tolist())))
