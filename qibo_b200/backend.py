"""`B200Backend`: the drop-in `qibo.backends.Backend` subclass.

Loaded through Qibo's own plugin convention -- ``qibo.set_backend("qibo_b200")`` ->
``construct_backend`` -> ``qibo_b200.MetaBackend.load`` (backends/__init__.py:325-350).  It derives from
the reference ``NumpyBackend`` (backends/numpy.py:20) so that the ~150-method array API, the gate-matrix
table (npmatrices.py) and ``matrix_fused`` keep working on the host for Qibo's non-hot-path code, and
overrides every method on the state-vector hot path with calls into libqibo_b200.so:

    apply_gate / apply_gate_density_matrix          abstract.py:2322-2361, 3176-3197   -> qb_apply_matrix
    execute_circuit / _execute_circuit (gate loop)   abstract.py:2442-2512, 3306-3344   -> qb_apply_program
    execute_circuit_repeated                         abstract.py:2532-2636
    execute_distributed_circuit                      abstract.py:2638-2647 (NotImplemented in-tree)
    calculate_probabilities                          abstract.py:2734-2758              -> qb_probabilities
    sample_shots / sample_frequencies                abstract.py:2760-2781              -> qb_sample
    collapse_state                                   abstract.py:2424-2440, 3279-3304   -> qb_collapse
    zero_state / plus_state / minus_state            abstract.py:2199-2273              -> qb_state_*

States returned to Qibo are :class:`qibo_b200.array.DeviceArray` objects (GPU resident, torch tensor
inside).  This module imports qibo; everything below it (engine, ops, C ABI) does not.
"""

import os
from collections import Counter

import numpy as np
import torch

from qibo import __version__ as qibo_version
from qibo.backends.numpy import NumpyBackend
from qibo.config import SHOT_BATCH_SIZE, log, raise_error
from qibo.gates.abstract import Gate
from qibo.gates.special import FusedGate
from qibo.result import CircuitResult, MeasurementOutcomes, QuantumState

from qibo_b200 import _lib
from qibo_b200.array import DeviceArray
from qibo_b200.array import torch_dtype as torch_dtype_of
from qibo_b200.engine import Engine, frequencies_from_samples

_PLAIN_NUMBERS = (float, int, np.float64, np.float32, np.int64)
# fixed-arity gates whose matrix is a function of their numeric parameters alone (backends/npmatrices.py)
_CACHED_MATRIX_GATES = frozenset(
    "H X Y Z S SDG T TDG SX SXDG RX RY RZ U1 U2 U3 GPI GPI2 CNOT CY CZ CSX CSXDG CRX CRY CRZ CU1 CU2 CU3 SWAP iSWAP SiSWAP SiSWAPDG "
    "FSWAP fSim SYC RXX RYY RZZ RZX RXXYY MS GIVENS RBS ECR TOFFOLI CCZ DEUTSCH".split())
from qibo_b200.ops import Op

# gates whose matrix the library evaluates from the angle (qb_program_set_params): class name -> QB_GATE_*
_FAMILIES = {"RX": _lib.QB_GATE_RX, "RY": _lib.QB_GATE_RY, "RZ": _lib.QB_GATE_RZ, "U1": _lib.QB_GATE_U1,
             "CRX": _lib.QB_GATE_CRX, "CRY": _lib.QB_GATE_CRY, "CRZ": _lib.QB_GATE_CRZ, "CU1": _lib.QB_GATE_CU1}

_COMPLEX = {"complex128": np.dtype("complex128"), "complex64": np.dtype("complex64"),
            "float64": np.dtype("complex128"), "float32": np.dtype("complex64")}


class B200Backend(NumpyBackend):
    def __init__(self, device=None, dtype="complex128"):
        super().__init__()
        self.name = "qibo_b200"
        self.platform = "cuda-sm100a"
        self.supports_multigpu = True
        self.oom_error = (torch.cuda.OutOfMemoryError, _lib.QiboB200OutOfMemory, MemoryError)
        # (DeviceArray is what this backend returns as a state: models/evolution.py:105, hamiltonians and parallel.py test
        # `isinstance(x, backend.tensor_types)` to tell a tensor from a result object)
        self.tensor_types = (np.ndarray, DeviceArray)
        self.versions = {"qibo": qibo_version, "numpy": np.__version__, "torch": torch.__version__,
                         "qibo_b200": _lib.load().qb_version()}
        self._engines = {}
        # Circuit objects keep their compiled program between executions (parameter slots, _compiled_circuit)
        self.compile_circuits = True
        # execute_circuit without an initial state: no 2^n-amplitude fill, the first sweep of the compiled program makes the
        # |0...0> tiles (QB_PROGRAM_INPUT_ZERO)
        self.lazy_zero_state = os.environ.get("QB_NO_LAZY_ZERO", "0") in ("", "0")
        self._matrix_cache = {}
        Engine._freeze_imports(again=True)  # qibo and its dependencies are imported by now
        index = 0
        if device is not None:
            index = self._parse_device(device)
        elif torch.distributed.is_available() and torch.distributed.is_initialized():

            index = int(os.environ.get("LOCAL_RANK", 0))
        self.device = f"/GPU:{index}"
        self.engine_gpu = self._engine(index)  # raises without a CUDA device: there is no CPU path
        if dtype != self.dtype:
            self.set_dtype(dtype)

    @property
    def exact_sampling(self):
        """True: shot sampling keeps NumPy's sequential-cumsum CDF (bit-identical to np.random.choice) for ANY number of
        bins; default: up to 2^22 bins, the parallel scan above (SURVEY 8a hazard 1: about 1e2 of 1e6 samples may then
        land in a neighbouring bin at 2^30 bins; the sequential scan costs seconds there)."""
        return self.engine_gpu.exact_scan_max_bins >= (1 << 62)

    @exact_sampling.setter
    def exact_sampling(self, on):
        from qibo_b200.engine import EXACT_SCAN_MAX_BINS

        for eng in self._engines.values():
            eng.exact_scan_max_bins = (1 << 62) if on else EXACT_SCAN_MAX_BINS

    # ------------------------------------------------------------------ configuration -------------
    @staticmethod
    def _parse_device(device):
        if isinstance(device, int):
            return device
        name = str(device)
        if not name.upper().startswith("/GPU:"):
            raise_error(ValueError, f"Device {device} is not available for qibo_b200 backend (GPU only).")
        return int(name.split(":")[1])

    def _engine(self, index):
        if index not in self._engines:
            if index >= torch.cuda.device_count():
                raise_error(ValueError, f"Device /GPU:{index} is not available ({torch.cuda.device_count()} GPUs visible).")
            self._engines[index] = Engine(index)
        return self._engines[index]

    def set_device(self, device):
        index = self._parse_device(device)
        self.engine_gpu = self._engine(index)
        self.device = f"/GPU:{index}"

    def set_threads(self, nthreads):
        """A CPU notion (backends/numpy.py:89-96): accepted and ignored, the work runs on the GPU."""
        if not isinstance(nthreads, int) or nthreads < 1:
            raise_error(TypeError if not isinstance(nthreads, int) else ValueError, "nthreads must be a positive integer.")
        self.nthreads = 1

    @property
    def _cdtype(self):
        """complex dtype the state is held in (float32/float64 backends hold a complex state)."""
        return _COMPLEX[str(np.dtype(self.dtype))]

    @property
    def _real_dtype(self):
        d = np.dtype(self.dtype)
        return d if d.kind == "f" else None

    # ------------------------------------------------------------------ casting / interop -----------
    def is_device(self, x):
        return isinstance(x, DeviceArray)

    def cast(self, array, dtype=None, copy=False):
        """DeviceArray -> host ndarray for Qibo's inherited NumPy code paths (interop).
        Hot-path methods never route a device state through here."""
        if isinstance(array, DeviceArray):
            host = array.numpy()
            return host if dtype is None else host.astype(dtype, copy=False)
        if isinstance(array, (list, tuple)) and any(isinstance(x, DeviceArray) for x in array):
            array = [np.asarray(x) for x in array]
        return super().cast(array, dtype=dtype, copy=copy)

    def to_numpy(self, array):
        if isinstance(array, DeviceArray):
            return array.numpy()
        return super().to_numpy(array)

    def _to_device(self, state, complex_dtype=None):
        """Accept a DeviceArray / ndarray / list state -> complex DeviceArray on this backend's GPU."""
        eng = self.engine_gpu
        if isinstance(state, DeviceArray):
            if state.tensor.device != eng.device:
                # a state made on another GPU (set_device since, or another rank's): move it rather than hand a foreign
                # pointer to this device's kernels
                state = DeviceArray(state.tensor.to(eng.device))
            if state.dtype.kind == "c":
                return state
            return DeviceArray(state.tensor.to(torch.complex128 if state.dtype == np.float64 else torch.complex64))
        host = np.asarray(state)
        if host.dtype.kind != "c":
            host = host.astype(np.complex64 if host.dtype == np.float32 else np.complex128)
        elif host.dtype not in (np.complex64, np.complex128):
            host = host.astype(np.complex128)
        if complex_dtype is not None:
            host = host.astype(complex_dtype, copy=False)
        return eng.upload(host)

    # ------------------------------------------------------------------ state constructors ----------
    def _state_dtype(self, dtype):
        if dtype is None:
            return self._cdtype
        d = np.dtype(dtype)
        return d if d.kind == "c" else (np.dtype("complex64") if d == np.float32 else np.dtype("complex128"))

    def zero_state(self, nqubits, density_matrix=False, dtype=None):
        self._validate_nqubits(nqubits, density_matrix=density_matrix)
        n = 2 * nqubits if density_matrix else nqubits
        state = self.engine_gpu.basis_state(n, self._state_dtype(dtype), 0)
        return state.reshape(2**nqubits, 2**nqubits) if density_matrix else state

    def plus_state(self, nqubits, density_matrix=False, dtype=None):
        self._validate_nqubits(nqubits, density_matrix=density_matrix)
        n = 2 * nqubits if density_matrix else nqubits
        value = 1.0 / 2**nqubits if density_matrix else 1.0 / np.sqrt(2**nqubits)
        state = self.engine_gpu.filled_state(n, value, self._state_dtype(dtype))
        return state.reshape(2**nqubits, 2**nqubits) if density_matrix else state

    def minus_state(self, nqubits, density_matrix=False, dtype=None):
        self._validate_nqubits(nqubits, density_matrix=density_matrix)
        # |-> on every qubit = Z on every qubit of |+...+>
        # (density matrix: |-><-| on every qubit = Z on every row AND column qubit of the all-equal matrix)
        n = 2 * nqubits if density_matrix else nqubits
        state = self.plus_state(nqubits, density_matrix=density_matrix, dtype=dtype)
        flat = state.reshape(-1) if density_matrix else state
        z = np.array([1, -1], dtype=np.complex128)
        ops = [Op(z, (q,), is_diagonal=True) for q in range(n)]
        self.engine_gpu.apply_program(flat, n, ops)
        return state

    # ------------------------------------------------------------------ gate -> Op ---------------------
    def _gate_ops(self, gate, nqubits, density_matrix=False):
        """Gate object -> list of Ops.  Named controlled gates arrive with their full matrix over
        gate.qubits (the library recovers the control structure exactly); `controlled_by` gates arrive as
        target matrix + controls (abstract.py:3176-3197).  Density matrices are 2n-qubit vectors: U on the
        row qubits, conj(U) on the column qubits (abstract.py:2341-2348)."""
        if isinstance(gate, FusedGate):
            # circuit.fuse() blocks (gates/special.py:28-157): the member gates go to the sweep planner one by one -- it
            # packs them into the same HBM sweep anyway, and neither the host-side scipy product of matrix_fused
            # (abstract.py:2680-2717, ~0.5 ms per block) nor a dense complex 2^r x 2^r multiply per amplitude is paid
            return [op for member in gate.gates for op in self._gate_ops(member, nqubits, density_matrix)]
        if gate.is_controlled_by:
            targets, controls = tuple(gate.target_qubits), tuple(gate.control_qubits)
        else:
            targets, controls = tuple(gate.qubits), ()
        # gate.matrix(backend) builds a fresh array per call (npmatrices.py); a queue repeats few distinct matrices (a QFT:
        # H, SWAP and one CU1 per distance), so they are kept by (gate class, shape, parameters).  Only plain numeric
        # parameters make a key: Unitary-like gates (array parameters) and symbolic ones are evaluated every time.
        key = None
        params = gate.parameters
        if gate.__class__.__name__ in _CACHED_MATRIX_GATES and not gate.symbolic_parameters and all(type(p) in _PLAIN_NUMBERS for p in params):
            key = (gate.__class__, len(targets), len(controls), tuple(params), self.dtype)
        matrix = self._matrix_cache.get(key) if key is not None else None
        if matrix is None:
            matrix = np.ascontiguousarray(np.asarray(gate.matrix(self)).astype(np.complex128, copy=False))
            if key is not None:
                if len(self._matrix_cache) >= 4096:
                    self._matrix_cache.clear()
                matrix.setflags(write=False)
                self._matrix_cache[key] = matrix
        if not density_matrix:
            return [Op(matrix, targets, controls, name=gate.__class__.__name__)]
        return [
            Op(np.conj(matrix), tuple(q + nqubits for q in targets), tuple(q + nqubits for q in controls)),
            Op(matrix, targets, controls),
        ]

    # ------------------------------------------------------------------ G1/G2: apply_gate ---------------
    def apply_gate(self, gate, state, nqubits):
        """In place on DeviceArray states (returns the same object, as qibojit documents,
        doc/source/getting-started/backends.rst:57-61); host arrays are uploaded first and left untouched."""
        density_matrix = len(state.shape) == 2
        dev = self._to_device(state)
        flat = dev.reshape(-1) if density_matrix else dev
        n = 2 * nqubits if density_matrix else nqubits
        for op in self._gate_ops(gate, nqubits, density_matrix):
            self.engine_gpu.apply_op(flat, n, op)
        return dev

    def apply_gate_density_matrix(self, gate, state, nqubits):
        """Older-API alias named by the north star; the reference dispatches on ndim inside apply_gate."""
        return self.apply_gate(gate, state, nqubits)

    def apply_gate_half_density_matrix(self, gate, state, nqubits):
        """abstract.py:2363-2383: only the left multiplication U rho."""
        if gate.is_controlled_by:  # pragma: no cover
            raise_error(NotImplementedError, "Gate density matrix half call is not implemented for ``controlled_by`` gates.")
        dev = self._to_device(state)
        flat = dev.reshape(-1)
        op = self._gate_ops(gate, nqubits, density_matrix=True)[1]
        self.engine_gpu.apply_op(flat, 2 * nqubits, op)
        return dev

    # ------------------------------------------------------------------ E1: circuits ----------------------
    @staticmethod
    def _is_plain(gate):
        """True for gates whose apply() is Gate.apply -> backend.apply_gate (everything but M, callbacks, channels)."""
        return type(gate).apply is Gate.apply

    def _run_queue(self, queue, state, nqubits, density_matrix, substitute_symbols=False, program_cache=None):
        """The gate loop of _execute_circuit (abstract.py:3321-3322), with maximal runs of plain gates handed
        to the sweep planner in one C call.  ``program_cache`` (a dict owned by the caller, one per circuit and
        nqubits/dtype): runs of plain gates without symbolic parameters are compiled on first use
        (qb_program_create) and re-launched afterwards -- execute_circuit_repeated simulates the same queue once per
        shot (abstract.py:2532-2636)."""
        flat_n = 2 * nqubits if density_matrix else nqubits
        pending, gates_of_run = [], []
        run_start = [0, True]  # index of the run's first gate in the queue, cacheable

        def flush(st):
            if gates_of_run:
                flat = st.reshape(-1) if density_matrix else st
                key = (run_start[0], len(gates_of_run), flat_n, str(flat.dtype))
                cacheable = program_cache is not None and run_start[1] and str(flat.dtype) in ("complex64", "complex128")
                prog = program_cache.get(key) if cacheable else None
                if prog is None:
                    for g in gates_of_run:
                        pending.extend(self._gate_ops(g, nqubits, density_matrix))
                    if cacheable:
                        prog = program_cache[key] = self.engine_gpu.compile(flat_n, flat.dtype, pending)
                if prog is not None:
                    self.engine_gpu.run_program(prog, flat)
                else:
                    self.engine_gpu.apply_program(flat, flat_n, pending)
                if flat is not st and flat.tensor.data_ptr() != st.tensor.data_ptr():
                    st.tensor = flat.tensor.reshape(st.shape)  # a permutation sweep re-pointed the flat view
                pending.clear()
                gates_of_run.clear()
            return st

        for index, gate in enumerate(queue):
            if substitute_symbols and gate.symbolic_parameters:
                # abstract.py:2590-2592: evaluated in queue order, i.e. after the collapses they depend on
                gate.substitute_symbols()
            if self._is_plain(gate):
                if not gates_of_run:
                    run_start[0], run_start[1] = index, True
                if gate.symbolic_parameters:
                    run_start[1] = False  # its matrix depends on earlier measurement outcomes: plan afresh
                gates_of_run.append(gate)
            else:
                state = flush(state)
                state = gate.apply(self, state, nqubits)
                if not isinstance(state, DeviceArray):
                    state = self._to_device(state)
        return flush(state)

    # ------------------------------------------------------------------ f2: compiled circuits with parameter slots
    def _param_slots(self, gate, first_op, nops, density_matrix):
        """-> [(op index, family, conjugate)] for a parametrised gate whose ops start at ``first_op``."""
        if not gate.parameters or isinstance(gate, FusedGate):
            return []
        family = _FAMILIES.get(gate.__class__.__name__, _lib.QB_GATE_MATRIX)
        if gate.is_controlled_by or len(gate.parameters) != 1:
            family = _lib.QB_GATE_MATRIX  # target matrix + controls, U2 / U3 / fSim ...: the host builds the matrix
        if density_matrix:  # (conj(U) on the column qubits first, then U on the row qubits: _gate_ops)
            return [(first_op, family, True), (first_op + 1, family, False)]
        return [(first_op, family, False)]

    def _compiled_circuit(self, circuit, nqubits, density_matrix, dtype):
        """The whole queue as ONE compiled program, cached on the circuit object and re-parametrised in place when
        ``Circuit.set_parameters`` (models/circuit.py:788-857) changed angles since the last execution: a variational loop
        pays kernel launches plus one small C call per step instead of 2 Python calls per gate and a fresh plan.  None for
        queues with special gates in the middle (callbacks, collapse) or symbolic parameters -- they take _run_queue."""
        queue = circuit.queue
        plain = [g for g in queue if self._is_plain(g)]
        if len(plain) != len(queue):
            # measurement gates without collapse at the END of the queue leave the state alone; anything else: general path
            tail = queue[len(plain):] if all(self._is_plain(g) for g in queue[: len(plain)]) else None
            if tail is None or any(g.__class__.__name__ != "M" or g.collapse for g in tail):
                return None
        if any(g.symbolic_parameters for g in plain):
            return None
        key = (tuple(map(id, plain)), nqubits, density_matrix, str(dtype), id(self.engine_gpu))
        cached = getattr(circuit, "_qb200_program", None)
        if cached is not None and cached.get("program") is not None and cached["key"] == key:
            self._update_parameters(cached)
            return cached["program"]
        ops, slots = [], []
        for gate in plain:
            members = gate.gates if isinstance(gate, FusedGate) else [gate]  # fused blocks are parametrised member by member
            for member in members:
                mops = self._gate_ops(member, nqubits, density_matrix)
                sl = self._param_slots(member, len(ops), len(mops), density_matrix)
                if sl:
                    slots.append((member, sl))
                ops.extend(mops)
        flat_n = 2 * nqubits if density_matrix else nqubits
        program = self.engine_gpu.compile(flat_n, dtype, ops)
        entry = None
        # (ops inside SWAP runs / wide blocks have no slot in a sweep program: such circuits are compiled on every call)
        if all(program.op_segment[i][0] >= 0 for _, sl in slots for i, _, _ in sl):
            fam = [(g, sl) for g, sl in slots if all(f != _lib.QB_GATE_MATRIX for _, f, _ in sl)]
            other = [(g, sl) for g, sl in slots if any(f == _lib.QB_GATE_MATRIX for _, f, _ in sl)]
            rec, seg = program.param_records([i for _, sl in fam for i, _, _ in sl], [f for _, sl in fam for _, f, _ in sl],
                                             [1 if c else 0 for _, sl in fam for _, _, c in sl])
            entry = {"key": key, "program": program, "fam_gates": [g for g, _ in fam], "fam_rec": rec, "fam_seg": seg,
                     "fam_row_gate": np.array([k for k, (_, sl) in enumerate(fam) for _ in sl], dtype=np.int64),
                     "fam_thetas": np.array([float(np.real(g.parameters[0])) for g, _ in fam], dtype=np.float64),
                     "other": other, "other_thetas": [tuple(g.parameters) for g, _ in other]}
        try:
            circuit._qb200_program = entry
        except Exception:  # (a circuit object that does not take attributes)
            pass
        return program

    def _update_parameters(self, cached):
        """Angles that changed since the cached program last ran -> qb_program_set_params.  Gates of the families the
        library evaluates itself (RX, RY, RZ, U1, CU1, CRX, CRY, CRZ) cost one float each here; any other parametrised gate
        sends its new matrix."""
        program = cached["program"]
        gates = cached["fam_gates"]
        if gates:
            thetas = np.fromiter((np.real(g.parameters[0]) for g in gates), dtype=np.float64, count=len(gates))
            changed = thetas != cached["fam_thetas"]
            if changed.any():
                rows = changed[cached["fam_row_gate"]]
                rec = cached["fam_rec"][rows]
                rec["theta"][:, 0] = thetas[cached["fam_row_gate"][rows]]
                program.set_param_records(rec, cached["fam_seg"][rows])
                cached["fam_thetas"] = thetas
        if cached["other"]:
            now = [tuple(g.parameters) for g, _ in cached["other"]]
            updates = []
            for (gate, sl), new, old in zip(cached["other"], now, cached["other_thetas"]):
                if len(new) == len(old) and all(np.array_equal(a, b) for a, b in zip(new, old)):
                    continue
                matrix = np.asarray(gate.matrix(self)).astype(np.complex128, copy=False)
                for index, family, conj in sl:
                    updates.append((index, _lib.QB_GATE_MATRIX, [], matrix, conj))
            program.set_params(updates)
            cached["other_thetas"] = now

    def _execute_circuit(self, circuit, initial_state=None, nshots=1000):
        nqubits = circuit.nqubits
        density_matrix = circuit.density_matrix
        self.engine_gpu.reclaim_before((16 if self.dtype != "complex64" else 8) << (2 * nqubits if density_matrix else nqubits))
        lazy_zero = False
        if initial_state is None:
            if self.compile_circuits and self.lazy_zero_state:
                # zero_state (abstract.py:2243-2273) without the fill: the compiled program's first sweep makes the |0...0>
                # tiles itself (QB_PROGRAM_INPUT_ZERO); written out below if the queue takes the general path after all
                self._validate_nqubits(nqubits, density_matrix=density_matrix)
                flat_n = 2 * nqubits if density_matrix else nqubits
                state = self.engine_gpu.uninitialised_state(flat_n, self._state_dtype(None))
                if density_matrix:
                    state = state.reshape(2**nqubits, 2**nqubits)
                lazy_zero = True
            else:
                state = self.zero_state(nqubits, density_matrix=density_matrix)
        else:
            # the caller's array is never clobbered (SURVEY 8b ownership): device inputs are cloned
            state = self._to_device(initial_state)
            if state is initial_state:
                state = state.copy()
        program = None
        if self.compile_circuits and str(state.dtype) in ("complex64", "complex128"):
            program = self._compiled_circuit(circuit, nqubits, density_matrix, state.dtype)
        if lazy_zero and program is None:
            self.engine_gpu._write_zero_state(state.reshape(-1) if density_matrix else state, 2 * nqubits if density_matrix else nqubits)
            lazy_zero = False
        if program is not None:
            flat = state.reshape(-1) if density_matrix else state
            self.engine_gpu.run_program(program, flat, input_zero=lazy_zero)
            if flat is not state and flat.tensor.data_ptr() != state.tensor.data_ptr():
                state.tensor = flat.tensor.reshape(state.shape)
            for gate in circuit.queue:
                if not self._is_plain(gate):
                    gate.result.backend = self  # what M.apply does for a non-collapsing measurement
        else:
            state = self._run_queue(circuit.queue, state, nqubits, density_matrix)
        init_dtype = getattr(initial_state, "dtype", None) if initial_state is not None else None
        if self._real_dtype is not None and (init_dtype is None or np.dtype(init_dtype).kind != "c"):
            # float32/float64 backends (real-matrix circuits, abstract.py:133-179): the kernels hold a complex state of
            # the same precision; hand back its real part in the backend's dtype (tests/test_backends_global.py:42-77)
            state = DeviceArray(state.tensor.real.contiguous().to(torch_dtype_of(self._real_dtype)))

        if circuit.measurements:
            circuit._final_state = CircuitResult(state, circuit.measurements, backend=self, nshots=nshots)
        else:
            circuit._final_state = QuantumState(state, backend=self)
        return circuit._final_state

    def execute_circuit(self, circuit, initial_state=None, nshots=1000):
        """abstract.py:2442-2512 with device-aware initial-state handling."""
        nqubits = circuit.nqubits
        density_matrix = circuit.density_matrix
        self._validate_nqubits(nqubits, density_matrix=density_matrix)

        if isinstance(initial_state, type(circuit)):
            if not bool(initial_state.density_matrix == density_matrix):
                raise_error(ValueError, f"Cannot set circuit with density_matrix {initial_state.density_matrix} as"
                            + f"initial state for circuit with density_matrix {density_matrix}.")
            if not bool(initial_state.accelerators == circuit.accelerators):  # pragma: no cover
                raise_error(ValueError, "Cannot set circuit with different accelerators as initial state.")
            return self.execute_circuit(initial_state + circuit, None, nshots)

        if initial_state is not None:
            valid_shape = 2 * (2**nqubits,) if density_matrix else (2**nqubits,)
            shape = tuple(initial_state.shape) if hasattr(initial_state, "shape") else tuple(np.shape(initial_state))
            if shape != valid_shape:
                raise_error(ValueError, f"Given initial state has shape {shape}" + f"instead of the expected {valid_shape}.")

        if circuit.repeated_execution:
            if not circuit.measurements and not circuit.has_collapse:
                raise_error(
                    RuntimeError,
                    "Attempting to perform noisy simulation with `density_matrix=False` "
                    + "and no Measurement gate in the Circuit. If you wish to retrieve the "
                    + "statistics of the outcomes please include measurements in the circuit, "
                    + "otherwise set `density_matrix=True` to recover the final state.",
                )
            return self.execute_circuit_repeated(circuit, nshots, initial_state)

        if circuit.accelerators:
            return self.execute_distributed_circuit(circuit, initial_state, nshots)

        try:
            return self._execute_circuit(circuit, initial_state=initial_state, nshots=nshots)
        except self.oom_error:
            raise_error(
                RuntimeError,
                f"State does not fit in {self.device} memory."
                "Please switch the execution device to a "
                "different one using ``qibo.set_device``.",
            )

    def execute_circuit_repeated(self, circuit, nshots, initial_state=None):
        """abstract.py:2532-2636: one full simulation per shot, state kept on the device throughout."""
        density_matrix = circuit.density_matrix
        if circuit.has_collapse and not circuit.measurements and not density_matrix:
            raise_error(
                RuntimeError,
                "The circuit contains only collapsing measurements (`collapse=True`) but "
                + "`density_matrix=False`. Please set `density_matrix=True` to retrieve "
                + "the final state after execution.",
            )
        results, final_states, samples = [], [], []
        nqubits = circuit.nqubits
        if initial_state is None:
            state_copy = self.zero_state(nqubits, density_matrix=density_matrix)
        else:
            state_copy = self._to_device(initial_state)

        program_cache = {}  # the plain-gate runs of the queue, compiled once for all shots
        sharded = False
        if not density_matrix and circuit.accelerators:
            from qibo_b200 import distributed

            sharded = distributed.world_size() > 1
        for _ in range(nshots):
            if sharded:
                # abstract.py:2579-2582: every shot is one distributed execution (collapses act on the sharded state)
                for gate in circuit.queue:
                    if gate.symbolic_parameters:
                        gate.substitute_symbols()
                state = distributed.execute_circuit(self, circuit, state_copy, return_state=True)
            else:
                state = state_copy.copy()
                state = self._run_queue(circuit.queue, state, nqubits, density_matrix, substitute_symbols=True, program_cache=program_cache)
            if density_matrix:
                final_states.append(state)
            if circuit.measurements:
                result = CircuitResult(state, circuit.measurements, backend=self, nshots=1)
                sample = result.samples()[0]
                results.append(sample)
                if not density_matrix:
                    samples.append("".join([str(int(s)) for s in sample]))
                for gate in circuit.measurements:
                    gate.result.reset()

        if density_matrix:  # this implies also it has_collapse
            acc = final_states[0].tensor.clone()
            for st in final_states[1:]:
                acc += st.tensor
            final_state = DeviceArray(acc / len(final_states))
            if circuit.measurements:
                final_result = CircuitResult(final_state, circuit.measurements, backend=self,
                                             samples=self.aggregate_shots(results), nshots=nshots)
            else:
                final_result = QuantumState(final_state, backend=self)
            circuit._final_state = final_result
            return final_result

        final_result = MeasurementOutcomes(circuit.measurements, backend=self, samples=self.aggregate_shots(results), nshots=nshots)
        final_result._repeated_execution_frequencies = self.calculate_frequencies(samples)
        circuit._final_state = final_result
        return final_result

    def execute_distributed_circuit(self, circuit, initial_state=None, nshots=None):
        """D2 (NotImplemented in the reference, abstract.py:2638-2647).  One process per GPU: when
        torch.distributed is initialised with W ranks the state is sharded over log2(W) global qubits
        (qibo_b200.distributed); in a single process the logical devices of ``accelerators`` collapse onto
        this GPU and the circuit runs as one shard -- same results, same return types."""
        from qibo_b200 import distributed

        # The reference marks a distributed circuit as "planned/executed" through its queues object
        # (models/circuit.py:357-361 refuses to reuse it as a subroutine afterwards).  Our planner never touches the
        # caller's gate objects, so only the marker is reproduced.
        queues = getattr(circuit, "queues", None)
        if queues is not None and not queues.queues:
            queues.queues = [[]]
        if distributed.world_size() > 1:
            return distributed.execute_circuit(self, circuit, initial_state, nshots)
        try:
            return self._execute_circuit(circuit, initial_state=initial_state, nshots=1000 if nshots is None else nshots)
        except self.oom_error:
            raise_error(RuntimeError, f"State does not fit in {self.device} memory.")

    # ------------------------------------------------------------------ f1: expectation values ----------------
    def exp_value_observable_symbolic(self, circuit, terms, term_qubits, term_coefficients, nqubits):
        """Sum of Pauli-string terms on the final state (abstract.py:2946-3054).  State vectors never leave the GPU:
        every term is one read pass of K9 (``qb_expval_pauli``) instead of an einsum over a host copy.  Density
        matrices and non-Pauli factors keep the reference's host contraction."""
        if circuit.density_matrix or any(f not in "IXYZ" for term in terms for f in term):
            return super().exp_value_observable_symbolic(circuit, terms, term_qubits, term_coefficients, nqubits)
        result = circuit._final_state if circuit._final_state is not None else self.execute_circuit(circuit)
        state = self._to_device(result.state())
        eng = self.engine_gpu
        expval = 0.0
        for term, qubits, coefficient in zip(terms, term_qubits, term_coefficients):
            kept = [(f, int(q)) for f, q in zip(term, qubits) if f != "I"]
            value = eng.expval_pauli(state, nqubits, "".join(f for f, _ in kept), [q for _, q in kept])
            expval += float(np.real(coefficient * value))
        return expval

    def expectation_value(self, hamiltonian, state, normalize):
        """<state| H |state> for a dense Hamiltonian matrix (abstract.py:2807-2825; Hamiltonian.expectation_from_state)
        with the state left on the device: H is uploaded, H|psi> is one library GEMV (cuBLAS through torch.matmul -- a
        plain dense product, the one place a library kernel is the right tool), the inner product is K9.  Density
        matrices: Re tr(H rho).  Sparse Hamiltonians and host states keep the reference's path."""
        if not isinstance(state, DeviceArray) or state.dtype.kind != "c" or self.is_sparse(hamiltonian):
            return super().expectation_value(hamiltonian, state, normalize)
        h = torch.as_tensor(np.asarray(hamiltonian)).to(state.tensor.device).to(state.tensor.dtype)
        if state.ndim == 2:
            ev = float(torch.einsum("ij,ji->", h, state.tensor).real.item())
            if normalize:
                ev /= float(torch.diagonal(state.tensor).sum().real.item())
            return ev
        n = int(np.log2(state.shape[0]))
        hpsi = DeviceArray(torch.matmul(h, state.tensor))
        ev = float(np.real(self.engine_gpu.vdot(state, hpsi, n)))
        if normalize:
            ev /= self.engine_gpu.norm2(state)
        return ev

    def overlap_statevector(self, state_1, state_2, dtype=None):
        """<state_1|state_2> (abstract.py:2180-2190) with K9 when both states already live on the device."""
        if (isinstance(state_1, DeviceArray) and isinstance(state_2, DeviceArray) and state_1.dtype == state_2.dtype
                and state_1.dtype.kind == "c" and state_1.ndim == 1 and state_1.shape == state_2.shape):
            n = int(np.log2(state_1.shape[0]))
            if 1 << n == state_1.shape[0]:
                return self.engine_gpu.vdot(state_1, state_2, n)
        return super().overlap_statevector(state_1, state_2, dtype=dtype)

    # ------------------------------------------------------------------ P1: probabilities ------------------
    def calculate_probabilities(self, state, qubits, nqubits, density_matrix=False):
        qubits = [int(q) for q in qubits]
        dev = self._to_device(state)
        if density_matrix:
            return self.engine_gpu.probabilities_dm(dev.reshape(-1), qubits, nqubits)
        return self.engine_gpu.probabilities(dev, qubits, nqubits)

    # ------------------------------------------------------------------ S1/S2: sampling ----------------------
    def _probs_to_device(self, probabilities):
        if isinstance(probabilities, DeviceArray):
            if probabilities.dtype.kind == "f":
                return probabilities
            probabilities = probabilities.numpy()
        host = np.asarray(probabilities)
        if host.dtype not in (np.float32, np.float64):
            host = host.astype(np.float64)
        return self.engine_gpu.upload(np.ascontiguousarray(host))

    def sample_shots(self, probabilities, nshots):
        """np.random.choice semantics (abstract.py:2774-2781): uniforms from the global legacy RNG
        (Backend.set_seed = np.random.seed), device scan + inverse-CDF search; int64 host array out."""
        probs = self._probs_to_device(probabilities)
        nshots = int(nshots)
        uniforms = np.random.random_sample(nshots)
        samples, total = self.engine_gpu.sample(probs, uniforms, return_total=True)
        atol = max(np.sqrt(np.finfo(np.float64).eps), np.sqrt(np.finfo(probs.dtype).eps))
        if abs(total - 1.0) > atol:  # numpy/random/mtrand.pyx: "probabilities do not sum to 1"
            raise_error(ValueError, "probabilities do not sum to 1")
        return samples

    def sample_frequencies(self, probabilities, nshots):
        """abstract.py:2760-2772: draws in 2^18 batches from one RNG stream (consecutive batches consume consecutive
        uniforms, so this equals one draw of nshots), histogram.  The CDF is built ONCE on the device and every batch is a
        search in it; it is normalised by its last entry inside the scan (the reference divides by the pairwise sum first
        and by cdf[-1] again inside np.random.choice -- the same bins except for uniforms within an ulp of an edge).  The
        histogram is built from the (host) sample vector -- no 2^m-entry Python loop."""
        probs = self._probs_to_device(probabilities)
        nshots = int(nshots)
        eng = self.engine_gpu
        cdf = eng.cdf(probs)
        freqs = Counter()
        batches = (nshots // SHOT_BATCH_SIZE) * [SHOT_BATCH_SIZE] + [nshots % SHOT_BATCH_SIZE]
        for b in batches:
            if b == 0:
                continue
            uniforms = eng.upload(np.random.random_sample(b))
            freqs.update(frequencies_from_samples(eng.sample_cdf(cdf, uniforms).numpy()))
        return Counter({int(k): int(v) for k, v in sorted(freqs.items())})

    def calculate_frequencies(self, samples):
        if isinstance(samples, DeviceArray):
            samples = samples.numpy()
        return super().calculate_frequencies(samples)

    # ------------------------------------------------------------------ C1: collapse ----------------------------
    def collapse_state(self, state, qubits, shot, nqubits, normalize=True, density_matrix=False):
        dev = self._to_device(state)
        outcome = int(np.asarray(shot).ravel()[0])
        if density_matrix:
            self.engine_gpu.collapse_dm(dev.reshape(-1), nqubits, [int(q) for q in qubits], outcome, normalize)
            return dev
        self.engine_gpu.collapse(dev, nqubits, [int(q) for q in qubits], outcome, normalize)
        return dev

    # ------------------------------------------------------------------ testing helpers ----------------------------
    def assert_allclose(self, value, target, rtol=1e-7, atol=0.0):
        if isinstance(value, (CircuitResult, QuantumState)):
            value = value.state()
        if isinstance(target, (CircuitResult, QuantumState)):
            target = target.state()
        np.testing.assert_allclose(self.to_numpy(value), self.to_numpy(target), rtol=rtol, atol=atol)
