"""Global-qubit distributed execution (one process per GPU) -- see DESIGN.md section 6."""

import torch


def world_size():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_world_size()
    return 1


def execute_circuit(backend, circuit, initial_state=None, nshots=None):  # pragma: no cover
    raise NotImplementedError("multi-rank execution is wired up in qibo_b200.distributed (work in progress)")
