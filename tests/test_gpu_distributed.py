"""Multi-GPU parity: ShardedProgram over NCCL (one process per GPU) against the single-process oracle.
Needs >= 2 GPUs (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_distributed.py -m gpu`)."""

import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, n, dtype, out):
    sys.path[:0] = [ROOT, HERE]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from helpers import ops_from_named, oracle_run, rand_state, random_zoo
        from oracle import numpy_oracle as orc
        from qibo_b200.distributed import ShardedProgram
        from qibo_b200.engine import Engine

        eng = Engine(rank)
        if case == "qft":
            ops = ops_from_named(orc.qft_ops(n))
        elif case == "variational":
            ops = ops_from_named(orc.variational_ops(n, 2, np.random.default_rng(1).random(4 * n) * 6))
        else:
            ops = random_zoo(n, 50, 5, max_dense=4)
        psi = rand_state(n, 11, dtype)
        prog = ShardedProgram(eng, n, dtype, ops, staging_elems=1 << 10)
        shard = prog.scatter(psi)
        stats = prog.run(shard)
        full = prog.gather(shard)
        ref = oracle_run(psi, ops, n)
        err = float(np.abs(full - ref).max())
        # zero-state constructor
        z = prog.basis_state(0)
        prog.run(z)
        zfull = prog.gather(z)
        zref = oracle_run(orc.zero_state(n, dtype), ops, n)
        err = max(err, float(np.abs(zfull - zref).max()))
        # the single-kernel NVLink peer-memory exchange must give the same state as the NCCL path
        ps = prog.peer_shard(None)
        ps.tensor.copy_(prog.scatter(psi).tensor)
        st2 = prog.run(ps)
        pfull = prog.gather(ps)
        err = max(err, float(np.abs(pfull - ref).max()))
        assert st2.nexchanges == stats.nexchanges
        if rank == 0:
            out.put((err, stats.nexchanges, stats.nsweeps))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case,n,dtype", [("qft", 16, "complex128"), ("qft", 17, "complex64"), ("variational", 15, "complex64"), ("zoo", 15, "complex128")])
def test_sharded_program_on_gpus(case, n, dtype):
    import torch.multiprocessing as mp

    world = 4 if torch.cuda.device_count() >= 4 else 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, n, dtype, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    err, nex, nsweeps = out.get()
    assert err < (1e-12 if dtype == "complex128" else 1e-5)
    assert nex >= 1 and nsweeps >= 1
