"""2-GPU probe of the NVLink half-shard swap kernel:  torchrun --nproc-per-node 2 scripts/p2p_probe.py [nlocal]"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qibo_b200.distributed import PeerShard  # noqa: E402
from qibo_b200.engine import Engine  # noqa: E402

rank = int(os.environ["RANK"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 30
eng = Engine(rank)
ps = PeerShard(eng, nl, "complex128")
ps.tensor.fill_(rank + 1)
peer = ps.peer_ptr[rank ^ 1]
res = []
for blocks in ("4", "8", "16", "32"):
    for unroll in ("2", "4", "8"):
        for lq in (0, 7):  # local qubit 0 = top bit (contiguous halves); 7 = strided chunks
            os.environ["QB_P2P_BLOCKS_PER_SM"], os.environ["QB_P2P_UNROLL"] = blocks, unroll
            ts = []
            for _ in range(3):
                ps.fence()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.swap_half_p2p(ps.array, peer, nl, lq, rank, rank, 2)
                e1.record()
                ps.fence()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            gbs = 16 * 2 ** (nl - 1) / (min(ts) * 1e-3) / 1e9
            res.append({"blocks_per_sm": int(blocks), "unroll": int(unroll), "local_qubit": lq, "ms": min(ts), "GBps_per_direction": gbs})
            if rank == 0:
                print(res[-1], flush=True)
# reference: torch peer copy of the same half
t = ps.tensor
half = t.numel() // 2
other = torch.empty(half, dtype=t.dtype, device=f"cuda:{rank}")
dist.barrier()
if rank == 0:
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "p2p_probe.json"), "w"), indent=1)
dist.destroy_process_group()
