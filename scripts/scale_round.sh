#!/bin/bash
# Multi-GPU measurement set (run under `gpurun --gpus N`):  bash scripts/scale_round.sh <tag> <N> [variants]
# One bench line per transport variant, each checked against the closed form before it is timed (bench.py "verify"):
#   default                 out-of-place all-to-all (remote stores), K8 permutation afterwards
#   swap                    QB_ALLTOALL_PUSH=0: in-place chunk swaps (remote loads + stores)
#   pairwise                QB_NO_ALLTOALL=1: one half-shard exchange kernel per global qubit
#   fuseperm (EXPERIMENTAL) QB_A2A_FUSE_PERM=1: the closing permutation written by the all-to-all (one K8 per chunk)
tag=${1:-rX}
n=${2:-2}
variants=${3:-"default swap pairwise fuseperm"}
out=gpurun_out
mkdir -p $out
port=29700
for v in $variants; do
  case $v in
    default) envs="" ;;
    swap) envs="QB_ALLTOALL_PUSH=0" ;;
    pairwise) envs="QB_NO_ALLTOALL=1" ;;
    fuseperm) envs="QB_A2A_FUSE_PERM=1" ;;
    *) echo "unknown variant $v"; continue ;;
  esac
  port=$((port + 1))
  env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 3 --warmup 3 > $out/${tag}_bench_${n}gpu_${v}.log 2>&1
  grep '^{' $out/${tag}_bench_${n}gpu_${v}.log | tail -1 > $out/${tag}_bench_${n}gpu_${v}.json
  python - "$out/${tag}_bench_${n}gpu_${v}.json" "$v" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    x = d.get("exchange", {})
    print(f"{sys.argv[2]:9s} {d['ms_per_step']:8.1f} ms/circuit  value {d['value']:9.0f}  exchange {x.get('ms_per_step_rank0', 0):6.1f} ms "
          f"({x.get('GBps_per_direction_rank0', 0):.0f} GB/s/dir, {x.get('launches_per_step')} launches)  verify {d.get('verify', {}).get('exchange_path')} "
          f"{d.get('verify', {}).get('max_rel_err')}")
except Exception as e:
    print(sys.argv[2], "FAILED:", e)
PY
done
