mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu -k "sharded_program and (qft-16 or zoo)" > gpurun_out/r4j_dist_pytest.log 2>&1
QB_ALLTOALL_MIN=1 QB_ALLTOALL_PUSH=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r4j_bench2_push.log 2>&1
echo finished
