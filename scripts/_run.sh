mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu -k sharded_program > gpurun_out/r4h_dist_pytest.log 2>&1
QB_ALLTOALL_MIN=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r4h_bench2.log 2>&1
echo finished
