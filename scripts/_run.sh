mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r4i_bench8.log 2>&1
echo finished
