mkdir -p gpurun_out
for n in 29 33; do timeout 120 python scripts/dist_local_repro.py $n 8 0 --check > gpurun_out/r4d_repro$n.log 2>&1; done
timeout 200 python scripts/dist_local_repro.py 35 8 0 > gpurun_out/r4d_repro35.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r4d_pytest.log 2>&1
echo finished
