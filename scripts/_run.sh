mkdir -p gpurun_out
timeout 100 python -m pytest tests -q -m gpu > gpurun_out/r4l_pytest.log 2>&1
echo finished
