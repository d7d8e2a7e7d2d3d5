export B200LP_SPIN_TIMEOUT_MS=8000
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -k "two_phase" 2>&1 | tail -12 > gpurun_out/r02_s_pytest_2gpu.log
tail -12 gpurun_out/r02_s_pytest_2gpu.log
