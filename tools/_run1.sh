export B200LP_SPIN_TIMEOUT_MS=8000
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r02_l_pytest.log
tail -12 gpurun_out/r02_l_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -2
python tools/small_solve_latency.py 2>&1 | tail -5
B200LP_SMALL=0 python tools/small_solve_latency.py 2>&1 | tail -5
