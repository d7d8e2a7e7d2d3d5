export B200LP_SPIN_TIMEOUT_MS=8000
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_k_pytest.log
tail -8 gpurun_out/r02_k_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -2
