export B200LP_SPIN_TIMEOUT_MS=8000
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "end_state" 2>&1 | tail -4 > gpurun_out/r02_t_pytest.log
tail -3 gpurun_out/r02_t_pytest.log
timeout 600 python -m pytest tests/test_gpu_solver_hook.py -m gpu -q --durations=6 2>&1 | tail -12 > gpurun_out/r02_t_hook_small_on.log
B200LP_SMALL=0 timeout 600 python -m pytest tests/test_gpu_solver_hook.py -m gpu -q --durations=6 2>&1 | tail -12 > gpurun_out/r02_t_hook_small_off.log
grep -E "s call|passed|failed" gpurun_out/r02_t_hook_small_on.log; echo ---; grep -E "s call|passed|failed" gpurun_out/r02_t_hook_small_off.log
timeout 900 python tools/two_phase_bench.py 8192 8192 1 > gpurun_out/r02_t_two_phase.jsonl 2>&1
cat gpurun_out/r02_t_two_phase.jsonl
