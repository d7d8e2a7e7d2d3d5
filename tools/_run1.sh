export B200LP_SPIN_TIMEOUT_MS=8000
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_a_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -x --deselect tests/test_gpu_parity.py::test_full_size_configs_prefix_bit_exact_then_invariants 2>&1 | tail -40 > gpurun_out/r02_a_pytest.log
timeout 900 python tools/loop_ab.py --shapes cfg3,slab8,cfg2,small --variants 10,2,20,21 --iters 400 --tag r02_a_loop_ab > gpurun_out/r02_a_loop_ab.log 2>&1
tail -5 gpurun_out/r02_a_pytest.log; cat gpurun_out/r02_a_loop_ab.log
