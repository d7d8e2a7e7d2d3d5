export B200LP_SPIN_TIMEOUT_MS=8000
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variant or look_grid or random_shapes or full_solve_bit_exact or iteration_limit or config2 or degenerate" 2>&1 | tail -5 > gpurun_out/r02_h_pytest.log
timeout 900 python tools/loop_ab.py --shapes slab8,cfg5 --variants 20 --iters 2000 --tag r02_h_y > gpurun_out/r02_h.log 2>&1
timeout 900 python tools/loop_ab.py --shapes cfg2,small --variants 21 --look 1,4 --iters 400 --tag r02_h_z >> gpurun_out/r02_h.log 2>&1
tail -3 gpurun_out/r02_h_pytest.log; cut -c1-600 gpurun_out/r02_h.log
