export B200LP_SPIN_TIMEOUT_MS=8000
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variant or look_grid or random_shapes or full_solve_bit_exact or iteration_limit or config2 or degenerate" 2>&1 | tail -5 > gpurun_out/r02_e_pytest.log
export B200LP_TILE_PROFILE=gpurun_out/r02_e_tileprof_slab8.csv
timeout 900 python tools/loop_ab.py --shapes slab8 --variants 10,20 --look 0,8 --iters 2000 --tag r02_e_y > gpurun_out/r02_e.log 2>&1
unset B200LP_TILE_PROFILE
timeout 900 python tools/loop_ab.py --shapes slab8c4,cfg5 --variants 10,20 --iters 1000 --tag r02_e_w >> gpurun_out/r02_e.log 2>&1
timeout 900 python tools/loop_ab.py --shapes cfg2,small --variants 2,21 --look 1 --iters 400 --tag r02_e_z >> gpurun_out/r02_e.log 2>&1
tail -3 gpurun_out/r02_e_pytest.log; cut -c1-600 gpurun_out/r02_e.log
