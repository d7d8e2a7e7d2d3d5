export B200LP_SPIN_TIMEOUT_MS=8000
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variant or look or random_shapes or iteration_limit or full_size" 2>&1 | tail -5 > gpurun_out/r02_n_pytest.log
tail -4 gpurun_out/r02_n_pytest.log
timeout 900 python tools/loop_ab.py --shapes cfg3,slab8,cfg5 --variants 10 --iters 1000 --tag r02_n_look2 > gpurun_out/r02_n.log 2>&1
B200LP_LOOK=1 timeout 900 python tools/loop_ab.py --shapes cfg3,slab8,cfg5 --variants 10 --iters 1000 --tag r02_n_look1 >> gpurun_out/r02_n.log 2>&1
B200LP_LOOK_CTAS=8 timeout 900 python tools/loop_ab.py --shapes slab8 --variants 10,13,11 --iters 1000 --tag r02_n_look2_g8 >> gpurun_out/r02_n.log 2>&1
cut -c1-330 gpurun_out/r02_n.log
