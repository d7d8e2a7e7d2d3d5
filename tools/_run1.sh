export B200LP_SPIN_TIMEOUT_MS=8000
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variant or look" 2>&1 | tail -4 > gpurun_out/r02_r_pytest.log
tail -3 gpurun_out/r02_r_pytest.log
timeout 900 python tools/loop_ab.py --shapes cfg3,slab8,cfg5,cfg2 --variants 10,30 --iters 1000 --tag r02_r_bulk_ab > gpurun_out/r02_r.log 2>&1
cut -c1-330 gpurun_out/r02_r.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cfg4 --no-parity --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_ncu_launches_cfg3.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_iter2 -s 6 -c 2 -o gpurun_out/r02_ncu_full_k_iter2_cfg3 $B > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r02_ncu_full.log
