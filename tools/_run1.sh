export B200LP_SPIN_TIMEOUT_MS=8000
timeout 900 python bench.py > gpurun_out/r02_j_bench_n1.json 2> gpurun_out/r02_j_bench_n1.err
tail -3 gpurun_out/r02_j_bench_n1.err; cat gpurun_out/r02_j_bench_n1.json
