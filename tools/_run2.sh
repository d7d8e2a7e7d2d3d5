export B200LP_SPIN_TIMEOUT_MS=8000
nvidia-smi -L > gpurun_out/r02_i_gpus.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_i_pytest_2gpu.log
tail -15 gpurun_out/r02_i_pytest_2gpu.log
for loop in iter persist; do
B200LP_LOOP=$loop timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 10 --no-e2e > gpurun_out/r02_i_bench_n2_$loop.json 2> gpurun_out/r02_i_bench_n2_$loop.err
cut -c1-1500 gpurun_out/r02_i_bench_n2_$loop.json; tail -3 gpurun_out/r02_i_bench_n2_$loop.err
done
