export B200LP_SPIN_TIMEOUT_MS=8000
nvidia-smi -L > gpurun_out/r02_q_gpus.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_q_pytest_8gpu.log
tail -6 gpurun_out/r02_q_pytest_8gpu.log
B200LP_LOOK=2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2000 --warmup 10 > gpurun_out/r02_q_bench_n8_look2.json 2> gpurun_out/r02_q_bench_n8_look2.err
B200LP_LOOK=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 2000 --warmup 10 --no-e2e --no-cfg4 > gpurun_out/r02_q_bench_n8_look1.json 2> gpurun_out/r02_q_bench_n8_look1.err
python - <<PY
import json
for look in (2,1):
    try:
        d=json.loads(open(f"gpurun_out/r02_q_bench_n8_look{look}.json").read().strip().splitlines()[-1])
        print("look",look, d["value"], d["ms_per_step"], d["overlapped"]["ms_look"], d["overlapped"]["us_look_split_max_over_ranks"], d["parity"], d["cfg4"], d["e2e"]["value"] if d["e2e"] else None)
    except Exception as e:
        print("look",look,"failed",e)
PY
tail -3 gpurun_out/r02_q_bench_n8_look2.err
