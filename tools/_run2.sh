export B200LP_SPIN_TIMEOUT_MS=8000
for look in 2 1; do
B200LP_LOOK=$look timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 10 --no-cfg4 --no-e2e > gpurun_out/r02_p_bench_n2_look$look.json 2> gpurun_out/r02_p_bench_n2_look$look.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_p_bench_n2_look$look.json").read().strip().splitlines()[-1])
print("look$look", d["value"], d["ms_per_step"], d["overlapped"])
PY
done
