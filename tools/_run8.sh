timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
for k in ('value','ms_per_step','e2e','fused_gather','gathered','check'):
    print(k, json.dumps(d.get(k))[:1500])
PY


