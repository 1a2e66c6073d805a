timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-pose > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -5 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
for k in ('value','ms_per_step','e2e','gather','fused_gather','gathered','check'):
    print(k, json.dumps(d.get(k))[:900])
PY
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3
