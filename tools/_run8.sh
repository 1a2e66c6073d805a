timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/pcie_probe.py > gpurun_out/r2_pcie_probe.log 2>&1
grep "^{" gpurun_out/r2_pcie_probe.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
for k in ('value','ms_per_step','e2e','gather','fused_gather','gathered','check','joints_on_gpu'):
    print(k, json.dumps(d.get(k))[:1100])
PY
python -c "
import json; b=json.load(open('gpurun_out/pcie_box.json')); print(b['topo'][:1500]); print(b['meminfo_nodes']); print(b['numactl'][:600]); print('\n'.join(l for l in b['lscpu'].splitlines() if any(k in l for k in ('Model name','Socket','NUMA','CPU(s):','Thread'))))"
