timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
cut -c1-300 gpurun_out/r02_bench_n8.json; cut -c1-300 gpurun_out/r02_bench_n4.json; tail -3 gpurun_out/r02_bench_n8.err
