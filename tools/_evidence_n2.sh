timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_multi_gpu_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
cat gpurun_out/r02_multi_gpu_tests.log; cut -c1-300 gpurun_out/r02_bench_n2.json; tail -3 gpurun_out/r02_bench_n2.err
