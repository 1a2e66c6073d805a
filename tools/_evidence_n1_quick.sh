# final state, without the ncu passes: GPU tests, smoke, bench line
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -2 gpurun_out/r02_gpu_tests.log; cat gpurun_out/r02_smoke.log; cut -c1-330 gpurun_out/r02_bench_n1.json
