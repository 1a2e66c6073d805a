timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_t5.log; cat gpurun_out/r2_t5.log
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -3 gpurun_out/r2_bench.err; cut -c1-1500 gpurun_out/r2_bench.json
