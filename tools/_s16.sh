timeout 900 python -m pytest tests -m gpu -q -x -k "bounds or run_kernel or edge or listed" 2>&1 | tail -4 > gpurun_out/s16_tests.log
timeout 600 python bench.py --no-cpu --no-configs --steps 10 --warmup 3 > gpurun_out/s16_bench.json 2> gpurun_out/s16_bench.err
cat gpurun_out/s16_tests.log; python -c "
import json; d=json.load(open('gpurun_out/s16_bench.json')); print(d['ms_per_step'], d['tight_bounds'])"
