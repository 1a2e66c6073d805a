# extended seeded fuzz over formats with the run kernel forced (two palette copies forced / library's choice) and by default
P3CS_FUZZ_CASES=12000 P3CS_PATH=run P3CS_RUN_COPIES=2 timeout 900 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x -n 6 2>&1 | tail -2 > gpurun_out/r02_fuzz_run_two_copies.log
P3CS_FUZZ_CASES=6000 P3CS_PATH=run timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x -n 6 2>&1 | tail -2 > gpurun_out/r02_fuzz_run.log
P3CS_FUZZ_CASES=6000 timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x -n 6 2>&1 | tail -2 > gpurun_out/r02_fuzz_default.log
for f in gpurun_out/r02_fuzz_*.log; do echo "== $f"; cat $f; done
