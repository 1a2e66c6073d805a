# typed-column parity on the GPU + a regression pass over the existing parity tests + strip-kernel timing check
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q 2>&1 | grep -v "^E  \|^$" | tail -40 > gpurun_out/r02_typed_tests.log
cat gpurun_out/r02_typed_tests.log

