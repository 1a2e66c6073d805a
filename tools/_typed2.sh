timeout 900 python -m pytest tests/test_panda_integration.py tests/test_gpu_parity.py -m gpu -q -k "integration or golden or shim or panda" 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r02_typed2_tests.log
cat gpurun_out/r02_typed2_tests.log
