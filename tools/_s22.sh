timeout 600 python -m pytest tests -m gpu -q -x -k "run_kernel or full_size or fullsize" 2>&1 | tail -3 > gpurun_out/s22_tests.log
P3CS_DEBUG_PLAN=1 timeout 200 python tools/kbench.py --configs c3,c3uv,c3tb,c3rand 2>&1 | cut -c1-200 > gpurun_out/s22_k.log
timeout 200 python tools/kbench.py --configs c3 --instances 1250 2>&1 | cut -c1-100 >> gpurun_out/s22_k.log
cat gpurun_out/s22_tests.log gpurun_out/s22_k.log
