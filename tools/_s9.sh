timeout 900 python -m pytest tests -m gpu -q -x -k "run_kernel or fullsize or full_size or non_affine or crowd or edge or golden" 2>&1 | tail -6 > gpurun_out/s9_tests.log
P3CS_DEBUG_PLAN=1 timeout 200 python tools/kbench.py --configs c3,c3uv,c3tb,c3rand > gpurun_out/s9_k.log 2>&1
timeout 200 python tools/kbench.py --configs c3 --instances 1250 >> gpurun_out/s9_k.log 2>&1
cat gpurun_out/s9_tests.log; sed 's/"vertices".*//' gpurun_out/s9_k.log | cut -c1-330
