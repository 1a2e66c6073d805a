timeout 300 python -m pytest tests -m gpu -q -x -k "run_kernel and texcoord" 2>&1 | tail -3 > gpurun_out/s18_tests.log
for c in 1 2; do P3CS_RUN_COPIES=$c P3CS_DEBUG_PLAN=1 timeout 200 python tools/kbench.py --configs c3uv 2>&1 | cut -c1-140 >> gpurun_out/s18_k.log; done
cat gpurun_out/s18_tests.log gpurun_out/s18_k.log
