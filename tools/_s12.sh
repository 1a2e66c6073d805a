timeout 900 python -m pytest tests -m gpu -q -x -k "run_kernel or fullsize or full_size or crowd or edge or extreme" 2>&1 | tail -4 > gpurun_out/s12_tests.log
timeout 200 python tools/kbench.py --configs c3,c3uv,c3tb,c3rand > gpurun_out/s12_k.log 2>&1
timeout 200 python tools/kbench.py --configs c3 --instances 1250 >> gpurun_out/s12_k.log 2>&1
cat gpurun_out/s12_tests.log; sed 's/"vertices".*//' gpurun_out/s12_k.log | cut -c1-330
