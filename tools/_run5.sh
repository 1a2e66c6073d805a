timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_t5.log; cat gpurun_out/r2_t5.log
timeout 300 python tools/kbench.py --configs c3,c3uv,c3tb,c3tbuv,c2x,c5x > gpurun_out/r2_k5.log 2>&1; sed 's/"vertices".*//' gpurun_out/r2_k5.log
