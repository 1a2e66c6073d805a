export P3CS_PATH=run
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_t3.log
tail -3 gpurun_out/r2_t3.log
timeout 200 python tools/kbench.py --configs c3,c3rand,c3fixed5,c3,c3tbuv,c3uv,c3tb,c2x,c3 > gpurun_out/r2_k1.log 2>&1
cut -c1-330 gpurun_out/r2_k1.log | sed 's/"vertices".*//'
