for path in strip run; do for c in 1 2; do echo "== $path copies $c" >> gpurun_out/s21_k.log; P3CS_PATH=$path P3CS_RUN_COPIES=$c P3CS_DEBUG_PLAN=1 timeout 200 python tools/kbench.py --configs c3tbuv,c5x 2>&1 | cut -c1-160 >> gpurun_out/s21_k.log; done; done
cat gpurun_out/s21_k.log
