for c in 1 2 1 2; do echo "== copies $c" >> gpurun_out/s10_k.log; P3CS_RUN_COPIES=$c P3CS_DEBUG_PLAN=1 timeout 200 python tools/kbench.py --configs c3,c3rand 2>&1 | cut -c1-120 >> gpurun_out/s10_k.log; done
cat gpurun_out/s10_k.log
