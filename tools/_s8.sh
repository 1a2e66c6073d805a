for ns in 0 100 200 400 800 0; do echo "== psleep $ns" >> gpurun_out/s8_k.log; P3CS_RUN_PSLEEP=$ns timeout 200 python tools/kbench.py --configs c3,c3uv >> gpurun_out/s8_k.log 2>&1; done
echo "== one copy, tiles of 1120" >> gpurun_out/s8_k.log; P3CS_RUN_COPIES=1 P3CS_RUN_PLAN=1120,2,12 timeout 200 python tools/kbench.py --configs c3 >> gpurun_out/s8_k.log 2>&1
echo "== one copy, default" >> gpurun_out/s8_k.log; P3CS_RUN_COPIES=1 timeout 200 python tools/kbench.py --configs c3 >> gpurun_out/s8_k.log 2>&1
sed 's/"vertices".*//' gpurun_out/s8_k.log
