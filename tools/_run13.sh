for cpi in 1 2 3 4; do echo "== cpi $cpi" >> gpurun_out/r2_k6.log; P3CS_RUN_CPI=$cpi timeout 200 python tools/kbench.py --configs c3,c3uv,c3tb >> gpurun_out/r2_k6.log 2>&1; done
echo "== 1250 instances" >> gpurun_out/r2_k6.log
for cpi in 1 2 3 4 6; do echo "== cpi $cpi" >> gpurun_out/r2_k6.log; P3CS_RUN_CPI=$cpi timeout 200 python tools/kbench.py --configs c3 --instances 1250 >> gpurun_out/r2_k6.log 2>&1; done
sed 's/"vertices".*//' gpurun_out/r2_k6.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
