export P3CS_PATH=run
for trim in 2 3 4 5; do
echo "== trim $trim" >> gpurun_out/r2_k4.log
P3CS_RUN_TRIM=$trim timeout 300 python tools/kbench.py --configs c3,c3uv,c3tb >> gpurun_out/r2_k4.log 2>&1
done
for plan in 1248,2,11 1248,2,10 1152,2,12 1200,2,12; do echo "== plan $plan" >> gpurun_out/r2_k4.log
P3CS_RUN_PLAN=$plan timeout 300 python tools/kbench.py --configs c3 >> gpurun_out/r2_k4.log 2>&1
done
sed 's/"vertices".*//' gpurun_out/r2_k4.log
