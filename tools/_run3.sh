export P3CS_PATH=run
cp panda3d_b200/libp3cudaskin.so /tmp/keep.so
echo "== 256x3" > gpurun_out/r2_k3.log
timeout 100 python tools/kbench.py --configs c3,c3fixed5,c3rand,c3tb >> gpurun_out/r2_k3.log 2>&1
for plan in 832,3,12 624,4,12 1248,2,9; do echo "plan $plan" >> gpurun_out/r2_k3.log; P3CS_RUN_PLAN=$plan timeout 100 python tools/kbench.py --configs c3 >> gpurun_out/r2_k3.log 2>&1; done
for v in 224x4 288x3 320x3; do
  cp panda3d_b200/libvar_$v.so panda3d_b200/libp3cudaskin.so
  echo "== $v" >> gpurun_out/r2_k3.log
  timeout 100 python tools/kbench.py --configs c3,c3fixed5,c3rand,c3tb >> gpurun_out/r2_k3.log 2>&1
done
cp /tmp/keep.so panda3d_b200/libp3cudaskin.so
sed 's/"vertices".*//; s/items of.*lockstep/lockstep/; s/(20000 lane.*//' gpurun_out/r2_k3.log
