export P3CS_PATH=run
cp panda3d_b200/libp3cudaskin.so /tmp/keep.so
echo "== 256x3" > gpurun_out/r2_k3.log
P3CS_DEBUG_PLAN=1 python tools/kbench.py --configs c3,c3fixed5,c3rand,c3tbuv,c3tb >> gpurun_out/r2_k3.log 2>&1
for v in 384x2 192x4 512x1; do
  cp panda3d_b200/libvar_$v.so panda3d_b200/libp3cudaskin.so
  echo "== $v" >> gpurun_out/r2_k3.log
  P3CS_DEBUG_PLAN=1 python tools/kbench.py --configs c3,c3fixed5,c3rand,c3tbuv,c3tb >> gpurun_out/r2_k3.log 2>&1
done
cp /tmp/keep.so panda3d_b200/libp3cudaskin.so
sed 's/"vertices".*//; s/items of.*lockstep/lockstep/; s/(20000 lane.*//' gpurun_out/r2_k3.log
