cp panda3d_b200/libp3cudaskin.so /tmp/keep.so
export P3CS_PATH=run
for v in 256x3 320x2 192x3 288x2; do
  cp panda3d_b200/libvar_$v.so panda3d_b200/libp3cudaskin.so
  echo "== $v" >> gpurun_out/r2_k3.log
  P3CS_DEBUG_PLAN=1 timeout 100 python tools/kbench.py --configs c3tbuv,c3rand,c2x >> gpurun_out/r2_k3.log 2>&1
done
cp /tmp/keep.so panda3d_b200/libp3cudaskin.so
sed 's/"vertices".*//; s/items of.*lockstep/lockstep/; s/(20000 lane.*//' gpurun_out/r2_k3.log
