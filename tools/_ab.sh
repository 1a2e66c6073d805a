# A/B of two builds of the library on the same box: tools/_ab.sh <configs>  (libvar_A.so, libvar_B.so beside the library)
export P3CS_PATH=run
cp panda3d_b200/libp3cudaskin.so /tmp/keep.so
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -3
for rep in 1 2; do for v in A B; do
  cp panda3d_b200/libvar_$v.so panda3d_b200/libp3cudaskin.so
  echo "== $v" >> gpurun_out/r2_ab.log
  timeout 200 python tools/kbench.py --configs $1 >> gpurun_out/r2_ab.log 2>&1
done; done
cp /tmp/keep.so panda3d_b200/libp3cudaskin.so
sed 's/"vertices".*//' gpurun_out/r2_ab.log
