cp panda3d_b200/libp3cudaskin.so /tmp/keep.so
cp panda3d_b200/libvar_checks.so panda3d_b200/libp3cudaskin.so
(echo "# debug build (P3CS_DEFINES=P3CS_RUN_CHECKS: device-side bounds checks of every item of the run kernel), tools/sanitize_cases.py and the run-kernel parity tests"; timeout 300 python tools/sanitize_cases.py 2>&1 | tail -12; P3CS_PATH=run timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "run_kernel or full_size or synthetic or golden" 2>&1 | tail -3) > gpurun_out/r02_run_kernel_checked_build.log 2>&1
cat gpurun_out/r02_run_kernel_checked_build.log | tail -8
cp /tmp/keep.so panda3d_b200/libp3cudaskin.so
for rep in 1 2; do for v in A B; do
  cp panda3d_b200/libvar_$v.so panda3d_b200/libp3cudaskin.so
  echo "== $v" >> gpurun_out/r2_ab.log
  timeout 200 python tools/kbench.py --configs c3,c3fixed5,c3uv,c3tb >> gpurun_out/r2_ab.log 2>&1
done; done
cp /tmp/keep.so panda3d_b200/libp3cudaskin.so
sed 's/"vertices".*//' gpurun_out/r2_ab.log
