python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_t3.log
tail -3 gpurun_out/r2_t3.log
echo "== 4 CTAs/SM" > gpurun_out/r2_k1.log
python tools/kbench.py --configs c3,c3rand,c3fixed5,c3tbuv >> gpurun_out/r2_k1.log 2>&1
P3CS_RUN_PLAN=1008,2,12 python tools/kbench.py --configs c3 >> gpurun_out/r2_k1.log 2>&1
cp panda3d_b200/libp3cudaskin.so /tmp/keep.so; cp panda3d_b200/libp3cudaskin_v3.so panda3d_b200/libp3cudaskin.so
echo "== 3 CTAs/SM" >> gpurun_out/r2_k1.log
python tools/kbench.py --configs c3,c3rand,c3fixed5,c3tbuv >> gpurun_out/r2_k1.log 2>&1
P3CS_RUN_PLAN=1424,2,12 python tools/kbench.py --configs c3 >> gpurun_out/r2_k1.log 2>&1
P3CS_RUN_PLAN=944,3,12 python tools/kbench.py --configs c3,c3fixed5 >> gpurun_out/r2_k1.log 2>&1
cp /tmp/keep.so panda3d_b200/libp3cudaskin.so
grep -v "^run plan" gpurun_out/r2_k1.log | cut -c1-100
