B="python bench.py --steps 3 --warmup 3 --no-cpu"
timeout 600 $B > gpurun_out/r02_plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_steps3_warmup3.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
K="python tools/kbench.py --configs c3,c3tbuv,c3a16tbuv --steps 1 --warmup 1"
timeout 300 $K > gpurun_out/r02_plain_kbench.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"run_kernel|strip_kernel" -c 6 -o gpurun_out/r02_hot_kernels $K > gpurun_out/r02_ncu_full.log 2>&1
P="python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --no-check"
timeout 300 $P > gpurun_out/r02_plain_pose.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:pose_kernel -s 2 -c 1 -o gpurun_out/r02_pose_kernel $P > gpurun_out/r02_ncu_pose.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log gpurun_out/r02_ncu_pose.log gpurun_out/r02_ncu_launches.log
ls -la gpurun_out/*.ncu-rep
