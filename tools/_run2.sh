export P3CS_PATH=run
CMD="python tools/kbench.py --configs c3 --steps 1 --warmup 1 --instances 2960"
$CMD > gpurun_out/r2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:run_kernel -c 2 -o gpurun_out/r2_run_v3 $CMD > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
