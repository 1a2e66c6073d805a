# evidence of the two-copy build: bench line, launch list, full capture of the run kernel; plus a look at K / trim with the new plan
timeout 900 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
B="python bench.py --steps 3 --warmup 3 --no-cpu"
timeout 600 $B > gpurun_out/r02b_plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_bench_steps3_warmup3.csv $B > gpurun_out/r02b_ncu_launches.log 2>&1
K="python tools/kbench.py --configs c3 --steps 1 --warmup 1"
timeout 300 $K > gpurun_out/r02b_plain_kbench.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"run_kernel" -c 2 -o gpurun_out/r02b_run_kernel $K > gpurun_out/r02b_ncu_full.log 2>&1
for plan in 1120,2,12 1120,2,14 1120,2,16 1120,2,10; do echo "== plan $plan" >> gpurun_out/s7_k.log; P3CS_RUN_PLAN=$plan timeout 200 python tools/kbench.py --configs c3 >> gpurun_out/s7_k.log 2>&1; done
for trim in 0 2 8; do echo "== trim $trim" >> gpurun_out/s7_k.log; P3CS_RUN_TRIM=$trim timeout 200 python tools/kbench.py --configs c3 >> gpurun_out/s7_k.log 2>&1; done
cat gpurun_out/r02b_bench_n1.json | cut -c1-600; tail -2 gpurun_out/r02b_ncu_full.log; sed 's/"vertices".*//' gpurun_out/s7_k.log; ls -la gpurun_out/*.ncu-rep
