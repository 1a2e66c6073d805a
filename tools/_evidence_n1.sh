# final-state evidence (N=1): GPU tests, smoke, bench line, reference arm, launch list, full capture of the run kernel
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err
B="python bench.py --steps 3 --warmup 3 --no-cpu"
timeout 600 $B > gpurun_out/r02_plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_steps3_warmup3.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
K="python tools/kbench.py --configs c3,c3uv --steps 1 --warmup 1"
timeout 300 $K > gpurun_out/r02_plain_kbench.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"run_kernel" -c 4 -o gpurun_out/r02_run_kernel_final $K > gpurun_out/r02_ncu_full.log 2>&1
cat gpurun_out/r02_gpu_tests.log gpurun_out/r02_smoke.log; cut -c1-300 gpurun_out/r02_bench_n1.json; cut -c1-400 gpurun_out/r02_bench_ref_n1.json; tail -2 gpurun_out/r02_ncu_full.log
