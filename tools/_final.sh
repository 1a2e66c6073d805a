# full GPU suite, smoke, default bench line (N=1)
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_final_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err
cat gpurun_out/r02_final_tests.log gpurun_out/r02_final_smoke.log | tail -12; cat gpurun_out/r02_bench_n1.json
