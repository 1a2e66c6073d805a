timeout 900 python -m pytest tests -m gpu -q -x -k "run_kernel or fullsize or full_size or non_affine or crowd or edge or golden or panda" 2>&1 | tail -6 > gpurun_out/s6_tests.log
for c in 1 2; do echo "== copies $c" >> gpurun_out/s6_k.log; P3CS_RUN_COPIES=$c timeout 300 python tools/kbench.py --configs c3,c3rand >> gpurun_out/s6_k.log 2>&1; done
cat gpurun_out/s6_tests.log; sed 's/"vertices".*//' gpurun_out/s6_k.log
