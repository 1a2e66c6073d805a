timeout 600 python -m pytest tests -m gpu -q -x -k "staged or bam or registered or mirror or panda or shim" 2>&1 | tail -6 > gpurun_out/s2_tests.log
for st in 0 1; do echo "== P3CS_STAGE=$st" >> gpurun_out/s2_breakdown.log; P3CS_STAGE=$st timeout 300 python tools/call_breakdown.py 2>&1 | grep animate_vertices >> gpurun_out/s2_breakdown.log; done
M=integration/_build/models
for st in 0 1; do P3CS_STAGE=$st timeout 120 integration/_build/panda_cuda_check panda3d_b200 $M/panda-model.egg $M/panda-walk4.egg 0 7 10 20 33 45 60 2>&1 | tail -1 | sed 's/.*"cpu_us_steady"/"cpu_us_steady"/; s/"frames".*//' >> gpurun_out/s2_breakdown.log; done
cat gpurun_out/s2_tests.log gpurun_out/s2_breakdown.log
