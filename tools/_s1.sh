# session 2 call 1: state of HEAD on a fresh box
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/s1_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s1_smoke.log 2>&1
timeout 300 python tools/call_breakdown.py > gpurun_out/s1_breakdown.log 2>&1
timeout 300 python tools/latency.py > gpurun_out/s1_latency.log 2>&1
M=integration/_build/models
timeout 120 integration/_build/panda_cuda_check panda3d_b200 $M/panda-model.egg $M/panda-walk4.egg 0 7 10 20 33 45 60 > gpurun_out/s1_panda.log 2>&1
timeout 200 python tools/kbench.py --configs c3 > gpurun_out/s1_k.log 2>&1
timeout 200 python tools/kbench.py --configs c3 --instances 1250 >> gpurun_out/s1_k.log 2>&1
cat gpurun_out/s1_tests.log gpurun_out/s1_smoke.log gpurun_out/s1_breakdown.log gpurun_out/s1_latency.log; tail -3 gpurun_out/s1_panda.log; sed 's/"vertices".*//' gpurun_out/s1_k.log
