M=integration/_build/models
timeout 120 integration/_build/panda_cuda_check panda3d_b200 $M/panda-model.egg $M/panda-walk4.egg 0 7 10 20 33 45 60 2>&1 | tail -1
timeout 120 integration/_build/panda_cuda_check panda3d_b200 $M/ripple.egg $M/ripple.egg 0 3 7 20 31 44 2>&1 | tail -1
timeout 300 integration/_build/panda_cuda_check panda3d_b200 $M/panda-model.egg $M/panda-walk4.egg --crowd 1000 8 2>&1 | tail -1
timeout 300 python tools/latency.py 2>&1 | tail -4
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
