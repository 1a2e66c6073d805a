timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -5
timeout 120 examples/_build/crowd_host 10000 20000 128 10 0 1 2>&1 | tail -2
timeout 120 examples/_build/crowd_host 10000 20000 128 10 0 2>&1 | tail -2
