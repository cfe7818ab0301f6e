timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_distributed.py 2>&1 | tail -3
