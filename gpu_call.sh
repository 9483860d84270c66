timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
