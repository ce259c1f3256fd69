timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "cycle" 2>&1 | grep -v "^$" | tail -30
