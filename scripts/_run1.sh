timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmarked.py -q -k "rows_ or c1 or C1" 2>&1 | tail -5
