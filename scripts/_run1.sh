timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmarked.py -q -x -n 2 -k "kfc or factor or Factor or sweep" 2>&1 | tail -3
python scripts/one_sweep.py 1 10
python scripts/one_sweep.py 8 10
