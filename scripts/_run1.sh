timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "finish_and_copy_variants" 2>&1 | tail -15
