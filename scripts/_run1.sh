timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmarked.py -q -k "rows_ or c1 or C1" 2>&1 | tail -2
timeout 100 python scripts/time_conv_c1.py 2>&1 | head -3
EINCONV_ROWS_PAIR=1 timeout 100 python scripts/time_conv_c1.py 2>&1 | head -3
EINCONV_ROWS_NO_SPEC=1 timeout 100 python scripts/time_conv_c1.py 2>&1 | head -3
