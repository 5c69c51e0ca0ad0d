timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmarked.py -q -x -k "rows_ or c1 or C1" 2>&1 | tail -2
timeout 100 python scripts/time_conv_c1.py 2>&1 | head -2
python scripts/time_conv_vjp.py 2>&1 | head -1
EINCONV_ROWS_NO_PDL=1 timeout 100 python scripts/time_conv_c1.py 2>&1 | head -2
