timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmarked.py -q -k "rows_ or c1 or C1" 2>&1 | tail -2
timeout 100 python scripts/time_conv_c1.py 2>&1 | head -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/one_conv_c1.py 2>/dev/null | grep -i "pack_rows" | tail -2 | awk -F'","' '{print $5, $NF}'
