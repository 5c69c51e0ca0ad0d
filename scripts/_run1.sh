timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "rows_forward" > gpurun_out/r3_rows_t.log 2>&1
tail -4 gpurun_out/r3_rows_t.log
grep -n "AssertionError: rows" gpurun_out/r3_rows_t.log | head
timeout 200 python scripts/time_conv_c1.py > gpurun_out/r3_c1_time.log 2>&1
head -3 gpurun_out/r3_c1_time.log
ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg --clock-control none --csv python scripts/one_conv_c1.py 2>/dev/null | grep -v "^==" | awk -F'","' '{print $5, $(NF-2), $(NF)}' | tail -4
