cp einconv_b200/_lib/variants/libeinconv_trace.so einconv_b200/_lib/libeinconv_b200.so
EINCONV_ROWS_SKIP_PACK=1 timeout 100 python scripts/time_rows_fixed.py 2>&1 | tail -8
EINCONV_ROWS_PAIR=1 EINCONV_ROWS_SKIP_PACK=1 timeout 100 python scripts/time_rows_fixed.py 2>&1 | tail -8
