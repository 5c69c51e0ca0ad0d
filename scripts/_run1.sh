cp einconv_b200/_lib/variants/lib_trace.so einconv_b200/_lib/libeinconv_b200.so
EINCONV_ROWS_TRACE=1 python scripts/one_conv_c1.py 2>&1 | tail -10
