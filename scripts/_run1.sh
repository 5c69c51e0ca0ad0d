ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5_l32.csv python scripts/one_sweep.py 8 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5_l256.csv python scripts/one_sweep.py 1 1 > /dev/null 2>&1
ls -la gpurun_out/r5_l*.csv
