#!/bin/bash
# Round-2 evidence pass (one GPU): ncu --set full of the dominant kernel of C1-C4 + launch list of the default bench.
# Run AFTER the plain commands exited 0.  Outputs under gpurun_out/; summarise with scripts/ncu_summary.py.
set -x
for w in unfold conv_fwd depthwise weight_vjp; do
  python scripts/one_call.py $w || exit 1
done
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:unfold_flat -c 1 -s 1 -o gpurun_out/r2f_unfold python scripts/one_call.py unfold > gpurun_out/r2f_ncu_unfold.log 2>&1
$NCU -k regex:conv_rows_kernel -c 1 -s 1 -o gpurun_out/r2f_conv_rows python scripts/one_call.py conv_fwd > gpurun_out/r2f_ncu_rows.log 2>&1
$NCU -k regex:dw_vec4 -c 2 -s 2 -o gpurun_out/r2f_depthwise python scripts/one_call.py depthwise > gpurun_out/r2f_ncu_dw.log 2>&1
$NCU -k regex:"wgrad_slab_kernel|pack_vec4|wgrad_finish" -c 4 -s 4 -o gpurun_out/r2f_wgrad python scripts/one_call.py weight_vjp > gpurun_out/r2f_ncu_wgrad.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1
ls -la gpurun_out/r2f_*
