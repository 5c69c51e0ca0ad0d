timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchmarked.py -q -x -n 2 -k "kfc or factor or Factor or sweep or wgrad or weight_vjp" 2>&1 | tail -3
python scripts/one_sweep.py 1 10
python scripts/one_sweep.py 8 10
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:kfc_finish_block -c 1 -s 1 -o gpurun_out/r2h_finish_block python scripts/one_layer_b32.py > gpurun_out/r2h_ncu_fin.log 2>&1
$NCU -k regex:flat_copy_kernel -c 1 -s 3 -o gpurun_out/r2h_flat_copy python scripts/one_layer_b32.py 256 > gpurun_out/r2h_ncu_flat.log 2>&1
$NCU -k regex:plane_sum_kernel -c 1 -s 1 -o gpurun_out/r2h_plane_sum python scripts/one_kfacr.py > gpurun_out/r2h_ncu_ps.log 2>&1
ls gpurun_out/r2h_*
