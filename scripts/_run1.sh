python scripts/one_sweep.py 1 10
python scripts/one_sweep.py 8 10
EINCONV_KFC_FINISH_BC=16 python scripts/one_sweep.py 1 10
EINCONV_KFC_FINISH_BC=16 python scripts/one_sweep.py 8 10
ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/one_layer_b32.py 2>/dev/null | grep -i "finish" | awk -F'","' '{print $5, $8, $NF}' | head -14
ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/one_layer_b32.py 256 2>/dev/null | grep -i "finish" | awk -F'","' '{print $5, $8, $NF}' | head -14
