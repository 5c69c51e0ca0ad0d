#!/bin/bash
# times the C2 unfold with every library variant under einconv_b200/_lib/variants
for lib in einconv_b200/_lib/libeinconv_b200.so einconv_b200/_lib/variants/libeinconv_*.so; do
  for c in 1 2; do
    r=$(EINCONV_B200_LIB=$PWD/$lib EINCONV_UNFOLD_CHUNKS=$c python bench.py --no-suite --no-cpu-baseline --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['roofline']['frac'],3))")
    echo "$(basename $lib) chunks=$c: $r"
  done
done
