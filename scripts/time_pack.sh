#!/bin/bash
# C4 weight VJP and C1 forward with different numbers of rows per CTA in the pack kernel
for rg in 1 2 4 8 16; do
  r=$(EINCONV_PACK_RG=$rg python bench.py --workload weight_vjp --no-suite --no-cpu-baseline --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1))")
  f=$(EINCONV_PACK_RG=$rg python bench.py --workload conv_fwd --no-suite --no-cpu-baseline --steps 20 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1))")
  echo "rows_per_cta=$rg: C4 $r us, C1 $f us"
done
