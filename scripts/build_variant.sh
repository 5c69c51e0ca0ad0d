#!/bin/bash
# usage: scripts/build_variant.sh NAME "-DEINCONV_FLAT_THREADS=512 ..."  -> einconv_b200/_lib/variants/libeinconv_NAME.so
# Rebuilds csrc/unfold.cu with extra defines and links it with the objects of the regular build
# (run `python -m einconv_b200.build` first).  Select at run time with EINCONV_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
mkdir -p einconv_b200/_lib/variants
src=${3:-unfold}
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
  $2 -c einconv_b200/csrc/$src.cu -o einconv_b200/_lib/variants/${src}_$1.o
objs=$(ls einconv_b200/_lib/obj/*.o | grep -v "/$src.o")
nvcc -shared -o einconv_b200/_lib/variants/libeinconv_$1.so $objs einconv_b200/_lib/variants/${src}_$1.o -lcudart_static
echo einconv_b200/_lib/variants/libeinconv_$1.so
