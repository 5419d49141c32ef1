#!/bin/bash
# Build libgsx variants with different -D tuning macros into 3dgsconverter_b200/lib/variants/ (dev aid).
# usage: build_variants.sh name1:"-DX=1 -DY=2" name2:"..."
set -e
cd "$(dirname "$0")/../3dgsconverter_b200/csrc"
mkdir -p ../lib/variants ../build_var
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  objs=""
  for f in gsx_abi gsx_sor gsx_stats gsx_masks gsx_density gsx_kmeans gsx_knn_exact gsx_radix gsx_compact; do
    if [ "$f" = "gsx_sor" ] || [ "$f" = "gsx_kmeans" ]; then
      /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -prec-div=true -prec-sqrt=true -std=c++17 -Xcompiler -fPIC,-O2 -ccbin /usr/bin/g++ $flags -c -o ../build_var/${f}_${name}.o $f.cu &
      objs="$objs ../build_var/${f}_${name}.o"
    else
      objs="$objs ../build/$f.o"
    fi
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/libgsx_${name}.so $objs -ccbin /usr/bin/g++
  echo built $name
done
