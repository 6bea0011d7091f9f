#!/bin/bash
# Builds alternative libdmhomo.so variants under tools/_var/<name>/ for A/B runs on the GPU box:
#   tools/build_variants.sh name1 "-DFOO=1" name2 "-DBAR=2" ...     then   DMH_LIB=tools/_var/name1/libdmhomo.so python bench.py
set -e
cd "$(dirname "$0")/.."
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  d=tools/_var/$name; mkdir -p $d
  for f in dmh_api dmh_warp dmh_warp_fast dmh_dlt dmh_flow dmh_persp dmh_next; do
    cp dmhomo_b200/csrc/$f.o $d/$f.o      # unchanged objects of the main build
  done
  nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -cudart static $flags \
    -Xptxas -v -c -o $d/dmh_warp_tile.o dmhomo_b200/csrc/dmh_warp_tile.cu 2>&1 | grep -A2 "ILi6ELi1ELb1E" | grep -E "spill|registers" | tr '\n' ' '
  echo " <- $name ($flags)"
  nvcc $ARCH -shared -cudart static -o $d/libdmhomo.so $d/*.o
done
