#!/bin/bash
# Build A/B variants of the library with different -D knobs into build/variants/<name>.so (dev tool; the product build
# is __graft_entry__.build()).  usage: tools/build_variants.sh name1="-DX=1 -DY=2" name2="..."
mkdir -p build/variants
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
      -o build/variants/$name.so stgcn_b200/csrc/stgcn_b200.cu && echo built $name ) &
done
wait
