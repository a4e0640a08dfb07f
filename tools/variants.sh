#!/bin/bash
# Builds tuning variants of the library: tools/variants.sh name1 "-DFLAG1=.." name2 "-DFLAG2=.." ...
# -> build/variants/lib_<name>.so (fast build: K=20 only).  Run one with VLG_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DVLG_FAST_BUILD $flags \
      -o build/variants/lib_$name.so video-layout-generation_b200/csrc/vlg_api.cu && echo "built $name" ) &
done
wait
