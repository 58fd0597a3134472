#!/bin/bash
# ablation builds of the library: message_t5.cu compiled with -DT5_EXP=n (1..6, see the kernel source); time with
#   ADK_LIB=adsorbdiff_b200/lib/libadsorbdiff_b200_exp$n.so T5_TIME_ONLY=1 python scripts/t5_check.py 1024
set -e
cd "$(dirname "$0")/../adsorbdiff_b200/csrc"
for n in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include -I. -DT5_EXP=$n -c message_t5.cu -o /tmp/t5_exp$n.o 2>/dev/null
  objs=$(ls ../lib/*.o | grep -v message_t5.o)
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../lib/libadsorbdiff_b200_exp$n.so $objs /tmp/t5_exp$n.o -lcudart
done
