#!/bin/bash
# bring-up experiment builds of the library: message_t5.cu compiled with -DT5_EXP=n (see the kernel source)
set -e
cd "$(dirname "$0")/../adsorbdiff_b200/csrc"
for n in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include -I. -DT5_EXP=$n -c message_t5.cu -o /tmp/t5_exp$n.o 2>/dev/null
  objs=$(ls ../lib/*.o | grep -v message_t5.o)
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../lib/libadsorbdiff_b200_exp$n.so $objs /tmp/t5_exp$n.o -lcudart
done
