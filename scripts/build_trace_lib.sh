#!/bin/bash
# debug build of the library with the message-kernel pipeline trace compiled in (scripts/t5_trace.py)
set -e
cd "$(dirname "$0")/../adsorbdiff_b200/csrc"
mkdir -p /tmp/adk_trace
for f in api neighbors node_ops linear linear_tc message message_t5 message_mma message_bwd train_ops se3_step; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include -I. -DT5_TRACE -c $f.cu -o /tmp/adk_trace/$f.o 2>/dev/null &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../lib/libadsorbdiff_b200_trace.so /tmp/adk_trace/*.o -lcudart
