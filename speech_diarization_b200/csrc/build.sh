#!/bin/bash
# Builds libsd_b200.so (sm_100a only) next to this script.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${SD_EXTRA_FLAGS} -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall"
SRCS=$(ls *.cu)
OBJS=""
for f in $SRCS; do
  o="build/${f%.cu}.o"
  mkdir -p build
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o" 2>/dev/null)" ] || [ ../../include/sd_b200.h -nt "$o" ]; then
    echo "nvcc $f"
    $NVCC $FLAGS ${SD_PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
  fi
  OBJS="$OBJS $o"
done
wait
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o libsd_b200.so $OBJS -lcudart_static -ldl -lrt -lpthread
echo "built $(pwd)/libsd_b200.so"
