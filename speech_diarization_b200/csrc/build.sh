#!/bin/bash
# Builds libsd_b200.so (sm_100a only) next to this script.
#   SD_EXTRA_FLAGS  extra nvcc flags (e.g. -DSD_EXPERIMENTS=1)
#   SD_BUILD_DIR    object directory (default build/); SD_OUT  output library (default libsd_b200.so)
#   SD_PTXAS_V=1    print ptxas register / spill statistics
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
BUILD=${SD_BUILD_DIR:-build}
OUT=${SD_OUT:-libsd_b200.so}
FLAGS="${SD_EXTRA_FLAGS} -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall"
mkdir -p "$BUILD"
# a change of flags rebuilds everything
if [ "$(cat "$BUILD/.flags" 2>/dev/null)" != "$FLAGS" ]; then rm -f "$BUILD"/*.o; echo "$FLAGS" > "$BUILD/.flags"; fi
OBJS=""
PIDS=""
for f in *.cu; do
  o="$BUILD/${f%.cu}.o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o" 2>/dev/null)" ] || [ ../../include/sd_b200.h -nt "$o" ]; then
    echo "nvcc $f"
    rm -f "$o" "$OUT"     # a failed compile must leave neither a stale object nor a stale library behind
    $NVCC $FLAGS ${SD_PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS $o"
done
for p in $PIDS; do wait "$p" || { echo "build.sh: a compile job failed" >&2; exit 1; }; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" $OBJS -lcudart_static -ldl -lrt -lpthread
echo "built $(cd "$(dirname "$OUT")" && pwd)/$(basename "$OUT")"
