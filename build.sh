#!/bin/bash
# Builds the C-ABI shared library for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=deepsee_b200/lib/libdeepsee_b200.so
mkdir -p deepsee_b200/lib build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default ${DSEE_NVCC_EXTRA}"
pids=()
for f in deepsee_b200/csrc/*.cu; do
  o=build/$(basename ${f%.cu}).o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ deepsee_b200/csrc/common.cuh -nt "$o" ] || [ include/deepsee_b200.h -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT build/*.o -lcudart
echo "built $OUT"
