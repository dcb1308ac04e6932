#!/bin/bash
# Builds libwap_b200.so (sm_100a) in-tree.  Usage: csrc/build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=../libwap_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p build
pids=()
for f in *.cu; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o")" ] || [ ../../include/wap_abi.h -nt "$o" ]; then
    nvcc $FLAGS "$@" -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/*.o
echo "built $(realpath $OUT)"
