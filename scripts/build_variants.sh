#!/bin/bash
# Build tuning variants of libolb (block size / min blocks per SM) into build/variants/.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for cfg in "128 1" "128 4" "256 1" "256 2" "256 3" "512 1"; do
  set -- $cfg
  out=build/variants/libolb_b$1_m$2.so
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC \
    -DOLB_BLOCK=$1 -DOLB_MIN_BLOCKS=$2 -o $out optiland_b200/csrc/olb_trace.cu &
done
wait
ls -la build/variants
