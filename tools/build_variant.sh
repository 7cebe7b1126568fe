#!/usr/bin/env bash
# Builds an A/B variant of the library into scratch/: tools/build_variant.sh NAME -DFOO=1 ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p scratch
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC \
  -Xcompiler -fvisibility=default -I include "$@" -o scratch/libsia_$name.so skin_image_analysis_b200/csrc/libsia_unity.cu
echo built scratch/libsia_$name.so
