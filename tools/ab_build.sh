#!/bin/bash
# Build an alternative library for same-box A/B timing: tools/ab_build.sh <tag> <extra nvcc flags...>
# -> build/ab/libvocalie_b200_<tag>.so, loaded with VT_LIB_PATH=build/ab/libvocalie_b200_<tag>.so
set -e
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p build/ab/$TAG
for f in vocalie-tts_b200/csrc/vt_*.cu; do
  o=build/ab/$TAG/$(basename $f .cu).o
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $o &
done
wait
nvcc -shared -o build/ab/libvocalie_b200_$TAG.so build/ab/$TAG/*.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo built build/ab/libvocalie_b200_$TAG.so
