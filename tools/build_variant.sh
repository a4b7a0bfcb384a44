#!/bin/bash
# Build a variant of libtmc_b200.so in which ONE source file is compiled with extra flags (A/B experiments):
#   tools/build_variant.sh <name> <source.cu> <extra nvcc flags...>   ->  torch_motion_correction_b200/libtmc_b200_<name>.so
# Select it at run time with TMC_B200_LIB=<path>.  The other objects come from the regular build (build/*.o).
set -e
name=$1; src=$2; shift 2
pkg=$(dirname "$0")/../torch_motion_correction_b200
base=$(basename "$src" .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DTMC_B200=1 "$@" \
     -c "$pkg/csrc/$base.cu" -o "$pkg/build/${base}_$name.o"
objs=""
for f in "$pkg"/csrc/*.cu; do b=$(basename "$f" .cu); [ "$b" != "$base" ] && objs="$objs $pkg/build/$b.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$pkg/libtmc_b200_$name.so" $objs "$pkg/build/${base}_$name.o"
echo "$pkg/libtmc_b200_$name.so"
