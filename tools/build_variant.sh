#!/bin/bash
# tools/build_variant.sh <tag> <file.cu> [-DFLAG=...]... -- an A/B build of libnuk.so that differs from the current
# one in ONE translation unit compiled with extra -D flags: numericals_b200/lib/variants/libnuk_<tag>.so
# (select it with NUK_LIB=... ; tools/sweep.py and friends honour it).
set -e
tag=$1; src=$2; shift 2
cd "$(dirname "$0")/../numericals_b200/csrc"
mkdir -p ../lib/variants
base=$(basename $src .cu)
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
  -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -cudart static "$@" -c $src -o ../lib/variants/${base}_$tag.o
objs=$(ls ../lib/obj/*.o | grep -v "/${base}.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../lib/variants/libnuk_$tag.so $objs ../lib/variants/${base}_$tag.o
rm -f ../lib/variants/${base}_$tag.o
echo built numericals_b200/lib/variants/libnuk_$tag.so
