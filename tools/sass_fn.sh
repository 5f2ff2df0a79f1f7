#!/bin/bash
# tools/sass_fn.sh <function> [f32|f64] [unary|binary] -- compile ONE call of a nukmath function
# into a trivial kernel and print its SASS opcode histogram grouped by pipe (no GPU needed).
fn=$1; ty=${2:-f32}; ar=${3:-unary}
T=float; [ "$ty" = f64 ] && T=double
d=$(dirname "$0")/_build; mkdir -p $d
if [ "$ar" = binary ]; then body="o[i] = nukmath::$fn(x[i], y[i]);"; else body="o[i] = nukmath::$fn(x[i]);"; fi
cat > $d/one.cu <<EOT
#include "../../numericals_b200/csrc/nukmath_f64.cuh"
__global__ void k(const $T *x, const $T *y, $T *o) { int i = blockIdx.x * blockDim.x + threadIdx.x; $body }
EOT
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false --expt-relaxed-constexpr -cubin -o $d/one.cubin $d/one.cu || exit 1
cuobjdump -sass $d/one.cubin | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/^\s*\/\*[0-9a-f]*\*\/\s*//' | awk '{op=$1; if (op ~ /^@/) op=$2; sub(/[.;].*/,"",op); c[op]++; t++}
END { dp=c["DFMA"]+c["DADD"]+c["DMUL"]+c["DSETP"]; xu=c["MUFU"]+c["F2F"]+c["I2F"]+c["F2I"]+c["FRND"];
  printf "total %d  DP %d  XU %d  other %d |", t, dp, xu, t-dp-xu; for (k in c) printf " %s:%d", k, c[k]; printf "\n" }'
