#!/bin/bash
# tools/ulp_report.sh -- regenerate profiles/ulp_report_r0N.jsonl (pass the path): EXHAUSTIVE (2^32 inputs) sweeps of
# every single-float unary function, 4e7-sample sweeps of every double function, 4e8 / 1e8 samples of
# the binary ones, against glibc (double / long double) and SLEEF u10.  ~40 min on 8 cores.
set -e
cd "$(dirname "$0")/.."
python -c "
import sys; sys.path.insert(0,'tests')
import _helpers as h; print(h.build_host_tool('ulp_sweep.cpp','ulp_sweep'))"
S=tests/ulp/_build/ulp_sweep
out=${1:-profiles/ulp_report_r01.jsonl}
: > $out.tmp
for f in exp log sin cos tanh tan asin acos atan sinh cosh asinh acosh atanh; do $S $f f32 | tail -1 >> $out.tmp; done
for f in exp log sin cos tan asin acos atan sinh cosh tanh asinh acosh atanh; do $S $f f64 40000000 | tail -1 >> $out.tmp; done
$S pow f64 40000000 | tail -1 >> $out.tmp
$S atan2 f64 40000000 | tail -1 >> $out.tmp
$S pow f32b 400000000 | tail -1 >> $out.tmp
$S atan2 f32b 100000000 | tail -1 >> $out.tmp
mv $out.tmp $out
