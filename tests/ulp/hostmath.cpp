// tests/ulp/hostmath.cpp -- TEST INFRASTRUCTURE: the product math headers compiled for the HOST
// (g++ -mfma -ffp-contract=off), exported as array functions.  tests compare the GPU results with
// these BIT FOR BIT: that is what lets the exhaustive CPU ULP sweeps (ulp_sweep.cpp) speak for
// the device code.
#include "../../numericals_b200/csrc/nukmath_f32.cuh"
#include "../../numericals_b200/csrc/nukmath_f64.cuh"
using namespace nukmath;
#define U32(name) extern "C" void hm_s##name(long n, const float *x, float *o) { for (long i = 0; i < n; ++i) o[i] = name##_f32(x[i]); }
#define U64(name) extern "C" void hm_d##name(long n, const double *x, double *o) { for (long i = 0; i < n; ++i) o[i] = name##_f64(x[i]); }
#define B32(name) extern "C" void hm_s##name(long n, const float *x, const float *y, float *o) { for (long i = 0; i < n; ++i) o[i] = name##_f32(x[i], y[i]); }
#define B64(name) extern "C" void hm_d##name(long n, const double *x, const double *y, double *o) { for (long i = 0; i < n; ++i) o[i] = name##_f64(x[i], y[i]); }
#define ALL1(X) X(sin) X(cos) X(tan) X(asin) X(acos) X(atan) X(sinh) X(cosh) X(tanh) X(asinh) X(acosh) X(atanh) X(exp) X(log)
ALL1(U32) ALL1(U64)
B32(pow) B32(atan2) B64(pow) B64(atan2)
