/*
 * nuk_bmas.h -- the BMAS-compatible symbol set exported by libnuk.so.
 *
 * These are the C functions the reference reaches through the ASDF system "bmas"
 * (cl-bmas; .dependencies:2) and selects per element type with c-name
 * (src/utils/translations.lisp:66-121).  The Lisp name bmas:<name> is the C symbol
 * BMAS_<name> (docs/common-lisp-and-numericals.md:233,253); dashes become underscores
 * (bmas:cast-ds -> BMAS_cast_ds).  Prefixes: s = float, d = double, iNN / uNN = signed /
 * unsigned integers.  The C header of BMAS is not in the reference tree; signatures
 * are the ones every call site uses:
 *
 *   binary   (n, x, incx, y, incy, out, incout)  two-arg-fn-non-comparison.lisp:133-134,175
 *   compare  same, out is uint8_t 0/1             two-arg-fn-comparison.lisp:100-102,166-167
 *   unary    (n, x, incx, out, incout)            one-arg-fn-float.lisp:71-74
 *   cast     (n, x, incx, out, incout)            copy-coerce.lisp:130-133
 *   reduce   T f(n, x, incx)                      sum.lisp:64,127; maximum.lisp:65
 *   argred   long f(n, x, incx)                   arg-maximum.lisp:31-73
 *   dot      T f(n, x, incx, y, incy)             linalg/vdot.lisp:23
 *   bitwise  (nbytes, x, incx, [y, incy,] out, incout), BYTE counts/strides
 *                                                 two-arg-fn-logical.lisp:98, lognot.lisp:43-47
 *
 * Replacing libbmas.so by libnuk.so therefore needs no change to the reference's
 * translation tables.
 */
#ifndef NUK_BMAS_H
#define NUK_BMAS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* X-macro lists: (prefix, C type) */
#define NUK_FLOAT_TYPES(X) X(s, float) X(d, double)
#define NUK_SINT_TYPES(X) X(i64, int64_t) X(i32, int32_t) X(i16, int16_t) X(i8, int8_t)
#define NUK_UINT_TYPES(X) X(u64, uint64_t) X(u32, uint32_t) X(u16, uint16_t) X(u8, uint8_t)
#define NUK_ALL_TYPES(X) NUK_FLOAT_TYPES(X) NUK_SINT_TYPES(X) NUK_UINT_TYPES(X)

#define NUK_DECL_BINARY(name, T, TO) \
  void BMAS_##name(const long n, T *x, const long incx, T *y, const long incy, TO *out, const long incout);
#define NUK_DECL_UNARY(name, T, TO) \
  void BMAS_##name(const long n, T *x, const long incx, TO *out, const long incout);
#define NUK_DECL_REDUCE(name, T) T BMAS_##name(const long n, T *x, const long incx);
#define NUK_DECL_ARGRED(name, T) long BMAS_##name(const long n, T *x, const long incx);
#define NUK_DECL_DOT(name, T) T BMAS_##name(const long n, T *x, const long incx, T *y, const long incy);

/* nu:add nu:subtract (floats + signed; unsigned arrays use the signed symbols,
 * src/basic-math/package.lisp:131-136), nu:multiply (all ten), nu:divide (floats) */
#define X(P, T) NUK_DECL_BINARY(P##add, T, T) NUK_DECL_BINARY(P##sub, T, T)
NUK_FLOAT_TYPES(X) NUK_SINT_TYPES(X)
#undef X
#define X(P, T) NUK_DECL_BINARY(P##mul, T, T)
NUK_ALL_TYPES(X)
#undef X
#define X(P, T) NUK_DECL_BINARY(P##div, T, T)
NUK_FLOAT_TYPES(X)
#undef X
/* nu:two-arg-max / nu:two-arg-min, package.lisp:142-147 */
#define X(P, T) NUK_DECL_BINARY(P##max, T, T) NUK_DECL_BINARY(P##min, T, T)
NUK_ALL_TYPES(X)
#undef X
/* nu:two-arg-< <= = /= >= >, package.lisp:182-199 */
#define X(P, T)                                                        \
  NUK_DECL_BINARY(P##lt, T, uint8_t) NUK_DECL_BINARY(P##le, T, uint8_t) \
  NUK_DECL_BINARY(P##eq, T, uint8_t) NUK_DECL_BINARY(P##neq, T, uint8_t) \
  NUK_DECL_BINARY(P##ge, T, uint8_t) NUK_DECL_BINARY(P##gt, T, uint8_t)
NUK_ALL_TYPES(X)
#undef X
/* nu:abs ffloor fceiling fround ftruncate, package.lisp:112-119 */
#define X(P, T)                                                   \
  NUK_DECL_UNARY(P##fabs, T, T) NUK_DECL_UNARY(P##round, T, T)     \
  NUK_DECL_UNARY(P##trunc, T, T) NUK_DECL_UNARY(P##floor, T, T)    \
  NUK_DECL_UNARY(P##ceil, T, T)
NUK_FLOAT_TYPES(X)
#undef X
#define X(P, T) NUK_DECL_UNARY(P##abs, T, T)
NUK_SINT_TYPES(X)
#undef X
/* nu:copy and the casts behind nu:astype / type promotion, package.lisp:121-129 */
#define X(P, T) NUK_DECL_UNARY(P##copy, T, T)
NUK_FLOAT_TYPES(X) NUK_SINT_TYPES(X)
#undef X
NUK_DECL_UNARY(cast_sd, float, double)
NUK_DECL_UNARY(cast_ds, double, float)
#define X(P, T) NUK_DECL_UNARY(cast_##P##s, T, float) NUK_DECL_UNARY(cast_##P##d, T, double)
NUK_SINT_TYPES(X) NUK_UINT_TYPES(X)
#undef X
/* transcendentals, src/transcendental/package.lisp:63-84 */
#define X(P, T)                                                                           \
  NUK_DECL_UNARY(P##sin, T, T) NUK_DECL_UNARY(P##cos, T, T) NUK_DECL_UNARY(P##tan, T, T)   \
  NUK_DECL_UNARY(P##asin, T, T) NUK_DECL_UNARY(P##acos, T, T) NUK_DECL_UNARY(P##atan, T, T) \
  NUK_DECL_UNARY(P##sinh, T, T) NUK_DECL_UNARY(P##cosh, T, T) NUK_DECL_UNARY(P##tanh, T, T) \
  NUK_DECL_UNARY(P##asinh, T, T) NUK_DECL_UNARY(P##acosh, T, T)                            \
  NUK_DECL_UNARY(P##atanh, T, T) NUK_DECL_UNARY(P##exp, T, T) NUK_DECL_UNARY(P##log, T, T) \
  NUK_DECL_UNARY(P##sqrt, T, T) NUK_DECL_BINARY(P##pow, T, T) NUK_DECL_BINARY(P##atan2, T, T)
NUK_FLOAT_TYPES(X)
#undef X
/* nu:sum (floats + signed), nu:maximum / nu:minimum (all ten), package.lisp:149-157 */
#define X(P, T) NUK_DECL_REDUCE(P##sum, T)
NUK_FLOAT_TYPES(X) NUK_SINT_TYPES(X)
#undef X
#define X(P, T) NUK_DECL_REDUCE(P##hmax, T) NUK_DECL_REDUCE(P##hmin, T)
NUK_ALL_TYPES(X)
#undef X
/* nu:arg-maximum / nu:arg-minimum, package.lisp:159-164 (first index of the extremum) */
#define X(P, T) NUK_DECL_ARGRED(P##himax, T) NUK_DECL_ARGRED(P##himin, T)
NUK_ALL_TYPES(X)
#undef X
/* vdot, src/linalg/package.lisp:39-43 */
#define X(P, T) NUK_DECL_DOT(P##dot, T)
NUK_FLOAT_TYPES(X) NUK_SINT_TYPES(X)
#undef X
/* bitwise family: n and strides are in BYTES, package.lisp:166-180 */
#define X(P, T)                                                                      \
  NUK_DECL_BINARY(P##and, uint8_t, uint8_t) NUK_DECL_BINARY(P##or, uint8_t, uint8_t)   \
  NUK_DECL_BINARY(P##xor, uint8_t, uint8_t) NUK_DECL_BINARY(P##andnot, uint8_t, uint8_t) \
  NUK_DECL_UNARY(P##not, uint8_t, uint8_t)
NUK_SINT_TYPES(X) NUK_UINT_TYPES(X)
#undef X
/* fixnum shim helper, src/basic-math/fixnum-c-funs.lisp:17 */
NUK_DECL_BINARY(i64sra, int64_t, int64_t)

#ifdef __cplusplus
}
#endif
#endif /* NUK_BMAS_H */
