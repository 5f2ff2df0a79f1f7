// bmas.cu -- the BMAS_<name> symbols (include/nuk_bmas.h): same names and argument meaning as the
// C library the reference binds (src/basic-math/package.lisp:111-203,
// src/transcendental/package.lisp:62-84, src/linalg/package.lisp:39-43).
//
// Every symbol accepts DEVICE pointers (dense-arrays device storage: the call is one async launch
// on the thread's current stream) or HOST pointers (a pinned Lisp vector, as CFFI passes today):
// host operands are staged through HBM in pipelined chunks -- H2D of chunk k+1, the kernel of
// chunk k and D2H of chunk k-1 overlap on three streams -- and the call returns when OUT is
// complete, which is the reference's synchronous contract.  There is no CPU compute path.
#include <vector>
#include "dispatch.cuh"
#include "staging.cuh"

namespace nuk {

static void host_scalar(const Opnd &o, Scalar64 *v) {
  memset(v, 0, sizeof(*v));
  memcpy(v->bytes, o.ptr, o.size);
}

// Launch = int(const MapProblem&)
template <class Launch>
static int run_map_1d(int arity, long n, Opnd x, Opnd y, Opnd out, Launch launch) {
  ApiRange range("BMAS/map");
  nuk_clear_error();  // nuk_last_status() describes the most recent BMAS_* call of this thread
  const DeviceInfo *dev = device_info();
  if (!dev) return nuk_last_status();
  if (n <= 0) return NUK_OK;
  if (n > 1 && out.inc == 0) {
    set_error(NUK_ERR_INVALID, "OUT stride 0 with n > 1");
    return NUK_ERR_INVALID;
  }
  x.kind = classify_ptr(x.ptr, x.inc != 0);
  y.kind = arity == 2 ? classify_ptr(y.ptr, y.inc != 0) : PK_NULL;
  out.kind = classify_ptr(out.ptr, true);
  const bool any_host_array = (on_host(x.kind) && x.inc != 0) || (on_host(y.kind) && y.inc != 0) ||
                              on_host(out.kind);
  const int64_t dims[1] = {n};

  if (!any_host_array) {
    // device-resident: one launch on the caller's stream
    MapProblem p;
    const int64_t sx[1] = {x.inc}, sy[1] = {y.inc}, so[1] = {out.inc};
    const void *xp = x.ptr, *yp = arity == 2 ? y.ptr : nullptr;
    Scalar64 xv, yv;
    memset(&xv, 0, sizeof(xv)); memset(&yv, 0, sizeof(yv));
    if (on_host(x.kind)) { host_scalar(x, &xv); xp = nullptr; }
    if (arity == 2 && on_host(y.kind)) { host_scalar(y, &yv); yp = nullptr; }
    NUK_TRY(make_problem(&p, arity, x.size, out.size, 1, dims, sx, sy, so, xp, yp,
                         const_cast<void *>(out.ptr), current_stream()));
    p.xv = xv; p.yv = yv;
    return launch(p);
  }

  // ---- staged (staging.cuh): chunk k on slot k mod S
  const size_t wide = x.size > out.size ? x.size : out.size;
  const size_t cap = stage_chunk_bytes((size_t)n * wide);
  const long chunk = (long)(cap / wide);
  Stager &sg = thread_stager();
  NUK_TRY(sg.ensure(cap, (int)opt(OPT_STAGE_SLOTS), dev->device));
  // operands already on the device may still be in flight on the caller's stream
  NUK_CUDA(cudaEventRecord(sg.entry, current_stream()));
  for (int s = 0; s < sg.nslots; ++s) NUK_CUDA(cudaStreamWaitEvent(sg.streams[s], sg.entry, 0));

  int status = NUK_OK;
  long c0 = 0;
  for (int k = 0; c0 < n && status == NUK_OK; ++k, c0 += chunk) {
    const long m = (n - c0 < chunk) ? (n - c0) : chunk;
    const int slot = k % sg.nslots;
    cudaStream_t st = sg.streams[slot];
    status = sg.begin_slot(slot);
    if (status != NUK_OK) break;
    const void *xp = nullptr, *yp = nullptr;
    long xi = 0, yi = 0;
    Scalar64 xv, yv;
    memset(&xv, 0, sizeof(xv)); memset(&yv, 0, sizeof(yv));
    auto prep = [&](const Opnd &o, int which, const void **dp, long *di, Scalar64 *v) -> int {
      if (o.inc == 0) {
        if (on_host(o.kind)) { host_scalar(o, v); *dp = nullptr; }
        else *dp = o.ptr;
        *di = 0;
        return NUK_OK;
      }
      if (on_host(o.kind)) return sg.in(slot, which, o, c0, m, dp, di);
      *dp = static_cast<const char *>(o.ptr) + (int64_t)c0 * o.inc * (int64_t)o.size;
      *di = o.inc;
      return NUK_OK;
    };
    status = prep(x, 0, &xp, &xi, &xv);
    if (status == NUK_OK && arity == 2) status = prep(y, 1, &yp, &yi, &yv);
    if (status != NUK_OK) break;
    void *op;
    long oi;
    if (on_host(out.kind)) {
      sg.out_target(slot, out, m, &op, &oi);
    } else {
      op = static_cast<char *>(const_cast<void *>(out.ptr)) + (int64_t)c0 * out.inc * (int64_t)out.size;
      oi = out.inc;
    }
    MapProblem p;
    const int64_t cd[1] = {m}, sx[1] = {xi}, sy[1] = {yi}, so[1] = {oi};
    status = make_problem(&p, arity, x.size, out.size, 1, cd, sx, sy, so, xp, yp, op, st);
    if (status != NUK_OK) break;
    p.xv = xv; p.yv = yv;
    status = launch(p);
    if (status == NUK_OK && on_host(out.kind)) status = sg.out(slot, out, c0, m);
  }
  const int fin = sg.finish_all();
  return status != NUK_OK ? status : fin;
}

// ------------------------------------------------------------------ reductions by value
static int reduce_value(int red, int dtype, long n, const void *x, long incx, void *result) {
  ApiRange range("BMAS/reduce");
  nuk_clear_error();
  const size_t sz = dtype_size(dtype);
  memset(result, 0, sz);   // a defined value even when the call fails (BMAS returns by value, no error channel)
  const DeviceInfo *dev = device_info();
  if (!dev) return nuk_last_status();
  if (n <= 0) return NUK_OK;
  cudaStream_t st = current_stream();
  void *pin = nullptr, *dres = nullptr;
  NUK_TRY(pinned_slot(&pin));
  const PtrKind kind = classify_ptr(x, incx != 0);
  if (kind == PK_DEVICE) {
    NUK_TRY(thread_buffer(BUF_RESULT, 64, &dres));
    NUK_TRY(reduce_1d_device(red, dtype, (size_t)n, x, incx, dres, st));
  } else {
    // host operand: stage chunk by chunk, one partial per chunk, fixed-order combine on device
    const size_t cap = stage_chunk_bytes((size_t)n * sz);
    const long chunk = (long)(cap / sz);
    const long nchunks = (n + chunk - 1) / chunk;
    Stager &sg = thread_stager();
    NUK_TRY(sg.ensure(cap, (int)opt(OPT_STAGE_SLOTS), dev->device));
    void *dpart = nullptr;
    NUK_TRY(thread_buffer(BUF_PARTS, (size_t)(nchunks + 8) * 8, &dpart));
    Opnd o{x, incx, sz, kind};
    long c0 = 0;
    for (long k = 0; c0 < n; ++k, c0 += chunk) {
      const long m = (n - c0 < chunk) ? (n - c0) : chunk;
      const int slot = (int)(k % sg.nslots);
      const void *dp;
      long di;
      NUK_TRY(sg.begin_slot(slot));
      NUK_TRY(sg.in(slot, 0, o, c0, m, &dp, &di));
      NUK_TRY(reduce_1d_device(red, dtype, (size_t)m, dp, di, static_cast<char *>(dpart) + k * sz,
                               sg.streams[slot], k == 0 ? 0 : NUK_REDUCE_NO_SEED));
    }
    NUK_TRY(sg.finish_all());
    dres = static_cast<char *>(dpart) + nchunks * sz;
    const int red2 = (red == NUK_SUMSQ) ? NUK_SUM : red;
    NUK_TRY(reduce_1d_device(red2, dtype, (size_t)nchunks, dpart, 1, dres, st));
  }
  NUK_CUDA(cudaMemcpyAsync(pin, dres, sz, cudaMemcpyDeviceToHost, st));
  NUK_CUDA(cudaStreamSynchronize(st));
  memcpy(result, pin, sz);
  return NUK_OK;
}

int reduce_partials_device(int red, int dtype, size_t count, const void *partials, void *out_dev,
                           cudaStream_t stream) {
  return reduce_1d_device(red == NUK_SUMSQ ? NUK_SUM : red, dtype, count, partials, 1, out_dev, stream);
}

// dot / arg-reductions live in reduce2.cu
int dot_value(int dtype, long n, const void *x, long incx, const void *y, long incy, void *result);
int argred_value(int is_max, int dtype, long n, const void *x, long incx, long *result);

}  // namespace nuk

using namespace nuk;

// =================================================================== symbol generation
#define DT_s NUK_F32
#define DT_d NUK_F64
#define DT_i64 NUK_I64
#define DT_i32 NUK_I32
#define DT_i16 NUK_I16
#define DT_i8 NUK_I8
#define DT_u64 NUK_U64
#define DT_u32 NUK_U32
#define DT_u16 NUK_U16
#define DT_u8 NUK_U8

#define BMAS_BINARY(P, name, T, TO, LAUNCH)                                                       \
  NUK_API void BMAS_##P##name(const long n, T *x, const long ix, T *y, const long iy, TO *o,      \
                              const long io) {                                                    \
    run_map_1d(2, n, Opnd{x, ix, sizeof(T), PK_NULL}, Opnd{y, iy, sizeof(T), PK_NULL},            \
               Opnd{o, io, sizeof(TO), PK_NULL}, [](const MapProblem &p) { return LAUNCH; });      \
  }
#define BMAS_UNARY_NAMED(sym, T, TO, LAUNCH)                                                      \
  NUK_API void BMAS_##sym(const long n, T *x, const long ix, TO *o, const long io) {              \
    run_map_1d(1, n, Opnd{x, ix, sizeof(T), PK_NULL}, Opnd{nullptr, 0, sizeof(T), PK_NULL},       \
               Opnd{o, io, sizeof(TO), PK_NULL}, [](const MapProblem &p) { return LAUNCH; });      \
  }
#define BMAS_UNARY(P, name, T, TO, LAUNCH) BMAS_UNARY_NAMED(P##name, T, TO, LAUNCH)

// arithmetic
#define X(P, T)                                              \
  BMAS_BINARY(P, add, T, T, launch_arith(NUK_ADD, DT_##P, p)) \
  BMAS_BINARY(P, sub, T, T, launch_arith(NUK_SUB, DT_##P, p))
NUK_FLOAT_TYPES(X) NUK_SINT_TYPES(X)
#undef X
#define X(P, T) BMAS_BINARY(P, mul, T, T, launch_arith(NUK_MUL, DT_##P, p))
NUK_ALL_TYPES(X)
#undef X
#define X(P, T) BMAS_BINARY(P, div, T, T, launch_arith(NUK_DIV, DT_##P, p))
NUK_FLOAT_TYPES(X)
#undef X
#define X(P, T)                                              \
  BMAS_BINARY(P, max, T, T, launch_arith(NUK_MAX, DT_##P, p)) \
  BMAS_BINARY(P, min, T, T, launch_arith(NUK_MIN, DT_##P, p))
NUK_ALL_TYPES(X)
#undef X
// comparisons
#define X(P, T)                                                   \
  BMAS_BINARY(P, lt, T, uint8_t, launch_cmp(NUK_LT, DT_##P, p))    \
  BMAS_BINARY(P, le, T, uint8_t, launch_cmp(NUK_LE, DT_##P, p))    \
  BMAS_BINARY(P, eq, T, uint8_t, launch_cmp(NUK_EQ, DT_##P, p))    \
  BMAS_BINARY(P, neq, T, uint8_t, launch_cmp(NUK_NEQ, DT_##P, p))  \
  BMAS_BINARY(P, ge, T, uint8_t, launch_cmp(NUK_GE, DT_##P, p))    \
  BMAS_BINARY(P, gt, T, uint8_t, launch_cmp(NUK_GT, DT_##P, p))
NUK_ALL_TYPES(X)
#undef X
// abs / rounding
#define X(P, T)                                                  \
  BMAS_UNARY(P, fabs, T, T, launch_unary(NUK_ABS, DT_##P, p))     \
  BMAS_UNARY(P, round, T, T, launch_unary(NUK_ROUND, DT_##P, p))  \
  BMAS_UNARY(P, trunc, T, T, launch_unary(NUK_TRUNC, DT_##P, p))  \
  BMAS_UNARY(P, floor, T, T, launch_unary(NUK_FLOOR, DT_##P, p))  \
  BMAS_UNARY(P, ceil, T, T, launch_unary(NUK_CEIL, DT_##P, p))    \
  BMAS_UNARY(P, sqrt, T, T, launch_unary(NUK_SQRT, DT_##P, p))
NUK_FLOAT_TYPES(X)
#undef X
#define X(P, T) BMAS_UNARY(P, abs, T, T, launch_unary(NUK_ABS, DT_##P, p))
NUK_SINT_TYPES(X)
#undef X
// copy / casts
#define X(P, T) BMAS_UNARY(P, copy, T, T, launch_unary(NUK_COPY, DT_##P, p))
NUK_FLOAT_TYPES(X) NUK_SINT_TYPES(X)
#undef X
BMAS_UNARY_NAMED(cast_sd, float, double, launch_cast(NUK_F32, NUK_F64, p))
BMAS_UNARY_NAMED(cast_ds, double, float, launch_cast(NUK_F64, NUK_F32, p))
#define X(P, T)                                                            \
  BMAS_UNARY_NAMED(cast_##P##s, T, float, launch_cast(DT_##P, NUK_F32, p))  \
  BMAS_UNARY_NAMED(cast_##P##d, T, double, launch_cast(DT_##P, NUK_F64, p))
NUK_SINT_TYPES(X) NUK_UINT_TYPES(X)
#undef X
// transcendentals
#define TR1(P, name, OP, T) BMAS_UNARY(P, name, T, T, launch_op1(OP, DT_##P, p))
#define X(P, T)                                                                       \
  TR1(P, sin, NUK_SIN, T) TR1(P, cos, NUK_COS, T) TR1(P, tan, NUK_TAN, T)              \
  TR1(P, asin, NUK_ASIN, T) TR1(P, acos, NUK_ACOS, T) TR1(P, atan, NUK_ATAN, T)        \
  TR1(P, sinh, NUK_SINH, T) TR1(P, cosh, NUK_COSH, T) TR1(P, tanh, NUK_TANH, T)        \
  TR1(P, asinh, NUK_ASINH, T) TR1(P, acosh, NUK_ACOSH, T) TR1(P, atanh, NUK_ATANH, T)  \
  TR1(P, exp, NUK_EXP, T) TR1(P, log, NUK_LOG, T)                                      \
  BMAS_BINARY(P, pow, T, T, launch_op2(NUK_POW, DT_##P, p))                            \
  BMAS_BINARY(P, atan2, T, T, launch_op2(NUK_ATAN2, DT_##P, p))
NUK_FLOAT_TYPES(X)
#undef X
// bitwise: byte counts, byte strides
#define X(P, T)                                                                        \
  BMAS_BINARY(P, and, uint8_t, uint8_t, launch_bitwise(NUK_AND, NUK_U8, p))             \
  BMAS_BINARY(P, or, uint8_t, uint8_t, launch_bitwise(NUK_OR, NUK_U8, p))               \
  BMAS_BINARY(P, xor, uint8_t, uint8_t, launch_bitwise(NUK_XOR, NUK_U8, p))             \
  BMAS_BINARY(P, andnot, uint8_t, uint8_t, launch_bitwise(NUK_ANDNOT, NUK_U8, p))       \
  BMAS_UNARY(P, not, uint8_t, uint8_t, launch_unary(NUK_NOT, NUK_U8, p))
NUK_SINT_TYPES(X) NUK_UINT_TYPES(X)
#undef X
BMAS_BINARY(i64, sra, int64_t, int64_t, launch_bitwise(NUK_SRA, NUK_I64, p))

// reductions returned by value
#define BMAS_REDUCE(P, name, T, RED)                                   \
  NUK_API T BMAS_##P##name(const long n, T *x, const long ix) {        \
    T r;                                                               \
    reduce_value(RED, DT_##P, n, x, ix, &r);                           \
    return r;                                                          \
  }
#define X(P, T) BMAS_REDUCE(P, sum, T, NUK_SUM)
NUK_FLOAT_TYPES(X) NUK_SINT_TYPES(X)
#undef X
#define X(P, T) BMAS_REDUCE(P, hmax, T, NUK_HMAX) BMAS_REDUCE(P, hmin, T, NUK_HMIN)
NUK_ALL_TYPES(X)
#undef X
#define X(P, T)                                                          \
  NUK_API long BMAS_##P##himax(const long n, T *x, const long ix) {      \
    long r = 0;                                                          \
    argred_value(1, DT_##P, n, x, ix, &r);                               \
    return r;                                                            \
  }                                                                      \
  NUK_API long BMAS_##P##himin(const long n, T *x, const long ix) {      \
    long r = 0;                                                          \
    argred_value(0, DT_##P, n, x, ix, &r);                               \
    return r;                                                            \
  }
NUK_ALL_TYPES(X)
#undef X
#define X(P, T)                                                                             \
  NUK_API T BMAS_##P##dot(const long n, T *x, const long ix, T *y, const long iy) {         \
    T r;                                                                                    \
    dot_value(DT_##P, n, x, ix, y, iy, &r);                                                 \
    return r;                                                                               \
  }
NUK_FLOAT_TYPES(X) NUK_SINT_TYPES(X)
#undef X
