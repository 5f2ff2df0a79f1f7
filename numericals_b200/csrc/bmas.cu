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

namespace nuk {

enum PtrKind { PK_NULL = 0, PK_DEVICE = 1, PK_HOST = 2 };
static PtrKind classify(const void *p) {
  if (!p) return PK_NULL;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return PK_HOST;
  }
  return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? PK_DEVICE : PK_HOST;
}

// ------------------------------------------------------------------ staging pipeline
struct Stager {
  static constexpr int kMaxSlots = 4;
  int nslots = 0;
  size_t cap = 0;                       // bytes per operand buffer
  cudaStream_t streams[kMaxSlots] = {};
  void *buf[kMaxSlots][3] = {};         // [slot][x, y, out]
  cudaEvent_t entry = nullptr;
  int device = -1;
  ~Stager() { release(); }
  void release() {
    for (int s = 0; s < nslots; ++s) {
      for (int o = 0; o < 3; ++o)
        if (buf[s][o]) cudaFree(buf[s][o]);
      if (streams[s]) cudaStreamDestroy(streams[s]);
    }
    if (entry) cudaEventDestroy(entry);
    nslots = 0; cap = 0; entry = nullptr;
    memset(buf, 0, sizeof(buf));
    memset(streams, 0, sizeof(streams));
  }
  int ensure(size_t want_cap, int want_slots, int dev) {
    if (want_slots > kMaxSlots) want_slots = kMaxSlots;
    if (want_slots < 1) want_slots = 1;
    if (nslots == want_slots && cap >= want_cap && device == dev) return NUK_OK;
    release();
    device = dev;
    for (int s = 0; s < want_slots; ++s) {
      NUK_CUDA(cudaStreamCreateWithFlags(&streams[s], cudaStreamNonBlocking));
      for (int o = 0; o < 3; ++o) NUK_CUDA(cudaMalloc(&buf[s][o], want_cap));
      nslots = s + 1;
    }
    NUK_CUDA(cudaEventCreateWithFlags(&entry, cudaEventDisableTiming));
    cap = want_cap;
    return NUK_OK;
  }
};
static thread_local Stager tl_stager;

// one operand of a BMAS-shaped (rank-1) call
struct Opnd {
  const void *ptr;
  long inc;
  size_t size;     // element size
  PtrKind kind;
};

// copy `m` elements starting at logical index c0 of a HOST strided operand into a contiguous
// device buffer (in ADDRESS order); returns the device pointer/inc that walk it in logical order
static int stage_in(const Opnd &o, long c0, long m, void *dbuf, cudaStream_t st, const void **dptr,
                    long *dinc) {
  const char *base = static_cast<const char *>(o.ptr);
  if (o.inc == 1) {
    NUK_CUDA(cudaMemcpyAsync(dbuf, base + (size_t)c0 * o.size, (size_t)m * o.size,
                             cudaMemcpyHostToDevice, st));
    *dptr = dbuf; *dinc = 1;
  } else if (o.inc == -1) {
    const char *lo = base + (int64_t)(c0 + m - 1) * o.inc * (int64_t)o.size;
    NUK_CUDA(cudaMemcpyAsync(dbuf, lo, (size_t)m * o.size, cudaMemcpyHostToDevice, st));
    *dptr = static_cast<char *>(dbuf) + (size_t)(m - 1) * o.size; *dinc = -1;
  } else {
    const long ainc = o.inc < 0 ? -o.inc : o.inc;
    const char *lo = o.inc > 0 ? base + (int64_t)c0 * o.inc * (int64_t)o.size
                               : base + (int64_t)(c0 + m - 1) * o.inc * (int64_t)o.size;
    NUK_CUDA(cudaMemcpy2DAsync(dbuf, o.size, lo, (size_t)ainc * o.size, o.size, (size_t)m,
                               cudaMemcpyHostToDevice, st));
    if (o.inc > 0) { *dptr = dbuf; *dinc = 1; }
    else { *dptr = static_cast<char *>(dbuf) + (size_t)(m - 1) * o.size; *dinc = -1; }
  }
  return NUK_OK;
}
static int stage_out(const Opnd &o, long c0, long m, const void *dbuf, cudaStream_t st) {
  char *base = static_cast<char *>(const_cast<void *>(o.ptr));
  if (o.inc == 1) {
    NUK_CUDA(cudaMemcpyAsync(base + (size_t)c0 * o.size, dbuf, (size_t)m * o.size,
                             cudaMemcpyDeviceToHost, st));
  } else if (o.inc == -1) {
    char *lo = base + (int64_t)(c0 + m - 1) * o.inc * (int64_t)o.size;
    NUK_CUDA(cudaMemcpyAsync(lo, dbuf, (size_t)m * o.size, cudaMemcpyDeviceToHost, st));
  } else {
    const long ainc = o.inc < 0 ? -o.inc : o.inc;
    char *lo = o.inc > 0 ? base + (int64_t)c0 * o.inc * (int64_t)o.size
                         : base + (int64_t)(c0 + m - 1) * o.inc * (int64_t)o.size;
    // only the view's elements are written (test-definition-macros.lisp:86-99)
    NUK_CUDA(cudaMemcpy2DAsync(lo, (size_t)ainc * o.size, dbuf, o.size, o.size, (size_t)m,
                               cudaMemcpyDeviceToHost, st));
  }
  return NUK_OK;
}

static void host_scalar(const Opnd &o, Scalar64 *v) {
  memset(v, 0, sizeof(*v));
  memcpy(v->bytes, o.ptr, o.size);
}

// Launch = int(const MapProblem&)
template <class Launch>
static int run_map_1d(int arity, long n, Opnd x, Opnd y, Opnd out, Launch launch) {
  nuk_clear_error();  // nuk_last_status() describes the most recent BMAS_* call of this thread
  const DeviceInfo *dev = device_info();
  if (!dev) return nuk_last_status();
  if (n <= 0) return NUK_OK;
  if (n > 1 && out.inc == 0) {
    set_error(NUK_ERR_INVALID, "OUT stride 0 with n > 1");
    return NUK_ERR_INVALID;
  }
  x.kind = classify(x.ptr);
  y.kind = arity == 2 ? classify(y.ptr) : PK_NULL;
  out.kind = classify(out.ptr);
  const bool any_host_array = (x.kind == PK_HOST && x.inc != 0) || (y.kind == PK_HOST && y.inc != 0) ||
                              out.kind == PK_HOST;
  const int64_t dims[1] = {n};

  if (!any_host_array) {
    // device-resident: one launch on the caller's stream
    MapProblem p;
    const int64_t sx[1] = {x.inc}, sy[1] = {y.inc}, so[1] = {out.inc};
    const void *xp = x.ptr, *yp = arity == 2 ? y.ptr : nullptr;
    Scalar64 xv, yv;
    memset(&xv, 0, sizeof(xv)); memset(&yv, 0, sizeof(yv));
    if (x.kind == PK_HOST) { host_scalar(x, &xv); xp = nullptr; }
    if (arity == 2 && y.kind == PK_HOST) { host_scalar(y, &yv); yp = nullptr; }
    NUK_TRY(make_problem(&p, arity, x.size, out.size, 1, dims, sx, sy, so, xp, yp,
                         const_cast<void *>(out.ptr), current_stream()));
    p.xv = xv; p.yv = yv;
    return launch(p);
  }

  // ---- staged
  const size_t wide = x.size > out.size ? x.size : out.size;
  size_t cap = (size_t)(option("stage_mb") > 0 ? option("stage_mb") : 32) << 20;
  if ((size_t)n * wide < cap) cap = (((size_t)n * wide + 255) / 256) * 256;
  const long chunk = (long)(cap / wide);
  Stager &sg = tl_stager;
  NUK_TRY(sg.ensure(cap, (int)option("stage_slots"), dev->device));
  // operands already on the device may still be in flight on the caller's stream
  NUK_CUDA(cudaEventRecord(sg.entry, current_stream()));
  for (int s = 0; s < sg.nslots; ++s) NUK_CUDA(cudaStreamWaitEvent(sg.streams[s], sg.entry, 0));

  int status = NUK_OK;
  long c0 = 0;
  for (int k = 0; c0 < n && status == NUK_OK; ++k, c0 += chunk) {
    const long m = (n - c0 < chunk) ? (n - c0) : chunk;
    const int slot = k % sg.nslots;
    cudaStream_t st = sg.streams[slot];
    const void *xp = nullptr, *yp = nullptr;
    long xi = 0, yi = 0;
    Scalar64 xv, yv;
    memset(&xv, 0, sizeof(xv)); memset(&yv, 0, sizeof(yv));
    auto prep = [&](const Opnd &o, int which, const void **dp, long *di, Scalar64 *v) -> int {
      if (o.inc == 0) {
        if (o.kind == PK_HOST) { host_scalar(o, v); *dp = nullptr; }
        else *dp = o.ptr;
        *di = 0;
        return NUK_OK;
      }
      if (o.kind == PK_HOST) return stage_in(o, c0, m, sg.buf[slot][which], st, dp, di);
      *dp = static_cast<const char *>(o.ptr) + (int64_t)c0 * o.inc * (int64_t)o.size;
      *di = o.inc;
      return NUK_OK;
    };
    status = prep(x, 0, &xp, &xi, &xv);
    if (status == NUK_OK && arity == 2) status = prep(y, 1, &yp, &yi, &yv);
    if (status != NUK_OK) break;
    void *op;
    long oi;
    if (out.kind == PK_HOST) {
      // staged OUT is contiguous in address order; walk it in the logical direction
      op = out.inc > 0 ? sg.buf[slot][2] : static_cast<char *>(sg.buf[slot][2]) + (size_t)(m - 1) * out.size;
      oi = out.inc > 0 ? 1 : -1;
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
    if (status == NUK_OK && out.kind == PK_HOST) status = stage_out(out, c0, m, sg.buf[slot][2], st);
  }
  for (int s = 0; s < sg.nslots; ++s) {
    cudaError_t e = cudaStreamSynchronize(sg.streams[s]);
    if (e != cudaSuccess && status == NUK_OK) status = cuda_fail(e, "staging stream sync", __FILE__, __LINE__);
  }
  return status;
}

// ------------------------------------------------------------------ reductions by value
static int reduce_value(int red, int dtype, long n, const void *x, long incx, void *result) {
  nuk_clear_error();
  const DeviceInfo *dev = device_info();
  if (!dev) return nuk_last_status();
  const size_t sz = dtype_size(dtype);
  memset(result, 0, sz);
  if (n <= 0) return NUK_OK;
  cudaStream_t st = current_stream();
  void *pin = nullptr, *dres = nullptr;
  NUK_TRY(pinned_slot(&pin));
  const PtrKind kind = classify(x);
  if (kind == PK_DEVICE) {
    // scratch layout: [0,64) result, then the reduction's own partials (separate arena call)
    static thread_local void *dslot = nullptr;
    if (!dslot) NUK_CUDA(cudaMalloc(&dslot, 64));
    dres = dslot;
    NUK_TRY(reduce_1d_device(red, dtype, (size_t)n, x, incx, dres, st));
  } else {
    // host operand: stage chunk by chunk, one partial per chunk, fixed-order combine on device
    size_t cap = (size_t)(option("stage_mb") > 0 ? option("stage_mb") : 32) << 20;
    if ((size_t)n * sz < cap) cap = (((size_t)n * sz + 255) / 256) * 256;
    const long chunk = (long)(cap / sz);
    const long nchunks = (n + chunk - 1) / chunk;
    Stager &sg = tl_stager;
    NUK_TRY(sg.ensure(cap, (int)option("stage_slots"), dev->device));
    static thread_local void *dpart = nullptr;
    static thread_local size_t dpart_cap = 0;
    if (dpart_cap < (size_t)(nchunks + 8) * 8) {
      if (dpart) cudaFree(dpart);
      dpart_cap = (size_t)(nchunks + 8) * 8;
      NUK_CUDA(cudaMalloc(&dpart, dpart_cap));
    }
    Opnd o{x, incx, sz, PK_HOST};
    long c0 = 0;
    for (long k = 0; c0 < n; ++k, c0 += chunk) {
      const long m = (n - c0 < chunk) ? (n - c0) : chunk;
      const int slot = (int)(k % sg.nslots);
      const void *dp;
      long di;
      if (incx == 0) {
        set_error(NUK_ERR_INVALID, "reduction over a stride-0 host operand");
        return NUK_ERR_INVALID;
      }
      NUK_TRY(stage_in(o, c0, m, sg.buf[slot][0], sg.streams[slot], &dp, &di));
      NUK_TRY(reduce_1d_device(red, dtype, (size_t)m, dp, di, static_cast<char *>(dpart) + k * sz,
                               sg.streams[slot]));
    }
    for (int s = 0; s < sg.nslots; ++s) NUK_CUDA(cudaStreamSynchronize(sg.streams[s]));
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
