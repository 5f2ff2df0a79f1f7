// stats.cu -- the elementwise tails of nu:mean and nu:variance, kept as separately rounded
// steps exactly as the reference composes them:
//   mean     : out <- out / n            (src/statistics/mean.lisp:90-97; true division)
//   variance : S <- S*S; S <- S/n; Q <- Q-S; Q <- Q/(n-ddof)   (src/statistics/variance.lisp:58-68)
// -fmad=false keeps nvcc from contracting Q - S*S/n into an FMA.
#include "dispatch.cuh"
#include "ops.cuh"

namespace nuk {

template <typename T> struct VarFinishOp {
  using In = T; using Out = T;
  static constexpr int kArity = 2;
  T cnt, denom;
  NUK_D T operator()(T s, T q) const {
    T ss = s * s;
    ss = ss / cnt;
    const T d = q - ss;
    return d / denom;
  }
};

int make_problem(MapProblem *p, int arity, size_t in_size, size_t out_size, int rank,
                 const int64_t *dims, const int64_t *sx, const int64_t *sy, const int64_t *so,
                 const void *x, const void *y, void *out, cudaStream_t stream);

}  // namespace nuk
using namespace nuk;

NUK_API int nuk_div_scalar_inplace(int dtype, int64_t n, void *out, double divisor, nuk_stream_t s) {
  if (!out || n < 0) {
    set_error(NUK_ERR_INVALID, "nuk_div_scalar_inplace: bad arguments");
    return NUK_ERR_INVALID;
  }
  const int64_t dims[1] = {n}, st[1] = {1};
  MapProblem p;
  NUK_TRY(make_problem(&p, 2, dtype_size(dtype), dtype_size(dtype), 1, dims, st, nullptr, st, out,
                       nullptr, out, (cudaStream_t)s));
  if (dtype == NUK_F32) p.yv.f = (float)divisor;
  else if (dtype == NUK_F64) p.yv.d = divisor;
  else {
    set_error(NUK_ERR_UNSUPPORTED, "mean/divide: float dtypes only (package.lisp:140)");
    return NUK_ERR_UNSUPPORTED;
  }
  return launch_arith(NUK_DIV, dtype, p);
}

NUK_API int nuk_variance_finish(int dtype, int64_t n, void *s_sum, void *q_sumsq, double count,
                                double ddof, nuk_stream_t s) {
  if (!s_sum || !q_sumsq || n < 0) {
    set_error(NUK_ERR_INVALID, "nuk_variance_finish: bad arguments");
    return NUK_ERR_INVALID;
  }
  const int64_t dims[1] = {n}, st[1] = {1};
  MapProblem p;
  NUK_TRY(make_problem(&p, 2, dtype_size(dtype), dtype_size(dtype), 1, dims, st, st, st, s_sum,
                       q_sumsq, q_sumsq, (cudaStream_t)s));
  if (dtype == NUK_F32) return launch_map(p, VarFinishOp<float>{(float)count, (float)(count - ddof)});
  if (dtype == NUK_F64) return launch_map(p, VarFinishOp<double>{count, count - ddof});
  set_error(NUK_ERR_UNSUPPORTED, "variance: float dtypes only");
  return NUK_ERR_UNSUPPORTED;
}
