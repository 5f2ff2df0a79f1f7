// nd.cu -- the nd-descriptor entry points: one launch per call instead of the reference's
// Lisp loop nest with one C call per innermost row
// (src/utils/ptr-iterate-but-inner.lisp:35-64; dense-numericals-src/utils/ptr-iterate-but-inner.lisp:42-79).
#include <algorithm>
#include "dispatch.cuh"

namespace nuk {

// Normalise (dims, strides of out/x/y) into the compact descriptor the kernels take:
//  * axes whose OUT stride is negative are walked backwards (valid for a map: no order
//    dependence), so OUT strides become non-negative;
//  * axes are ordered by decreasing OUT stride (a transposed OUT becomes row-major);
//  * size-1 axes are dropped and adjacent axes merged when all operands allow it.
// Pointers are moved to the element the normalised walk starts from.
int build_problem(MapProblem *p, int arity, int rank, const int64_t *dims, const int64_t *sx,
                  const int64_t *sy, const int64_t *so) {
  if (rank < 0 || rank > NUK_MAX_RANK) {
    set_error(NUK_ERR_INVALID, "rank %d outside [0, %d]", rank, NUK_MAX_RANK);
    return NUK_ERR_INVALID;
  }
  int64_t d[NUK_MAX_RANK], st[3 * NUK_MAX_RANK];
  for (int a = 0; a < rank; ++a) {
    if (dims[a] < 0) {
      set_error(NUK_ERR_INVALID, "negative dimension %lld at axis %d", (long long)dims[a], a);
      return NUK_ERR_INVALID;
    }
    d[a] = dims[a];
    st[0 * NUK_MAX_RANK + a] = so ? so[a] : 0;
    st[1 * NUK_MAX_RANK + a] = (p->x && sx) ? sx[a] : 0;
    st[2 * NUK_MAX_RANK + a] = (arity == 2 && p->y && sy) ? sy[a] : 0;
  }
  size_t total = 1;
  for (int a = 0; a < rank; ++a) total *= (size_t)d[a];
  // flip axes with a negative OUT stride
  // (the matching pointer shift is applied by shift() below)
  if (total > 0) {
    for (int a = 0; a < rank; ++a) {
      if (st[a] < 0) {
        for (int o = 0; o < 3; ++o) st[o * NUK_MAX_RANK + a] = -st[o * NUK_MAX_RANK + a];
      }
    }
  }
  // order axes by decreasing OUT stride (stable; broadcast-out axes with stride 0 stay put
  // relative to each other at the end would break row-major order, so only permute when
  // every OUT stride is non-zero)
  bool all_nonzero = true;
  for (int a = 0; a < rank; ++a) all_nonzero = all_nonzero && (st[a] != 0 || d[a] == 1);
  if (all_nonzero && rank > 1) {
    int perm[NUK_MAX_RANK];
    for (int a = 0; a < rank; ++a) perm[a] = a;
    std::stable_sort(perm, perm + rank, [&](int a, int b) {
      const int64_t sa = d[a] == 1 ? INT64_MAX : st[a], sb = d[b] == 1 ? INT64_MAX : st[b];
      return sa > sb;
    });
    int64_t d2[NUK_MAX_RANK], st2[3 * NUK_MAX_RANK];
    for (int a = 0; a < rank; ++a) {
      d2[a] = d[perm[a]];
      for (int o = 0; o < 3; ++o) st2[o * NUK_MAX_RANK + a] = st[o * NUK_MAX_RANK + perm[a]];
    }
    memcpy(d, d2, sizeof(d2));
    memcpy(st, st2, sizeof(st2));
  }
  const int r = coalesce(3, rank, d, st);
  p->rank = r;
  for (int a = 0; a < r; ++a) {
    p->dims[a] = d[a];
    p->so[a] = st[0 * NUK_MAX_RANK + a];
    p->sx[a] = st[1 * NUK_MAX_RANK + a];
    p->sy[a] = st[2 * NUK_MAX_RANK + a];
  }
  if (total == 0) {  // empty array: keep one zero-sized axis so that launchers see n == 0
    p->rank = 1;
    p->dims[0] = 0;
    p->so[0] = p->sx[0] = p->sy[0] = 1;
  }
  return NUK_OK;
}

static int check_ptrs(const void *out, const void *x, int need_x) {
  if (!out || (need_x && !x)) {
    set_error(NUK_ERR_INVALID, "null array pointer");
    return NUK_ERR_INVALID;
  }
  return NUK_OK;
}

static void shift(MapProblem *p, int arity, size_t in_size, size_t out_size, int rank,
                  const int64_t *dims, const int64_t *sx, const int64_t *sy, const int64_t *so) {
  // apply the pointer offsets of flipped axes (recomputed here to keep build_problem pure)
  size_t total = 1;
  for (int a = 0; a < rank; ++a) total *= (size_t)dims[a];
  if (total == 0) return;
  int64_t oo = 0, ox = 0, oy = 0;
  for (int a = 0; a < rank; ++a) {
    if (so && so[a] < 0) {
      oo += (dims[a] - 1) * so[a];
      if (p->x && sx) ox += (dims[a] - 1) * sx[a];
      if (arity == 2 && p->y && sy) oy += (dims[a] - 1) * sy[a];
    }
  }
  p->out = static_cast<char *>(p->out) + oo * (int64_t)out_size;
  if (p->x) p->x = static_cast<const char *>(p->x) + ox * (int64_t)in_size;
  if (p->y) p->y = static_cast<const char *>(p->y) + oy * (int64_t)in_size;
}

int make_problem(MapProblem *p, int arity, size_t in_size, size_t out_size, int rank,
                 const int64_t *dims, const int64_t *sx, const int64_t *sy, const int64_t *so,
                 const void *x, const void *y, void *out, cudaStream_t stream) {
  memset(p, 0, sizeof(*p));
  p->x = x;
  p->y = y;
  p->out = out;
  p->stream = stream;
  NUK_TRY(build_problem(p, arity, rank, dims, sx, sy, so));
  shift(p, arity, in_size, out_size, rank, dims, sx, sy, so);
  return NUK_OK;
}

static int load_scalar(int dtype, const void *host_scalar, Scalar64 *v) {
  const size_t sz = dtype_size(dtype);
  if (!sz || !host_scalar) {
    set_error(NUK_ERR_INVALID, "bad scalar operand (dtype %d)", dtype);
    return NUK_ERR_INVALID;
  }
  memset(v, 0, sizeof(*v));
  memcpy(v->bytes, host_scalar, sz);
  return NUK_OK;
}

}  // namespace nuk

using namespace nuk;

NUK_API int nuk_map2(int op, int dtype, int rank, const int64_t *dims, const int64_t *sx,
                     const int64_t *sy, const int64_t *so, const void *x, const void *y, void *out,
                     nuk_stream_t s) {
  ApiRange range("nuk/map2");
  NUK_TRY(check_ptrs(out, x, 1));
  NUK_TRY(check_ptrs(out, y, 1));
  const size_t sz = dtype_size(dtype);
  MapProblem p;
  NUK_TRY(make_problem(&p, 2, sz, sz, rank, dims, sx, sy, so, x, y, out, (cudaStream_t)s));
  return launch_op2(op, dtype, p);
}

NUK_API int nuk_map1(int op, int dtype, int rank, const int64_t *dims, const int64_t *sx,
                     const int64_t *so, const void *x, void *out, nuk_stream_t s) {
  ApiRange range("nuk/map1");
  NUK_TRY(check_ptrs(out, x, 1));
  const size_t sz = dtype_size(dtype);
  MapProblem p;
  NUK_TRY(make_problem(&p, 1, sz, sz, rank, dims, sx, nullptr, so, x, nullptr, out, (cudaStream_t)s));
  return launch_op1(op, dtype, p);
}

NUK_API int nuk_cmp2(int cmp, int dtype, int rank, const int64_t *dims, const int64_t *sx,
                     const int64_t *sy, const int64_t *so, const void *x, const void *y,
                     uint8_t *out, nuk_stream_t s) {
  ApiRange range("nuk/cmp2");
  NUK_TRY(check_ptrs(out, x, 1));
  NUK_TRY(check_ptrs(out, y, 1));
  MapProblem p;
  NUK_TRY(make_problem(&p, 2, dtype_size(dtype), 1, rank, dims, sx, sy, so, x, y, out, (cudaStream_t)s));
  return launch_cmp(cmp, dtype, p);
}

NUK_API int nuk_cast(int src_dtype, int dst_dtype, int rank, const int64_t *dims, const int64_t *sx,
                     const int64_t *so, const void *x, void *out, nuk_stream_t s) {
  ApiRange range("nuk/cast");
  NUK_TRY(check_ptrs(out, x, 1));
  MapProblem p;
  NUK_TRY(make_problem(&p, 1, dtype_size(src_dtype), dtype_size(dst_dtype), rank, dims, sx, nullptr,
                       so, x, nullptr, out, (cudaStream_t)s));
  return launch_cast(src_dtype, dst_dtype, p);
}

NUK_API int nuk_map2_scalar(int op, int dtype, int rank, const int64_t *dims, const int64_t *sa,
                            const int64_t *so, const void *a, const void *host_scalar,
                            int scalar_is_x, void *out, nuk_stream_t s) {
  ApiRange range("nuk/map2_scalar");
  NUK_TRY(check_ptrs(out, a, 1));
  Scalar64 v;
  NUK_TRY(load_scalar(dtype, host_scalar, &v));
  const size_t sz = dtype_size(dtype);
  MapProblem p;
  if (scalar_is_x) {
    NUK_TRY(make_problem(&p, 2, sz, sz, rank, dims, nullptr, sa, so, nullptr, a, out, (cudaStream_t)s));
    p.xv = v;
  } else {
    NUK_TRY(make_problem(&p, 2, sz, sz, rank, dims, sa, nullptr, so, a, nullptr, out, (cudaStream_t)s));
    p.yv = v;
  }
  return launch_op2(op, dtype, p);
}

NUK_API int nuk_cmp2_scalar(int cmp, int dtype, int rank, const int64_t *dims, const int64_t *sa,
                            const int64_t *so, const void *a, const void *host_scalar,
                            int scalar_is_x, uint8_t *out, nuk_stream_t s) {
  ApiRange range("nuk/cmp2_scalar");
  NUK_TRY(check_ptrs(out, a, 1));
  Scalar64 v;
  NUK_TRY(load_scalar(dtype, host_scalar, &v));
  const size_t sz = dtype_size(dtype);
  MapProblem p;
  if (scalar_is_x) {
    NUK_TRY(make_problem(&p, 2, sz, 1, rank, dims, nullptr, sa, so, nullptr, a, out, (cudaStream_t)s));
    p.xv = v;
  } else {
    NUK_TRY(make_problem(&p, 2, sz, 1, rank, dims, sa, nullptr, so, a, nullptr, out, (cudaStream_t)s));
    p.yv = v;
  }
  return launch_cmp(cmp, dtype, p);
}

NUK_API int nuk_fill(int dtype, int rank, const int64_t *dims, const int64_t *so,
                     const void *host_scalar, void *out, nuk_stream_t s) {
  ApiRange range("nuk/fill");
  NUK_TRY(check_ptrs(out, nullptr, 0));
  Scalar64 v;
  NUK_TRY(load_scalar(dtype, host_scalar, &v));
  const size_t sz = dtype_size(dtype);
  MapProblem p;
  NUK_TRY(make_problem(&p, 1, sz, sz, rank, dims, nullptr, nullptr, so, nullptr, nullptr, out,
                       (cudaStream_t)s));
  p.xv = v;
  return launch_unary(NUK_COPY, dtype, p);
}
