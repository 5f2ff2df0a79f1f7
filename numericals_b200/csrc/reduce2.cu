// reduce2.cu -- SURVEY 8(f) "next" rows that share the reduction skeleton:
//   BMAS_<t>dot        (vdot = multiply + sum fused: 2 reads per element instead of the
//                       reference's multiply pass + sum pass; src/linalg/vdot.lisp:15-25)
//   BMAS_<t>himax/himin (nu:arg-maximum / nu:arg-minimum, src/basic-math/arg-maximum.lisp:31-73):
//                       index of the FIRST extremum.
// Deterministic two-pass reductions like reduce.cu; operands may be device or host pointers.
#include <limits>
#include "dispatch.cuh"
#include "ops.cuh"
#include "staging.cuh"

namespace nuk {

template <typename T> struct DTypeOf;
template <> struct DTypeOf<float> { static constexpr int value = NUK_F32; };
template <> struct DTypeOf<double> { static constexpr int value = NUK_F64; };
template <> struct DTypeOf<int64_t> { static constexpr int value = NUK_I64; };
template <> struct DTypeOf<int32_t> { static constexpr int value = NUK_I32; };
template <> struct DTypeOf<int16_t> { static constexpr int value = NUK_I16; };
template <> struct DTypeOf<int8_t> { static constexpr int value = NUK_I8; };

template <typename T> struct AccT {
  using type = typename std::conditional<std::is_floating_point<T>::value, T, typename WrapT<T>::type>::type;
};

template <typename A> NUK_D A shfl_down_any(A v, int delta) {
  constexpr int W = (sizeof(A) + 3) / 4;
  unsigned w[W];
  memset(w, 0, sizeof(w));
  memcpy(w, &v, sizeof(A));
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = __shfl_down_sync(0xffffffffu, w[i], delta);
  A r;
  memcpy(&r, w, sizeof(A));
  return r;
}
template <typename A, class Combine> NUK_D A block_fold(A v, A init, Combine comb) {
  __shared__ alignas(16) unsigned char raw[sizeof(A) * (kThreads / 32)];
  A *sm = reinterpret_cast<A *>(raw);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = comb(v, shfl_down_any(v, d));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = (lane < kThreads / 32) ? sm[lane] : init;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = comb(v, shfl_down_any(v, d));
  }
  return v;
}

// ------------------------------------------------------------------ dot
// One launch: every CTA publishes its partial, the last one to arrive (atomic ticket in the arena header) folds
// them in index order and re-arms the ticket (as reduce.cu:finish_in_last_cta).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_dot(size_t n, const T *x, int64_t ix, const T *y, int64_t iy, typename AccT<T>::type *partials,
      int vec, unsigned *ticket, T *out) {
  using A = typename AccT<T>::type;
  constexpr int E = 16 / sizeof(T);
  using P = Pack<T, E>;
  A acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = (A)0;
  auto comb = [](A a, A b) { return a + b; };
  if (vec) {
    const P *xp = reinterpret_cast<const P *>(x), *yp = reinterpret_cast<const P *>(y);
    const size_t npack = n / E;
    constexpr int U = 4;
    const size_t tile = (size_t)kThreads * U, nfull = npack / tile;
    for (size_t t = blockIdx.x; t < nfull; t += gridDim.x) {
      const size_t base = t * tile + threadIdx.x;
      P a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) a[u] = ld_stream(xp + base + (size_t)u * kThreads);
#pragma unroll
      for (int u = 0; u < U; ++u) b[u] = ld_stream(yp + base + (size_t)u * kThreads);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if constexpr (std::is_same<T, int8_t>::value) {
          // four byte products per IDP.4A; the wrap-around sum is the same in any grouping (i8dot: 0.80 of the
          // HBM peak with a sign extension and an IMAD per byte)
          uint32_t wa[4], wb[4];
          memcpy(wa, &a[u], 16);
          memcpy(wb, &b[u], 16);
#pragma unroll
          for (int w = 0; w < 4; ++w) acc[w] = (A)__dp4a((int)wa[w], (int)wb[w], (int)acc[w]);
        } else {
#pragma unroll
          for (int e = 0; e < E; ++e) acc[e] = acc[e] + (A)a[u].v[e] * (A)b[u].v[e];  // product rounded, then added
        }
      }
    }
    if (blockIdx.x == nfull % gridDim.x) {
      for (size_t p = nfull * tile + threadIdx.x; p < npack; p += kThreads) {
        const P a = ld_stream(xp + p), b = ld_stream(yp + p);
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] = acc[e] + (A)a.v[e] * (A)b.v[e];
      }
      const size_t i = npack * E + threadIdx.x;
      if (i < n) acc[0] = acc[0] + (A)x[i] * (A)y[i];
    }
  } else {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
      acc[0] = acc[0] + (A)x[(int64_t)i * ix] * (A)y[(int64_t)i * iy];
  }
  A v = acc[0];
#pragma unroll
  for (int e = 1; e < E; ++e) v = v + acc[e];
  v = block_fold(v, (A)0, comb);
  __shared__ int last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = v;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  A tot = (A)0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += kThreads) tot = tot + __ldcg(partials + i);
  tot = block_fold(tot, (A)0, comb);
  if (threadIdx.x == 0) {
    out[0] = (T)tot;
    *ticket = 0;
  }
}

template <typename T>
static int dot_device(size_t n, const T *x, int64_t ix, const T *y, int64_t iy, T *out_dev, cudaStream_t st) {
  using A = typename AccT<T>::type;
  constexpr int E = 16 / sizeof(T);
  const int vec = (ix == 1 && iy == 1 && aligned_to(x, 16) && aligned_to(y, 16)) ? 1 : 0;
  const size_t blocks = n / E / ((size_t)kThreads * 4) + 1;
  const int grid = grid_for(blocks, (int)opt(OPT_REDUCE_GRID_MULT));
  void *part = nullptr;
  unsigned *ticket = nullptr;
  NUK_TRY(scratch_with_ticket(st, sizeof(A) * (size_t)grid, &part, &ticket));
  k_dot<T><<<grid, kThreads, 0, st>>>(n, x, ix, y, iy, static_cast<A *>(part), vec, ticket, out_dev);
  return post_launch("k_dot");
}

// ------------------------------------------------------------------ arg-reductions
template <typename T> struct ValIdx { T v; int64_t i; };

template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(kThreads)
k_argred(size_t n, const T *x, int64_t ix, int64_t index_base, ValIdx<T> *partials, T init) {
  using A = ValIdx<T>;
  auto comb = [](A a, A b) {
    const bool better = IS_MAX ? (b.v > a.v) : (b.v < a.v);
    const bool tie_first = (b.v == a.v) && (b.i < a.i);
    return (better || tie_first) ? b : a;
  };
  A acc{init, INT64_MAX};
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
    acc = comb(acc, A{x[(int64_t)i * ix], (int64_t)i + index_base});
  acc = block_fold(acc, A{init, INT64_MAX}, comb);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(kThreads)
k_argred_combine(size_t count, const ValIdx<T> *partials, ValIdx<T> *out, int accumulate, T init) {
  using A = ValIdx<T>;
  auto comb = [](A a, A b) {
    const bool better = IS_MAX ? (b.v > a.v) : (b.v < a.v);
    const bool tie_first = (b.v == a.v) && (b.i < a.i);
    return (better || tie_first) ? b : a;
  };
  A acc = (accumulate && threadIdx.x == 0) ? out[0] : A{init, INT64_MAX};
  for (size_t i = threadIdx.x; i < count; i += kThreads) acc = comb(acc, partials[i]);
  acc = block_fold(acc, A{init, INT64_MAX}, comb);
  if (threadIdx.x == 0) out[0] = acc;
}

template <typename T, bool IS_MAX>
static int argred_device(size_t n, const T *x, int64_t ix, int64_t index_base, ValIdx<T> *out_dev,
                         int accumulate, cudaStream_t st) {
  T init;
  if (std::is_floating_point<T>::value)
    init = IS_MAX ? -std::numeric_limits<T>::infinity() : std::numeric_limits<T>::infinity();
  else
    init = IS_MAX ? std::numeric_limits<T>::lowest() : std::numeric_limits<T>::max();
  const size_t blocks = (n + (size_t)kThreads * 4 - 1) / ((size_t)kThreads * 4);
  const int grid = grid_for(blocks ? blocks : 1, (int)opt(OPT_REDUCE_GRID_MULT));
  void *part = nullptr;
  NUK_TRY(scratch(st, sizeof(ValIdx<T>) * (size_t)grid, &part));
  k_argred<T, IS_MAX><<<grid, kThreads, 0, st>>>(n, x, ix, index_base, static_cast<ValIdx<T> *>(part), init);
  NUK_TRY(post_launch("k_argred"));
  k_argred_combine<T, IS_MAX><<<1, kThreads, 0, st>>>((size_t)grid, static_cast<const ValIdx<T> *>(part),
                                                      out_dev, accumulate, init);
  return post_launch("k_argred_combine");
}

// ------------------------------------------------------------------ value-returning wrappers
// Host operands travel through the staging ring (staging.cuh): chunk k on slot k mod S produces ONE
// partial (a dot product / a (value, index) pair) in a per-thread device buffer; the partials are then
// combined in chunk order by the same pass-2 kernels, so the result is a fixed function of (n, chunk size).
template <typename T>
static int dot_typed(long n, const void *x, long ix, const void *y, long iy, void *result) {
  using A = typename AccT<T>::type;
  memset(result, 0, sizeof(T));
  if (n <= 0) return NUK_OK;
  const DeviceInfo *dev = device_info();
  if (!dev) return nuk_last_status();
  cudaStream_t st = current_stream();
  void *pin = nullptr, *dresv = nullptr;
  NUK_TRY(pinned_slot(&pin));
  NUK_TRY(thread_buffer(BUF_RESULT, 64, &dresv));
  T *dres = static_cast<T *>(dresv);
  const PtrKind xk = classify_ptr(x, true), yk = classify_ptr(y, true);
  if (xk == PK_DEVICE && yk == PK_DEVICE) {
    NUK_TRY(dot_device<T>((size_t)n, static_cast<const T *>(x), ix, static_cast<const T *>(y), iy, dres, st));
  } else {
    const size_t cap = stage_chunk_bytes((size_t)n * sizeof(T));
    const long chunk = (long)(cap / sizeof(T));
    const long nchunks = (n + chunk - 1) / chunk;
    Stager &sg = thread_stager();
    NUK_TRY(sg.ensure(cap, (int)opt(OPT_STAGE_SLOTS), dev->device));
    NUK_CUDA(cudaEventRecord(sg.entry, st));
    for (int s = 0; s < sg.nslots; ++s) NUK_CUDA(cudaStreamWaitEvent(sg.streams[s], sg.entry, 0));
    void *dpartv = nullptr;
    NUK_TRY(thread_buffer(BUF_PARTS, (size_t)(nchunks + 8) * sizeof(A), &dpartv));
    T *dpart = static_cast<T *>(dpartv);
    const Opnd ox{x, ix, sizeof(T), xk}, oy{y, iy, sizeof(T), yk};
    long c0 = 0;
    for (long k = 0; c0 < n; ++k, c0 += chunk) {
      const long m = (n - c0 < chunk) ? (n - c0) : chunk;
      const int slot = (int)(k % sg.nslots);
      NUK_TRY(sg.begin_slot(slot));
      const void *xp, *yp;
      long xi, yi;
      if (xk == PK_DEVICE) { xp = static_cast<const char *>(x) + (int64_t)c0 * ix * (int64_t)sizeof(T); xi = ix; }
      else NUK_TRY(sg.in(slot, 0, ox, c0, m, &xp, &xi));
      if (yk == PK_DEVICE) { yp = static_cast<const char *>(y) + (int64_t)c0 * iy * (int64_t)sizeof(T); yi = iy; }
      else NUK_TRY(sg.in(slot, 1, oy, c0, m, &yp, &yi));
      NUK_TRY(dot_device<T>((size_t)m, static_cast<const T *>(xp), xi, static_cast<const T *>(yp), yi, dpart + k,
                            sg.streams[slot]));
    }
    NUK_TRY(sg.finish_all());
    NUK_TRY(reduce_partials_device(NUK_SUM, DTypeOf<T>::value, (size_t)nchunks, dpart, dres, st));
  }
  NUK_CUDA(cudaMemcpyAsync(pin, dres, sizeof(T), cudaMemcpyDeviceToHost, st));
  NUK_CUDA(cudaStreamSynchronize(st));
  memcpy(result, pin, sizeof(T));
  return NUK_OK;
}

template <typename T, bool IS_MAX>
static int argred_typed(long n, const void *x, long ix, long *result) {
  *result = 0;
  if (n <= 0 || ix == 0) return NUK_OK;
  const DeviceInfo *dev = device_info();
  if (!dev) return nuk_last_status();
  cudaStream_t st = current_stream();
  void *pin = nullptr, *dresv = nullptr;
  NUK_TRY(pinned_slot(&pin));
  NUK_TRY(thread_buffer(BUF_RESULT, 64, &dresv));
  ValIdx<T> *dres = static_cast<ValIdx<T> *>(dresv);
  const PtrKind xk = classify_ptr(x, true);
  if (xk == PK_DEVICE) {
    NUK_TRY((argred_device<T, IS_MAX>((size_t)n, static_cast<const T *>(x), ix, 0, dres, 0, st)));
  } else {
    const size_t cap = stage_chunk_bytes((size_t)n * sizeof(T));
    const long chunk = (long)(cap / sizeof(T));
    const long nchunks = (n + chunk - 1) / chunk;
    Stager &sg = thread_stager();
    NUK_TRY(sg.ensure(cap, (int)opt(OPT_STAGE_SLOTS), dev->device));
    void *dpartv = nullptr;
    NUK_TRY(thread_buffer(BUF_PARTS, (size_t)(nchunks + 8) * sizeof(ValIdx<T>), &dpartv));
    ValIdx<T> *dpart = static_cast<ValIdx<T> *>(dpartv);
    const Opnd ox{x, ix, sizeof(T), xk};
    long c0 = 0;
    for (long k = 0; c0 < n; ++k, c0 += chunk) {
      const long m = (n - c0 < chunk) ? (n - c0) : chunk;
      const int slot = (int)(k % sg.nslots);
      NUK_TRY(sg.begin_slot(slot));
      const void *xp;
      long xi;
      NUK_TRY(sg.in(slot, 0, ox, c0, m, &xp, &xi));
      NUK_TRY((argred_device<T, IS_MAX>((size_t)m, static_cast<const T *>(xp), xi, c0, dpart + k, 0, sg.streams[slot])));
    }
    NUK_TRY(sg.finish_all());
    T init;
    if (std::is_floating_point<T>::value)
      init = IS_MAX ? -std::numeric_limits<T>::infinity() : std::numeric_limits<T>::infinity();
    else
      init = IS_MAX ? std::numeric_limits<T>::lowest() : std::numeric_limits<T>::max();
    k_argred_combine<T, IS_MAX><<<1, kThreads, 0, st>>>((size_t)nchunks, dpart, dres, 0, init);
    NUK_TRY(post_launch("k_argred_combine"));
  }
  NUK_CUDA(cudaMemcpyAsync(pin, dres, sizeof(ValIdx<T>), cudaMemcpyDeviceToHost, st));
  NUK_CUDA(cudaStreamSynchronize(st));
  ValIdx<T> r;
  memcpy(&r, pin, sizeof(r));
  *result = (r.i == INT64_MAX) ? 0 : (long)r.i;  // all-NaN input: index 0, like the scalar loop
  return NUK_OK;
}

int dot_value(int dtype, long n, const void *x, long incx, const void *y, long incy, void *result) {
  nuk_clear_error();
  switch (dtype) {
    case NUK_F32: return dot_typed<float>(n, x, incx, y, incy, result);
    case NUK_F64: return dot_typed<double>(n, x, incx, y, incy, result);
    case NUK_I64: return dot_typed<int64_t>(n, x, incx, y, incy, result);
    case NUK_I32: return dot_typed<int32_t>(n, x, incx, y, incy, result);
    case NUK_I16: return dot_typed<int16_t>(n, x, incx, y, incy, result);
    case NUK_I8: return dot_typed<int8_t>(n, x, incx, y, incy, result);
  }
  set_error(NUK_ERR_UNSUPPORTED, "dot: no kernel for dtype %d", dtype);
  return NUK_ERR_UNSUPPORTED;
}

template <bool IS_MAX> static int argred_by_type(int dtype, long n, const void *x, long incx, long *r) {
  switch (dtype) {
    case NUK_F32: return argred_typed<float, IS_MAX>(n, x, incx, r);
    case NUK_F64: return argred_typed<double, IS_MAX>(n, x, incx, r);
    case NUK_I64: return argred_typed<int64_t, IS_MAX>(n, x, incx, r);
    case NUK_I32: return argred_typed<int32_t, IS_MAX>(n, x, incx, r);
    case NUK_I16: return argred_typed<int16_t, IS_MAX>(n, x, incx, r);
    case NUK_I8: return argred_typed<int8_t, IS_MAX>(n, x, incx, r);
    case NUK_U64: return argred_typed<uint64_t, IS_MAX>(n, x, incx, r);
    case NUK_U32: return argred_typed<uint32_t, IS_MAX>(n, x, incx, r);
    case NUK_U16: return argred_typed<uint16_t, IS_MAX>(n, x, incx, r);
    case NUK_U8: return argred_typed<uint8_t, IS_MAX>(n, x, incx, r);
  }
  set_error(NUK_ERR_UNSUPPORTED, "arg-reduction: no kernel for dtype %d", dtype);
  return NUK_ERR_UNSUPPORTED;
}
int argred_value(int is_max, int dtype, long n, const void *x, long incx, long *result) {
  nuk_clear_error();
  return is_max ? argred_by_type<true>(dtype, n, x, incx, result) : argred_by_type<false>(dtype, n, x, incx, result);
}

}  // namespace nuk

// ------------------------------------------------------------------ nd arg-reduction along one axis
// nu:arg-maximum / nu:arg-minimum with :axis (src/basic-math/arg-maximum.lisp:31-73): per output
// element the index of the FIRST extremum along `axis`, as (signed-byte 64).  One warp per output,
// lanes stride over the axis, (value, index) pairs folded by shuffles with the first-index rule.
namespace nuk {
struct ArgAxisDesc {
  int orank;
  int64_t odims[NUK_MAX_RANK], oxs[NUK_MAX_RANK], oos[NUK_MAX_RANK];
  int64_t nout, rlen, rstride;
};
template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(kThreads)
k_argred_axis(ArgAxisDesc d, const T *x, int64_t *out, T init) {
  using A = ValIdx<T>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t j = (int64_t)blockIdx.x * (kThreads / 32) + warp;
  if (j >= d.nout) return;
  int64_t ox = 0, oo = 0;
  for (int a = d.orank - 1; a >= 0; --a) {
    const int64_t q = j / d.odims[a], k = j - q * d.odims[a];
    ox += k * d.oxs[a];
    oo += k * d.oos[a];
    j = q;
  }
  auto comb = [](A a, A b) {
    const bool better = IS_MAX ? (b.v > a.v) : (b.v < a.v);
    const bool tie_first = (b.v == a.v) && (b.i < a.i);
    return (better || tie_first) ? b : a;
  };
  A acc{init, INT64_MAX};
  for (int64_t r = lane; r < d.rlen; r += 32) acc = comb(acc, A{x[ox + r * d.rstride], r});
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc = comb(acc, shfl_down_any(acc, s));
  if (lane == 0) out[oo] = acc.i == INT64_MAX ? 0 : acc.i;
}
template <typename T, bool IS_MAX> static int argaxis_launch(const ArgAxisDesc &d, const void *x, int64_t *out, cudaStream_t st) {
  T init;
  if (std::is_floating_point<T>::value)
    init = IS_MAX ? -std::numeric_limits<T>::infinity() : std::numeric_limits<T>::infinity();
  else
    init = IS_MAX ? std::numeric_limits<T>::lowest() : std::numeric_limits<T>::max();
  const unsigned grid = (unsigned)((d.nout + kThreads / 32 - 1) / (kThreads / 32));
  k_argred_axis<T, IS_MAX><<<grid, kThreads, 0, st>>>(d, static_cast<const T *>(x), out, init);
  return post_launch("k_argred_axis");
}
template <bool IS_MAX> static int argaxis_by_type(int dtype, const ArgAxisDesc &d, const void *x, int64_t *out, cudaStream_t st) {
  switch (dtype) {
    case NUK_F32: return argaxis_launch<float, IS_MAX>(d, x, out, st);
    case NUK_F64: return argaxis_launch<double, IS_MAX>(d, x, out, st);
    case NUK_I64: return argaxis_launch<int64_t, IS_MAX>(d, x, out, st);
    case NUK_I32: return argaxis_launch<int32_t, IS_MAX>(d, x, out, st);
    case NUK_I16: return argaxis_launch<int16_t, IS_MAX>(d, x, out, st);
    case NUK_I8: return argaxis_launch<int8_t, IS_MAX>(d, x, out, st);
    case NUK_U64: return argaxis_launch<uint64_t, IS_MAX>(d, x, out, st);
    case NUK_U32: return argaxis_launch<uint32_t, IS_MAX>(d, x, out, st);
    case NUK_U16: return argaxis_launch<uint16_t, IS_MAX>(d, x, out, st);
    case NUK_U8: return argaxis_launch<uint8_t, IS_MAX>(d, x, out, st);
  }
  set_error(NUK_ERR_UNSUPPORTED, "arg-reduction: no kernel for dtype %d", dtype);
  return NUK_ERR_UNSUPPORTED;
}
}  // namespace nuk

NUK_API int nuk_argreduce(int is_max, int dtype, int rank, const int64_t *dims, const int64_t *sx, int axis,
                          const int64_t *so, const void *x, int64_t *out, nuk_stream_t s) {
  using namespace nuk;
  if (rank < 1 || rank > NUK_MAX_RANK || axis < 0 || axis >= rank || !x || !out) {
    set_error(NUK_ERR_INVALID, "bad arg-reduce descriptor (rank %d, axis %d)", rank, axis);
    return NUK_ERR_INVALID;
  }
  if (!device_info()) return nuk_last_status();
  ArgAxisDesc d;
  memset(&d, 0, sizeof(d));
  d.rlen = dims[axis];
  d.rstride = sx[axis];
  d.nout = 1;
  int k = 0;
  for (int a = 0; a < rank; ++a) {
    if (a == axis) continue;
    d.odims[k] = dims[a];
    d.oxs[k] = sx[a];
    d.oos[k] = so ? so[k] : 0;
    d.nout *= dims[a];
    ++k;
  }
  d.orank = k;
  if (d.nout == 0) return NUK_OK;
  return is_max ? argaxis_by_type<true>(dtype, d, x, out, (cudaStream_t)s)
                : argaxis_by_type<false>(dtype, d, x, out, (cudaStream_t)s);
}
