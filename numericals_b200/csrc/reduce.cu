// reduce.cu -- nu:sum / nu:maximum / nu:minimum (+ fused raw moments for nu:variance).
//
// Reference decomposition being replaced (src/basic-math/sum.lisp:52-71, maximum.lisp:43-75):
//   full     : ONE single-threaded BMAS_<t>sum(total, ptr, 1)
//   axis, out-size <  n : per output element a strided BMAS_<t>sum(n, in+j, stride)
//   axis, out-size >= n : n sequential row accumulations out <- out + row_i
//
// Here every reduction is deterministic: the assignment of elements to threads, the warp
// shuffle tree, the shared-memory block tree and the second pass over per-CTA partials are
// fixed functions of the problem shape (no atomics, no arrival-order dependence).
//   k_reduce_flat : contiguous full reduction, 128-bit loads, persistent grid, pass 1
//   k_reduce_nd   : strided full reduction (index decomposition), pass 1
//   k_combine     : pass 2 over partials (one CTA, fixed order)
//   k_reduce_rows : reduced axis is the contiguous one; a warp or a CTA per output
//   k_reduce_cols : inner output axis contiguous, reduced axis strided; threads own columns
//                   and walk the rows IN ORDER (== the reference's row-accumulate order);
//                   rows may be split into chunks whose partials are combined in order
//   k_reduce_any  : anything else, one thread per output
#include <cuda.h>   // CUtensorMap types only; the encoder is fetched with cudaGetDriverEntryPoint
#include <atomic>
#include <cfloat>
#include <limits>
#include <type_traits>
#include "dispatch.cuh"
#include "ops.cuh"
#include "comm.cuh"

namespace nuk {

// ------------------------------------------------------------------ reducers
template <typename T> struct AccOf {
  using type = typename std::conditional<std::is_floating_point<T>::value, T, typename WrapT<T>::type>::type;
};

template <typename T> struct SumR {
  using In = T;
  using Acc = typename AccOf<T>::type;
  static constexpr int kOuts = 1;
  NUK_D static Acc load(T v) { return (Acc)v; }
  NUK_D static Acc combine(Acc a, Acc b) { return a + b; }
  NUK_D static void store(T *o1, T *, int64_t i, Acc a) { o1[i] = (T)a; }
};
template <typename T> struct SumSqR {  // sum of x*x; the product is rounded on its own (-fmad=false)
  using In = T;
  using Acc = typename AccOf<T>::type;
  static constexpr int kOuts = 1;
  NUK_D static Acc load(T v) { return (Acc)v * (Acc)v; }
  NUK_D static Acc combine(Acc a, Acc b) { return a + b; }
  NUK_D static void store(T *o1, T *, int64_t i, Acc a) { o1[i] = (T)a; }
};
template <typename T> struct MaxR {
  using In = T;
  using Acc = T;
  static constexpr int kOuts = 1;
  NUK_D static Acc load(T v) { return v; }
  NUK_D static Acc combine(Acc a, Acc b) { return (b > a) ? b : a; }  // NaN in b is ignored
  NUK_D static void store(T *o1, T *, int64_t i, Acc a) { o1[i] = a; }
};
template <typename T> struct MinR {
  using In = T;
  using Acc = T;
  static constexpr int kOuts = 1;
  NUK_D static Acc load(T v) { return v; }
  NUK_D static Acc combine(Acc a, Acc b) { return (b < a) ? b : a; }
  NUK_D static void store(T *o1, T *, int64_t i, Acc a) { o1[i] = a; }
};
template <typename T> struct Pair { T s, q; };
template <typename T> struct MomentsR {  // (sum x, sum x*x) in one pass
  using In = T;
  using Acc = Pair<T>;
  static constexpr int kOuts = 2;
  NUK_D static Acc load(T v) { return Acc{v, v * v}; }
  NUK_D static Acc combine(Acc a, Acc b) { return Acc{a.s + b.s, a.q + b.q}; }
  NUK_D static void store(T *o1, T *o2, int64_t i, Acc a) { o1[i] = a.s; o2[i] = a.q; }
};

// ---- 8/16-bit inputs: accumulate a whole 32-bit word per instruction (dp4a / dp2a for sums, packed
// SIMD max/min for extrema) instead of extracting bytes; without this the narrow reductions are
// issue-bound (i8sum 61%, i8hmax 50% of the HBM roofline in profiles/sweep_r01_a.jsonl).
// The accumulator of a word lane is a pair of 32-bit registers (the 8-bit extrema keep even and odd bytes apart).
struct WordAcc { uint32_t a, b; };
template <class R> struct WordRed { static constexpr bool ok = false; };
#define NUK_WORDRED(RED, IDENT, ACC_EXPR, FINISH_BODY)                                          \
  template <> struct WordRed<RED> {                                                             \
    static constexpr bool ok = true;                                                            \
    using A = typename RED::Acc;                                                                \
    NUK_D static WordAcc identity() { return WordAcc{(IDENT), 0u}; }                            \
    NUK_D static WordAcc acc(WordAcc s, uint32_t w) { const uint32_t a = s.a; return WordAcc{(ACC_EXPR), 0u}; } \
    NUK_D static A finish(A r, WordAcc s) { const uint32_t a = s.a; FINISH_BODY }               \
  };
NUK_WORDRED(SumR<int8_t>, 0u, (uint32_t)__dp4a((int)w, 0x01010101, (int)a), return r + a;)
NUK_WORDRED(SumR<int16_t>, 0u, (uint32_t)__dp2a_lo((int)w, 0x00000101, (int)a), return r + a;)
#define NUK_UNPACK2(T, OP) for (int k = 0; k < 2; ++k) { const T v = (T)(a >> (16 * k)); r = (v OP r) ? v : r; } return r;
NUK_WORDRED(MaxR<int16_t>, 0x80008000u, __vmaxs2(a, w), NUK_UNPACK2(int16_t, >))
NUK_WORDRED(MinR<int16_t>, 0x7fff7fffu, __vmins2(a, w), NUK_UNPACK2(int16_t, <))
NUK_WORDRED(MaxR<uint16_t>, 0u, __vmaxu2(a, w), NUK_UNPACK2(uint16_t, >))
NUK_WORDRED(MinR<uint16_t>, 0xffffffffu, __vminu2(a, w), NUK_UNPACK2(uint16_t, <))
// 8-bit extrema: sm_100 has VIMNMX for 16-bit pairs but emulates the 4 x 8-bit forms in 6-7 instructions, so a word
// is split into its even bytes (moved up: b0 00 b2 00 read as two 16-bit values = the bytes times 256, order and
// sign preserved) and its odd bytes (masked in place), each folded with ONE 16x2 instruction (i8hmax / u8hmax: 0.82
// -> of the HBM peak with __vmaxs4 / __vmaxu4).
#define NUK_WORDRED8(RED, T, IDENT16, OP2, CMP)                                                  \
  template <> struct WordRed<RED> {                                                             \
    static constexpr bool ok = true;                                                            \
    using A = typename RED::Acc;                                                                \
    NUK_D static WordAcc identity() { return WordAcc{(IDENT16), (IDENT16)}; }                   \
    NUK_D static WordAcc acc(WordAcc s, uint32_t w) {                                           \
      return WordAcc{OP2(s.a, __byte_perm(w, 0u, 0x2404)), OP2(s.b, w & 0xff00ff00u)};          \
    }                                                                                           \
    NUK_D static A finish(A r, WordAcc s) {                                                     \
      const uint32_t h[2] = {s.a, s.b};                                                         \
      for (int k = 0; k < 4; ++k) { const T v = (T)(h[k >> 1] >> (8 + 16 * (k & 1))); r = (v CMP r) ? v : r; } \
      return r;                                                                                 \
    }                                                                                           \
  };
NUK_WORDRED8(MaxR<int8_t>, int8_t, 0x80008000u, __vmaxs2, >)
NUK_WORDRED8(MinR<int8_t>, int8_t, 0x7f007f00u, __vmins2, <)
NUK_WORDRED8(MaxR<uint8_t>, uint8_t, 0u, __vmaxu2, >)
NUK_WORDRED8(MinR<uint8_t>, uint8_t, 0xff00ff00u, __vminu2, <)

template <typename A> NUK_D A shfl_down(A v, int delta) {
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

// block tree: every thread contributes v; thread 0 returns the total.  Fixed order:
// lanes are folded 16,8,4,2,1 apart, then warp 0 folds the per-warp values the same way.
template <class R> NUK_D typename R::Acc block_reduce(typename R::Acc v, typename R::Acc init) {
  using A = typename R::Acc;
  __shared__ alignas(16) unsigned char smem_raw[sizeof(A) * (kThreads / 32)];
  A *smem = reinterpret_cast<A *>(smem_raw);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = R::combine(v, shfl_down(v, d));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // smem reuse across calls
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = (lane < kThreads / 32) ? smem[lane] : init;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = R::combine(v, shfl_down(v, d));
  }
  return v;
}

// Pass 2 inside pass 1: every CTA publishes its partial, the LAST one to arrive (an atomic ticket) folds all of
// them -- in index order through the block tree, so the result does not depend on which CTA that is -- and
// re-arms the ticket.  One launch instead of two: a full reduction of a 256 MiB array is ~45 us of
// streaming, and the second launch was ~4 us of it (8-bit sums / extrema: 0.82 -> of the HBM peak).
template <class R>
NUK_D void finish_in_last_cta(typename R::Acc v, typename R::Acc *partials, unsigned *ticket, typename R::In *o1,
                              typename R::In *o2, typename R::Acc init, const typename R::In *first) {
  using A = typename R::Acc;
  __shared__ int last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = v;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  A acc = init;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += kThreads) acc = R::combine(acc, ld_cg_any<A>(partials + i));
  acc = block_reduce<R>(acc, init);
  if (threadIdx.x == 0) {
    // `first` (float extrema only): the element at index 0.  BMAS hmax / hmin fold from x[0]
    // (oracle/bmas_oracle.c DEF_HMINMAX: r = x[0]; r = (a > r) ? a : r), so a NaN survives exactly when it sits
    // in the FIRST position and is ignored anywhere else; the tree ignores every NaN, this puts the first one back.
    if constexpr (std::is_floating_point<typename R::In>::value && std::is_same<A, typename R::In>::value) {
      if (first) {
        const typename R::In f0 = *first;
        if (f0 != f0) acc = f0;
      }
    }
    R::store(o1, o2, 0, acc);
    *ticket = 0;
  }
}

// ------------------------------------------------------------------ full, contiguous
template <class R, int UNROLL>
__global__ void __launch_bounds__(kThreads)
k_reduce_flat(size_t n, const typename R::In *x, typename R::Acc *partials, typename R::Acc init, unsigned *ticket,
              typename R::In *o1, typename R::In *o2, const typename R::In *first) {
  using T = typename R::In;
  using A = typename R::Acc;
  constexpr int E = 16 / sizeof(T);
  using P = Pack<T, E>;
  const P *xp = reinterpret_cast<const P *>(x);
  const size_t npack = n / E;
  constexpr size_t kTile = (size_t)kThreads * UNROLL;
  const size_t nfull = npack / kTile;
  A acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = init;
  WordAcc wacc[4];
  if constexpr (WordRed<R>::ok) {
#pragma unroll
    for (int w = 0; w < 4; ++w) wacc[w] = WordRed<R>::identity();
  }
  for (size_t t = blockIdx.x; t < nfull; t += gridDim.x) {
    const size_t base = t * kTile + threadIdx.x;
    P v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = ld_stream(xp + base + (size_t)u * kThreads);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if constexpr (WordRed<R>::ok) {
        uint32_t w4[4];
        memcpy(w4, &v[u], 16);
#pragma unroll
        for (int w = 0; w < 4; ++w) wacc[w] = WordRed<R>::acc(wacc[w], w4[w]);
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] = R::combine(acc[e], R::load(v[u].v[e]));
      }
    }
  }
  if constexpr (WordRed<R>::ok) {
#pragma unroll
    for (int w = 0; w < 4; ++w) acc[w] = WordRed<R>::finish(acc[w], wacc[w]);
  }
  if (blockIdx.x == nfull % gridDim.x) {
    for (size_t p = nfull * kTile + threadIdx.x; p < npack; p += kThreads) {
      const P v = ld_stream(xp + p);
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = R::combine(acc[e], R::load(v.v[e]));
    }
    const size_t i = npack * E + threadIdx.x;
    if (i < n) acc[0] = R::combine(acc[0], R::load(x[i]));
  }
  A v = acc[0];
#pragma unroll
  for (int e = 1; e < E; ++e) v = R::combine(v, acc[e]);
  v = block_reduce<R>(v, init);
  if (ticket) finish_in_last_cta<R>(v, partials, ticket, o1, o2, init, first);
  else if (threadIdx.x == 0) partials[blockIdx.x] = v;
}

// ------------------------------------------------------------------ full, strided nd
template <class R>
__global__ void __launch_bounds__(kThreads)
k_reduce_nd(size_t n, NdDesc d, const typename R::In *x, typename R::Acc *partials,
            typename R::Acc init, unsigned *ticket, typename R::In *o1, typename R::In *o2,
            const typename R::In *first) {
  using A = typename R::Acc;
  A acc = init;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    int64_t ox = 0;
    size_t rem = i;
    for (int ax = d.rank - 1; ax > 0; --ax) {
      const size_t q = rem / (size_t)d.dims[ax];
      ox += (int64_t)(rem - q * (size_t)d.dims[ax]) * d.stride[1][ax];
      rem = q;
    }
    ox += (int64_t)rem * d.stride[1][0];
    acc = R::combine(acc, R::load(x[ox]));
  }
  acc = block_reduce<R>(acc, init);
  if (ticket) finish_in_last_cta<R>(acc, partials, ticket, o1, o2, init, first);
  else if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// ------------------------------------------------------------------ pass 2
// partials laid out [count][nout]; output j combines partials[0..count)[j] in index order.
template <class R>
__global__ void __launch_bounds__(kThreads)
k_combine_cols(int64_t count, int64_t nout, const typename R::Acc *partials, typename R::In *o1,
               typename R::In *o2, int64_t ostride_outer, int64_t ncols, int64_t ostride_col,
               typename R::Acc init) {
  // nout = nouter * ncols outputs; output (a, c) lives at a*ostride_outer + c*ostride_col
  using A = typename R::Acc;
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= nout) return;
  A acc = init;
  for (int64_t k = 0; k < count; ++k) acc = R::combine(acc, partials[k * nout + j]);
  const int64_t a = j / ncols, c = j - a * ncols;
  R::store(o1, o2, a * ostride_outer + c * ostride_col, acc);
}

// ------------------------------------------------------------------ multi-GPU combine (comm.cuh)
template <class R>
__global__ void __launch_bounds__(kThreads)
k_combine_allgather(CommLink link, size_t count, const typename R::Acc *partials, typename R::In *o1,
                    typename R::In *o2, typename R::Acc init, const typename R::In *first) {
  combine_allgather_body<R>(link, count, partials, o1, o2, init, first,
                            [](typename R::Acc v, typename R::Acc i) { return block_reduce<R>(v, i); });
}
template <class R>
__global__ void __launch_bounds__(kThreads)
k_allreduce(CommLink link, int64_t count, const typename R::In *local, typename R::In *out, int vec) {
  allreduce_body<R>(link, count, local, out, vec);
}

// ------------------------------------------------------------------ axis reductions
// Output index space: up to NUK_MAX_RANK-1 coalesced dims with x strides and out strides.
struct AxisDesc {
  int orank;
  int64_t odims[NUK_MAX_RANK];
  int64_t oxs[NUK_MAX_RANK];   // x stride per output dim
  int64_t oos[NUK_MAX_RANK];   // out stride per output dim
  int64_t nout;                // product of odims
  int64_t rlen, rstride;       // reduced axis
};

NUK_D void out_offsets(const AxisDesc &d, int64_t j, int64_t *ox, int64_t *oo) {
  int64_t x = 0, o = 0;
  for (int a = d.orank - 1; a > 0; --a) {
    const int64_t q = j / d.odims[a], k = j - q * d.odims[a];
    x += k * d.oxs[a];
    o += k * d.oos[a];
    j = q;
  }
  if (d.orank > 0) {
    x += j * d.oxs[0];
    o += j * d.oos[0];
  }
  *ox = x;
  *oo = o;
}

// reduced axis contiguous (rstride == 1): WARPS_PER_OUT = 1 -> a warp per output,
// otherwise the whole CTA per output.
template <class R, bool CTA_PER_OUT>
__global__ void __launch_bounds__(kThreads)
k_reduce_rows(AxisDesc d, const typename R::In *x, typename R::In *o1, typename R::In *o2,
              typename R::Acc init, int vec_ok) {
  using T = typename R::In;
  using A = typename R::Acc;
  constexpr int E = 16 / sizeof(T);
  using P = Pack<T, E>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t j = CTA_PER_OUT ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * (kThreads / 32) + warp;
  const bool active = j < d.nout;
  int64_t ox = 0, oo = 0;
  if (active) out_offsets(d, j, &ox, &oo);
  const T *row = x + ox;
  const int tid = CTA_PER_OUT ? threadIdx.x : lane;
  const int nthr = CTA_PER_OUT ? kThreads : 32;
  A acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = init;
  if (active) {
    if (vec_ok) {
      const P *rp = reinterpret_cast<const P *>(row);
      const int64_t npack = d.rlen / E;
      int64_t p = tid;
      for (; p + 3 * nthr < npack; p += 4 * nthr) {
        P v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_stream(rp + p + u * nthr);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int e = 0; e < E; ++e) acc[e] = R::combine(acc[e], R::load(v[u].v[e]));
      }
      for (; p < npack; p += nthr) {
        const P v = ld_stream(rp + p);
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] = R::combine(acc[e], R::load(v.v[e]));
      }
      const int64_t i = npack * E + tid;
      if (i < d.rlen) acc[0] = R::combine(acc[0], R::load(row[i]));
    } else {
      for (int64_t i = tid; i < d.rlen; i += nthr) acc[0] = R::combine(acc[0], R::load(row[i * d.rstride]));
    }
  }
  A v = acc[0];
#pragma unroll
  for (int e = 1; e < E; ++e) v = R::combine(v, acc[e]);
  if (CTA_PER_OUT) {
    v = block_reduce<R>(v, init);
    if (threadIdx.x == 0 && active) R::store(o1, o2, oo, v);
  } else {
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) v = R::combine(v, shfl_down(v, dd));
    if (lane == 0 && active) R::store(o1, o2, oo, v);
  }
}

// inner output axis contiguous in x (stride 1): grid.x = column tiles, grid.y = outer * chunks.
// Rows of a chunk are walked in order; with one chunk this IS the reference's
// out <- out + row_i order (bit-identical for float sums).
template <class R, int UNROLL>
__global__ void __launch_bounds__(kThreads)
k_reduce_cols(AxisDesc d, int64_t ncols, int64_t nouter, int chunks, int64_t rows_per_chunk,
              const typename R::In *x, typename R::In *o1, typename R::In *o2,
              typename R::Acc *partials, typename R::Acc init, int vec_ok) {
  using T = typename R::In;
  using A = typename R::Acc;
  constexpr int E = 16 / sizeof(T);
  using P = Pack<T, E>;
  const int64_t a = (int64_t)blockIdx.y / chunks;          // outer output index
  const int chunk = (int)((int64_t)blockIdx.y - a * chunks);
  int64_t ox = 0, oo = 0;
  {  // outer dims = all output dims but the last
    AxisDesc outer = d;
    outer.orank = d.orank - 1;
    out_offsets(outer, a, &ox, &oo);
  }
  const int64_t r0 = (int64_t)chunk * rows_per_chunk;
  const int64_t r1 = (r0 + rows_per_chunk < d.rlen) ? r0 + rows_per_chunk : d.rlen;
  const int64_t c0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * (vec_ok ? E : 1);
  if (c0 >= ncols) return;
  const T *base = x + ox + c0;  // column stride is 1
  const int64_t ocs = d.oos[d.orank - 1];
  if (vec_ok) {
    A acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = init;
    int64_t r = r0;
    for (; r + UNROLL <= r1; r += UNROLL) {
      P v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        v[u] = ld_stream(reinterpret_cast<const P *>(base + (r + u) * d.rstride));
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] = R::combine(acc[e], R::load(v[u].v[e]));
    }
    for (; r < r1; ++r) {
      const P v = ld_stream(reinterpret_cast<const P *>(base + r * d.rstride));
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = R::combine(acc[e], R::load(v.v[e]));
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (chunks == 1) R::store(o1, o2, oo + (c0 + e) * ocs, acc[e]);
      else partials[((int64_t)chunk * nouter + a) * ncols + c0 + e] = acc[e];
    }
  } else {
    A acc = init;
    int64_t r = r0;
    for (; r + UNROLL <= r1; r += UNROLL) {
      T v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) v[u] = base[(r + u) * d.rstride];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) acc = R::combine(acc, R::load(v[u]));
    }
    for (; r < r1; ++r) acc = R::combine(acc, R::load(base[r * d.rstride]));
    if (chunks == 1) R::store(o1, o2, oo + c0 * ocs, acc);
    else partials[((int64_t)chunk * nouter + a) * ncols + c0] = acc;
  }
}

// ------------------------------------------------------------------ sequential column walk, TMA-fed
// The reference's axis-0 order (out <- out + row_i, i = 0..n-1; sum.lisp:69-71) makes every column
// a strictly sequential chain, so the only parallelism is ACROSS columns: 16384 columns = 128 CTAs
// of 128 threads, far too few loads in flight for HBM with ordinary ld.global (0.9 TB/s measured).
// Here one elected thread streams the CTA's 128-column slab through a ring of shared-memory stages
// with 1-D bulk async copies (cp.async.bulk, the TMA engine, completion on an mbarrier), 128 KB in
// flight per SM, while the 128 threads fold the rows of each landed stage IN ORDER.  Result: bit-
// identical to the reference order at the HBM roofline.
NUK_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
NUK_D void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
NUK_D void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
NUK_D void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
NUK_D bool mbar_try_wait(uint64_t *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

constexpr int kSeqThreads = 128;  // columns per CTA
constexpr int kSeqRows = 16;      // rows per stage
constexpr int kSeqStages = 8;     // stages in flight

template <class R>
__global__ void __launch_bounds__(kSeqThreads)
k_reduce_cols_seq(AxisDesc d, int64_t ncols, const typename R::In *x, typename R::In *o1,
                  typename R::In *o2, typename R::Acc init) {
  using T = typename R::In;
  using A = typename R::Acc;
  extern __shared__ __align__(128) unsigned char seq_smem[];
  T *buf = reinterpret_cast<T *>(seq_smem);  // [kSeqStages][kSeqRows][kSeqThreads]
  __shared__ __align__(8) uint64_t full[kSeqStages];

  const int tid = threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.x * kSeqThreads;
  const int ncol_here = (int)((ncols - c0 < kSeqThreads) ? (ncols - c0) : kSeqThreads);
  int64_t ox = 0, oo = 0;
  {
    AxisDesc outer = d;
    outer.orank = d.orank - 1;
    out_offsets(outer, (int64_t)blockIdx.y, &ox, &oo);
  }
  const T *base = x + ox + c0;
  const unsigned seg_bytes = (unsigned)(ncol_here * sizeof(T));
  const int64_t nstages = (d.rlen + kSeqRows - 1) / kSeqRows;

  if (tid == 0) {
    for (int s = 0; s < kSeqStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // warp 0 fills slot it % kSeqStages with stage `it`: lane 0 arms the barrier with the byte count,
  // then lane r issues the bulk copy of row r (one UBLKCP per lane, so the 16 row copies of a
  // stage are issued together instead of serially by one thread)
  auto issue = [&](int64_t it) {
    const int slot = (int)(it % kSeqStages);
    const int64_t r0 = it * kSeqRows;
    const int nr = (int)((d.rlen - r0 < kSeqRows) ? (d.rlen - r0) : kSeqRows);
    if (tid == 0) mbar_expect_tx(&full[slot], (unsigned)nr * seg_bytes);
    __syncwarp();
    if (tid < nr)
      bulk_g2s(buf + ((size_t)slot * kSeqRows + tid) * kSeqThreads, base + (r0 + tid) * d.rstride, seg_bytes,
               &full[slot]);
  };
  if (tid < 32)
    for (int64_t it = 0; it < kSeqStages && it < nstages; ++it) issue(it);

  A acc = init;
  for (int64_t it = 0; it < nstages; ++it) {
    const int slot = (int)(it % kSeqStages);
    const unsigned parity = (unsigned)((it / kSeqStages) & 1);
    while (!mbar_try_wait(&full[slot], parity)) {
    }
    const int64_t r0 = it * kSeqRows;
    const int nr = (int)((d.rlen - r0 < kSeqRows) ? (d.rlen - r0) : kSeqRows);
    if (tid < ncol_here) {
      const T *col = buf + (size_t)slot * kSeqRows * kSeqThreads + tid;
      if (nr == kSeqRows) {
        T v[kSeqRows];
#pragma unroll
        for (int r = 0; r < kSeqRows; ++r) v[r] = col[r * kSeqThreads];
#pragma unroll
        for (int r = 0; r < kSeqRows; ++r) acc = R::combine(acc, R::load(v[r]));   // rows in order
      } else {
        for (int r = 0; r < nr; ++r) acc = R::combine(acc, R::load(col[r * kSeqThreads]));
      }
    }
    __syncthreads();  // every thread is done reading the slot before it is refilled
    if (tid < 32 && it + kSeqStages < nstages) issue(it + kSeqStages);
  }
  if (tid < ncol_here) R::store(o1, o2, oo + (c0 + tid) * d.oos[d.orank - 1], acc);
}

// Same walk fed by ONE tensor-map TMA request per stage (cp.async.bulk.tensor, a 128-column x
// 16-row box = 16 KB for doubles) instead of 16 one-row bulk copies: the per-request cost of the
// TMA engine is what limited the 1-D version (2.9 TB/s).  Tensor: (columns, rows, outer) with byte
// strides; out-of-bounds parts of a box are zero-filled and still counted, so every stage expects
// the full box.
template <class R>
__global__ void __launch_bounds__(kSeqThreads)
k_reduce_cols_tma(const __grid_constant__ CUtensorMap tmap, AxisDesc d, int64_t ncols,
                  typename R::In *o1, typename R::In *o2, typename R::Acc init) {
  using T = typename R::In;
  using A = typename R::Acc;
  extern __shared__ __align__(128) unsigned char seq_smem[];
  T *buf = reinterpret_cast<T *>(seq_smem);  // [kSeqStages][kSeqRows][kSeqThreads]
  __shared__ __align__(8) uint64_t full[kSeqStages];
  const int tid = threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.x * kSeqThreads;
  const int ncol_here = (int)((ncols - c0 < kSeqThreads) ? (ncols - c0) : kSeqThreads);
  const int outer = (int)blockIdx.y;
  const int64_t oo = d.orank >= 2 ? (int64_t)outer * d.oos[0] : 0;
  const int64_t nstages = (d.rlen + kSeqRows - 1) / kSeqRows;
  constexpr unsigned kBoxBytes = kSeqRows * kSeqThreads * sizeof(T);

  if (tid == 0) {
    for (int s = 0; s < kSeqStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int64_t it) {  // thread 0
    const int slot = (int)(it % kSeqStages);
    mbar_expect_tx(&full[slot], kBoxBytes);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(buf + (size_t)slot * kSeqRows * kSeqThreads)),
        "l"(&tmap), "r"((int)c0), "r"((int)(it * kSeqRows)), "r"(outer), "r"(smem_u32(&full[slot]))
        : "memory");
  };
  if (tid == 0)
    for (int64_t it = 0; it < kSeqStages && it < nstages; ++it) issue(it);

  A acc = init;
  for (int64_t it = 0; it < nstages; ++it) {
    const int slot = (int)(it % kSeqStages);
    const unsigned parity = (unsigned)((it / kSeqStages) & 1);
    while (!mbar_try_wait(&full[slot], parity)) {
    }
    const int64_t r0 = it * kSeqRows;
    const int nr = (int)((d.rlen - r0 < kSeqRows) ? (d.rlen - r0) : kSeqRows);
    if (tid < ncol_here) {
      const T *col = buf + (size_t)slot * kSeqRows * kSeqThreads + tid;
      if (nr == kSeqRows) {
        T v[kSeqRows];
#pragma unroll
        for (int r = 0; r < kSeqRows; ++r) v[r] = col[r * kSeqThreads];
#pragma unroll
        for (int r = 0; r < kSeqRows; ++r) acc = R::combine(acc, R::load(v[r]));   // rows in order
      } else {
        for (int r = 0; r < nr; ++r) acc = R::combine(acc, R::load(col[r * kSeqThreads]));
      }
    }
    __syncthreads();
    if (tid == 0 && it + kSeqStages < nstages) issue(it + kSeqStages);
  }
  if (tid < ncol_here) R::store(o1, o2, oo + (c0 + tid) * d.oos[d.orank - 1], acc);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libnuk does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// fallback: one thread per output, rows in order
template <class R>
__global__ void __launch_bounds__(kThreads)
k_reduce_any(AxisDesc d, const typename R::In *x, typename R::In *o1, typename R::In *o2,
             typename R::Acc init) {
  using A = typename R::Acc;
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= d.nout) return;
  int64_t ox, oo;
  out_offsets(d, j, &ox, &oo);
  A acc = init;
  for (int64_t r = 0; r < d.rlen; ++r) acc = R::combine(acc, R::load(x[ox + r * d.rstride]));
  R::store(o1, o2, oo, acc);
}

// ------------------------------------------------------------------ host side
template <class R> static int run_full(const ReduceProblem &p, typename R::Acc init) {
  using T = typename R::In;
  using A = typename R::Acc;
  const T *x = static_cast<const T *>(p.x);
  T *o1 = static_cast<T *>(p.out), *o2 = static_cast<T *>(p.out2);
  size_t n = 1;
  for (int a = 0; a < p.rank; ++a) n *= (size_t)p.dims[a];
  const bool contig = p.rank == 0 || (p.rank == 1 && (p.sx[0] == 1 || n <= 1));
  constexpr int UNROLL = 4;
  constexpr int E = 16 / sizeof(T);
  int grid;
  void *scratch_ptr = nullptr;
  unsigned *ticket = nullptr;   // non-null: ONE launch, the last CTA to arrive folds the partials (finish_in_last_cta)
  // pass-1 partials: the thread's arena, or the communicator's own buffer for a sharded reduction (whose pass 2
  // is the exchange kernel)
  auto partial_space = [&](size_t bytes, void **ptr) -> int {
    if (!p.link) return scratch_with_ticket(p.stream, bytes, ptr, &ticket);
    if (bytes > p.link_partial_bytes) {
      set_error(NUK_ERR_INVALID, "internal: communicator partial buffer too small (%zu > %zu)", bytes,
                p.link_partial_bytes);
      return NUK_ERR_INVALID;
    }
    *ptr = p.link_partials;
    return NUK_OK;
  };
  const T *first = ((p.red == NUK_HMAX || p.red == NUK_HMIN) && !(p.flags & NUK_REDUCE_NO_SEED) && n > 0) ? x : nullptr;
  if (contig && aligned_to(x, 16)) {
    const size_t tiles = n / E / ((size_t)kThreads * UNROLL) + 1;
    // 8-bit extrema: a quarter of the CTAs (4 per SM instead of 16).  Their fold is ALU work (two extracts and two 16x2
    // maxima per word, issued at half rate) and a 2^28-element array is only ~7 tiles per CTA at 16 per SM, so the
    // per-CTA epilogue (unpack, block tree, ticket) weighs in: i8hmax 0.868 -> 0.917, u8hmax 0.856 -> 0.930 of the
    // measured peak, the sums and the wider extrema are best at 16 (profiles/ab_reduce_grid_r02.txt)
    constexpr bool kByteExtremum = sizeof(T) == 1 && WordRed<R>::ok && !std::is_same<R, SumR<T>>::value;
    const int mult = (int)opt(OPT_REDUCE_GRID_MULT);
    grid = grid_for(tiles, kByteExtremum ? (mult + 3) / 4 : mult);
    NUK_TRY(partial_space(sizeof(A) * (size_t)grid, &scratch_ptr));
    k_reduce_flat<R, UNROLL><<<grid, kThreads, 0, p.stream>>>(n, x, static_cast<A *>(scratch_ptr), init, ticket, o1, o2,
                                                              first);
    NUK_TRY(post_launch("k_reduce_flat"));
  } else {
    NdDesc d;
    memset(&d, 0, sizeof(d));
    d.rank = p.rank > 0 ? p.rank : 1;
    d.dims[0] = 1;
    for (int a = 0; a < p.rank; ++a) {
      d.dims[a] = p.dims[a];
      d.stride[1][a] = p.sx[a];
    }
    const size_t blocks = (n + kThreads * 4 - 1) / (kThreads * 4);
    grid = grid_for(blocks ? blocks : 1, (int)opt(OPT_REDUCE_GRID_MULT));
    NUK_TRY(partial_space(sizeof(A) * (size_t)grid, &scratch_ptr));
    k_reduce_nd<R><<<grid, kThreads, 0, p.stream>>>(n, d, x, static_cast<A *>(scratch_ptr), init, ticket, o1, o2, first);
    NUK_TRY(post_launch("k_reduce_nd"));
  }
  if (p.link) {
    k_combine_allgather<R><<<1, kThreads, 0, p.stream>>>(*p.link, (size_t)grid, static_cast<const A *>(scratch_ptr), o1,
                                                         o2, init, first);
    return post_launch("k_combine_allgather");
  }
  return NUK_OK;
}

template <class R> static int run_axis(const ReduceProblem &p, typename R::Acc init) {
  using T = typename R::In;
  using A = typename R::Acc;
  const T *x = static_cast<const T *>(p.x);
  T *o1 = static_cast<T *>(p.out), *o2 = static_cast<T *>(p.out2);
  const DeviceInfo *dev = device_info();
  if (!dev) return nuk_last_status();
  constexpr int E = 16 / sizeof(T);

  // output index space = dims without the axis, coalesced over (x stride, out stride)
  AxisDesc d;
  memset(&d, 0, sizeof(d));
  d.rlen = p.dims[p.axis];
  d.rstride = p.sx[p.axis];
  {
    int64_t od[NUK_MAX_RANK], st[3 * NUK_MAX_RANK];
    memset(st, 0, sizeof(st));
    int r = 0;
    for (int a = 0; a < p.rank; ++a) {
      if (a == p.axis) continue;
      od[r] = p.dims[a];
      st[0 * NUK_MAX_RANK + r] = p.so[a];
      st[1 * NUK_MAX_RANK + r] = p.sx[a];
      ++r;
    }
    r = coalesce(2, r, od, st);
    d.orank = r;
    d.nout = 1;
    for (int a = 0; a < r; ++a) {
      d.odims[a] = od[a];
      d.oos[a] = st[0 * NUK_MAX_RANK + a];
      d.oxs[a] = st[1 * NUK_MAX_RANK + a];
      d.nout *= od[a];
    }
  }
  if (d.nout == 0) return NUK_OK;
  if (d.rlen == 0) {  // empty axis: every output is the identity
    AxisDesc e = d;
    e.rlen = 0;
    const unsigned g = (unsigned)((d.nout + kThreads - 1) / kThreads);
    k_reduce_any<R><<<g, kThreads, 0, p.stream>>>(e, x, o1, o2, init);
    return post_launch("k_reduce_any");
  }

  auto rows_aligned = [&](int64_t elt_multiple) {
    if (!aligned_to(x, 16)) return false;
    for (int a = 0; a < d.orank; ++a)
      if (d.oxs[a] % elt_multiple != 0) return false;
    return true;
  };

  // ---- reduced axis contiguous
  if (d.rstride == 1 || d.rlen == 1) {
    const int vec_ok = (d.rstride == 1 && rows_aligned(E)) ? 1 : 0;
    if (d.rlen >= 2048) {
      k_reduce_rows<R, true><<<(unsigned)d.nout, kThreads, 0, p.stream>>>(d, x, o1, o2, init, vec_ok);
    } else {
      const unsigned g = (unsigned)((d.nout + kThreads / 32 - 1) / (kThreads / 32));
      k_reduce_rows<R, false><<<g, kThreads, 0, p.stream>>>(d, x, o1, o2, init, vec_ok);
    }
    return post_launch("k_reduce_rows");
  }
  // ---- inner output axis contiguous in x: column walk
  if (d.orank >= 1 && d.oxs[d.orank - 1] == 1) {
    const int64_t ncols = d.odims[d.orank - 1];
    const int64_t nouter = d.nout / ncols;
    bool vec = aligned_to(x, 16) && (d.rstride % E == 0) && (ncols % E == 0);
    for (int a = 0; a + 1 < d.orank; ++a) vec = vec && (d.oxs[a] % E == 0);
    // TMA-fed sequential walk: reference order (bit-exact) at HBM speed.  Needs 16-byte aligned row
    // segments; chosen whenever it covers most of the chip, or when the caller asks for the
    // reference order.
    {
      const int64_t seq_tiles = (ncols + kSeqThreads - 1) / kSeqThreads;
      bool ok = sizeof(T) >= 4 && aligned_to(x, 16) && ((d.rstride * (int64_t)sizeof(T)) % 16 == 0) &&
                ((ncols * (int64_t)sizeof(T)) % 16 == 0) && d.rlen >= 4 * kSeqRows && nouter <= 65535;
      for (int a = 0; a + 1 < d.orank; ++a) ok = ok && ((d.oxs[a] * (int64_t)sizeof(T)) % 16 == 0);
      const bool covers = seq_tiles * nouter * 4 >= (int64_t)dev->sm_count * 3;   // >= 75% of the SMs
      if (ok && (covers || (p.flags & NUK_REDUCE_REF_ORDER)) && opt(OPT_SEQ_COLS) != 0) {
        const size_t smem = (size_t)kSeqStages * kSeqRows * kSeqThreads * sizeof(T);
        // the opt-in is per (kernel instantiation, device): one bit per device ordinal
        static std::atomic<uint64_t> attr_set{0};
        const uint64_t dev_bit = 1ull << (dev->device & 63);
        if (!(attr_set.load(std::memory_order_acquire) & dev_bit)) {
          NUK_CUDA(cudaFuncSetAttribute(k_reduce_cols_seq<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
          attr_set.fetch_or(dev_bit, std::memory_order_release);
        }
        dim3 grid((unsigned)seq_tiles, (unsigned)nouter);
        // preferred: one tensor-map TMA request per stage
        EncodeTiledFn enc = encode_tiled_fn();
        if (enc && d.orank <= 2 && d.rstride > 0 && (d.orank < 2 || d.oxs[0] > 0) && opt(OPT_SEQ_COLS) != 2) {
          CUtensorMap tmap;
          const cuuint64_t gdim[3] = {(cuuint64_t)ncols, (cuuint64_t)d.rlen, (cuuint64_t)nouter};
          const cuuint64_t outer_stride = d.orank == 2 ? (cuuint64_t)d.oxs[0] : (cuuint64_t)(d.rlen * d.rstride);
          const cuuint64_t gstr[2] = {(cuuint64_t)d.rstride * sizeof(T), outer_stride * sizeof(T)};
          const cuuint32_t box[3] = {(cuuint32_t)kSeqThreads, (cuuint32_t)kSeqRows, 1u};
          const cuuint32_t estr[3] = {1u, 1u, 1u};
          const CUresult r = enc(&tmap, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32,
                                 3, const_cast<T *>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r == CUDA_SUCCESS) {
            static std::atomic<uint64_t> attr2_set{0};
            if (!(attr2_set.load(std::memory_order_acquire) & dev_bit)) {
              NUK_CUDA(cudaFuncSetAttribute(k_reduce_cols_tma<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
              attr2_set.fetch_or(dev_bit, std::memory_order_release);
            }
            k_reduce_cols_tma<R><<<grid, kSeqThreads, smem, p.stream>>>(tmap, d, ncols, o1, o2, init);
            return post_launch("k_reduce_cols_tma");
          }
        }
        k_reduce_cols_seq<R><<<grid, kSeqThreads, smem, p.stream>>>(d, ncols, x, o1, o2, init);
        return post_launch("k_reduce_cols_seq");
      }
    }
    const int64_t cols_per_cta = (int64_t)kThreads * (vec ? E : 1);
    const int64_t col_tiles = (ncols + cols_per_cta - 1) / cols_per_cta;
    int chunks = 1;
    int64_t rows_per_chunk = d.rlen;
    if (!(p.flags & NUK_REDUCE_REF_ORDER) && d.orank <= 2) {
      const int64_t want = (int64_t)dev->sm_count * 8;
      const int64_t have = col_tiles * nouter;
      if (have < want) {
        int64_t c = (want + have - 1) / have;
        const int64_t min_rows = opt(OPT_COL_SPLIT_ROWS) > 0 ? opt(OPT_COL_SPLIT_ROWS) : 256;
        const int64_t maxc = (d.rlen + min_rows - 1) / min_rows;
        if (c > maxc) c = maxc;
        if (c < 1) c = 1;
        rows_per_chunk = (d.rlen + c - 1) / c;
        chunks = (int)((d.rlen + rows_per_chunk - 1) / rows_per_chunk);
      }
    }
    if (nouter * chunks <= 65535 && col_tiles <= 0x7fffffff) {
      void *part = nullptr;
      if (chunks > 1) NUK_TRY(scratch(p.stream, sizeof(A) * (size_t)chunks * (size_t)d.nout, &part));
      dim3 grid((unsigned)col_tiles, (unsigned)(nouter * chunks));
      k_reduce_cols<R, 4><<<grid, kThreads, 0, p.stream>>>(d, ncols, nouter, chunks, rows_per_chunk, x, o1,
                                                           o2, static_cast<A *>(part), init, vec ? 1 : 0);
      NUK_TRY(post_launch("k_reduce_cols"));
      if (chunks > 1) {
        // outputs: (a, c) -> out offset. k_combine_cols needs a single outer stride, so it is
        // only used when the outer output space is one dim; otherwise combine per outer index.
        if (d.orank <= 2) {
          const int64_t os_outer = d.orank == 2 ? d.oos[0] : 0;
          const unsigned g = (unsigned)((d.nout + kThreads - 1) / kThreads);
          k_combine_cols<R><<<g, kThreads, 0, p.stream>>>(chunks, d.nout, static_cast<const A *>(part), o1, o2,
                                                          os_outer, ncols, d.oos[d.orank - 1], init);
          return post_launch("k_combine_cols");
        }
        set_error(NUK_ERR_INVALID, "internal: split column reduce with outer rank > 1");
        return NUK_ERR_INVALID;
      }
      return NUK_OK;
    }
  }
  // ---- anything else
  const unsigned g = (unsigned)((d.nout + kThreads - 1) / kThreads);
  k_reduce_any<R><<<g, kThreads, 0, p.stream>>>(d, x, o1, o2, init);
  return post_launch("k_reduce_any");
}

template <typename T> static T type_lowest_full() {
  if (std::is_floating_point<T>::value) return -std::numeric_limits<T>::infinity();
  return std::numeric_limits<T>::lowest();
}
template <typename T> static T type_highest_full() {
  if (std::is_floating_point<T>::value) return std::numeric_limits<T>::infinity();
  return std::numeric_limits<T>::max();
}

template <class R> static int run(const ReduceProblem &p, typename R::Acc init_full, typename R::Acc init_axis) {
  if (p.axis < 0) return run_full<R>(p, init_full);
  return run_axis<R>(p, init_axis);
}

template <typename T> static int by_red(const ReduceProblem &p) {
  using A = typename AccOf<T>::type;
  if (p.out2) {
    if constexpr (std::is_floating_point<T>::value) {
      if (p.red == NUK_SUM) return run<MomentsR<T>>(p, Pair<T>{0, 0}, Pair<T>{0, 0});
    }
    set_error(NUK_ERR_UNSUPPORTED, "moments need a float dtype and red == NUK_SUM");
    return NUK_ERR_UNSUPPORTED;
  }
  switch (p.red) {
    case NUK_SUM: return run<SumR<T>>(p, (A)0, (A)0);
    case NUK_SUMSQ: return run<SumSqR<T>>(p, (A)0, (A)0);
    // full reductions start from -inf/+inf (== BMAS hmax over the data itself); axis
    // reductions start from the reference's finite sentinels type-min / type-max
    // (common.lisp:46-131, used at maximum.lisp:52)
    case NUK_HMAX: return run<MaxR<T>>(p, type_lowest_full<T>(), std::numeric_limits<T>::lowest());
    case NUK_HMIN: return run<MinR<T>>(p, type_highest_full<T>(), std::numeric_limits<T>::max());
  }
  set_error(NUK_ERR_UNSUPPORTED, "unknown reduction %d", p.red);
  return NUK_ERR_UNSUPPORTED;
}

int launch_reduce(const ReduceProblem &p) {
  int dt = p.dtype;
  // nu:sum of unsigned arrays goes through the signed symbol (package.lisp:149-151): same bits
  if (p.red == NUK_SUM || p.red == NUK_SUMSQ) {
    if (dt == NUK_U64) dt = NUK_I64;
    if (dt == NUK_U32) dt = NUK_I32;
    if (dt == NUK_U16) dt = NUK_I16;
    if (dt == NUK_U8) dt = NUK_I8;
  }
  switch (dt) {
    case NUK_F32: return by_red<float>(p);
    case NUK_F64: return by_red<double>(p);
    case NUK_I64: return by_red<int64_t>(p);
    case NUK_I32: return by_red<int32_t>(p);
    case NUK_I16: return by_red<int16_t>(p);
    case NUK_I8: return by_red<int8_t>(p);
    case NUK_U64: return by_red<uint64_t>(p);
    case NUK_U32: return by_red<uint32_t>(p);
    case NUK_U16: return by_red<uint16_t>(p);
    case NUK_U8: return by_red<uint8_t>(p);
  }
  set_error(NUK_ERR_UNSUPPORTED, "no reduction kernel for dtype %d", p.dtype);
  return NUK_ERR_UNSUPPORTED;
}

// rank-ordered all-reduce of `count` elements (one mailbox slot at most; comm.cu loops over larger payloads)
template <typename T> static int allreduce_typed(const CommLink &link, int red, int64_t count, const void *local,
                                                 void *out, cudaStream_t stream) {
  // enough CTAs to keep the NVLink stores in flight, never more than the flag table holds; a function of
  // `count` only, so every rank launches the same grid
  int64_t g = (count * (int64_t)sizeof(T) + 2047) / 2048;     // 2 KB (128 threads x 16 B) per CTA and peer
  if (g < 1) g = 1;
  if (g > kCommMaxCtas) g = kCommMaxCtas;
  const T *l = static_cast<const T *>(local);
  T *o = static_cast<T *>(out);
  // a per-rank choice (the cut of the payload into CTA pieces does not depend on it, comm.cuh)
  const int vec = (aligned_to(local, 16) && aligned_to(out, 16)) ? 1 : 0;
  switch (red) {
    case NUK_SUM:
    case NUK_SUMSQ: k_allreduce<SumR<T>><<<(unsigned)g, kThreads, 0, stream>>>(link, count, l, o, vec); break;
    case NUK_HMAX: k_allreduce<MaxR<T>><<<(unsigned)g, kThreads, 0, stream>>>(link, count, l, o, vec); break;
    case NUK_HMIN: k_allreduce<MinR<T>><<<(unsigned)g, kThreads, 0, stream>>>(link, count, l, o, vec); break;
    default:
      set_error(NUK_ERR_UNSUPPORTED, "all-reduce: unknown reduction %d", red);
      return NUK_ERR_UNSUPPORTED;
  }
  return post_launch("k_allreduce");
}
int launch_allreduce(const CommLink &link, int red, int dtype, int64_t count, const void *local, void *out,
                     cudaStream_t stream) {
  int dt = dtype;
  if (red == NUK_SUM || red == NUK_SUMSQ) {   // unsigned sums through the signed symbol: same bits
    if (dt == NUK_U64) dt = NUK_I64;
    if (dt == NUK_U32) dt = NUK_I32;
    if (dt == NUK_U16) dt = NUK_I16;
    if (dt == NUK_U8) dt = NUK_I8;
  }
  switch (dt) {
    case NUK_F32: return allreduce_typed<float>(link, red, count, local, out, stream);
    case NUK_F64: return allreduce_typed<double>(link, red, count, local, out, stream);
    case NUK_I64: return allreduce_typed<int64_t>(link, red, count, local, out, stream);
    case NUK_I32: return allreduce_typed<int32_t>(link, red, count, local, out, stream);
    case NUK_I16: return allreduce_typed<int16_t>(link, red, count, local, out, stream);
    case NUK_I8: return allreduce_typed<int8_t>(link, red, count, local, out, stream);
    case NUK_U64: return allreduce_typed<uint64_t>(link, red, count, local, out, stream);
    case NUK_U32: return allreduce_typed<uint32_t>(link, red, count, local, out, stream);
    case NUK_U16: return allreduce_typed<uint16_t>(link, red, count, local, out, stream);
    case NUK_U8: return allreduce_typed<uint8_t>(link, red, count, local, out, stream);
  }
  set_error(NUK_ERR_UNSUPPORTED, "all-reduce: no kernel for dtype %d", dtype);
  return NUK_ERR_UNSUPPORTED;
}

int reduce_1d_device(int red, int dtype, size_t n, const void *x, int64_t incx, void *out_dev,
                     cudaStream_t stream, int flags) {
  ReduceProblem p;
  memset(&p, 0, sizeof(p));
  p.flags = flags;
  p.red = red;
  p.dtype = dtype;
  p.rank = 1;
  p.dims[0] = (int64_t)n;
  p.sx[0] = incx;
  p.axis = -1;
  p.x = x;
  p.out = out_dev;
  p.stream = stream;
  return launch_reduce(p);
}

}  // namespace nuk

using namespace nuk;

static int reduce_entry(int red, int dtype, int rank, const int64_t *dims, const int64_t *sx, int axis,
                        const int64_t *so, const void *x, void *out, void *out2, int flags,
                        nuk_stream_t s) {
  ApiRange range("nuk/reduce");
  if (rank < 0 || rank > NUK_MAX_RANK || !x || !out || axis >= rank) {
    set_error(NUK_ERR_INVALID, "bad reduce descriptor (rank %d, axis %d)", rank, axis);
    return NUK_ERR_INVALID;
  }
  if (!device_info()) return nuk_last_status();
  ReduceProblem p;
  memset(&p, 0, sizeof(p));
  p.red = red;
  p.dtype = dtype;
  p.axis = axis;
  p.x = x;
  p.out = out;
  p.out2 = out2;
  p.flags = flags;
  p.stream = (cudaStream_t)s;
  if (axis < 0) {
    // full: coalesce x alone so that a contiguous array becomes rank <= 1
    int64_t d[NUK_MAX_RANK], st[NUK_MAX_RANK * 1 + NUK_MAX_RANK];
    for (int a = 0; a < rank; ++a) {
      d[a] = dims[a];
      st[a] = sx[a];
    }
    const int r = coalesce(1, rank, d, st);
    p.rank = r;
    for (int a = 0; a < r; ++a) {
      p.dims[a] = d[a];
      p.sx[a] = st[a];
    }
    for (int a = 0; a < rank; ++a)
      if (dims[a] == 0) { p.rank = 1; p.dims[0] = 0; p.sx[0] = 1; }
  } else {
    p.rank = rank;
    int k = 0;
    for (int a = 0; a < rank; ++a) {
      p.dims[a] = dims[a];
      p.sx[a] = sx[a];
      p.so[a] = (a == axis) ? 0 : (so ? so[k++] : 0);
    }
  }
  return launch_reduce(p);
}

NUK_API int nuk_reduce(int red, int dtype, int rank, const int64_t *dims, const int64_t *sx, int axis,
                       const int64_t *so, const void *x, void *out, int flags, nuk_stream_t s) {
  return reduce_entry(red, dtype, rank, dims, sx, axis, so, x, out, nullptr, flags, s);
}

NUK_API int nuk_moments(int dtype, int rank, const int64_t *dims, const int64_t *sx, int axis,
                        const int64_t *so, const void *x, void *sum_out, void *sumsq_out, int flags,
                        nuk_stream_t s) {
  if (!sumsq_out) {
    set_error(NUK_ERR_INVALID, "nuk_moments: null sumsq_out");
    return NUK_ERR_INVALID;
  }
  return reduce_entry(NUK_SUM, dtype, rank, dims, sx, axis, so, x, sum_out, sumsq_out, flags, s);
}
