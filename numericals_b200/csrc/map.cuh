// map.cuh -- the three elementwise kernels of libnuk and their launcher.
//
//   k_map_flat : everything contiguous (or scalar).  ONE tile of kThreads*UNROLL packs per CTA (the
//                hardware CTA scheduler spreads tiles over the HBM channels better than a persistent
//                grid-stride loop: 6.8 vs 5.9 TB/s, profiles/membench_r01.txt); one pack = 16 bytes of
//                the widest operand type, all loads of a tile issued before the first use (UNROLL
//                independent 128-bit requests per operand per thread in flight), streaming stores.
//                Functor traits: kMapUnroll, kMapMinBlocks (register cap), kSmemBytes + stage() (lookup
//                tables copied into shared memory once per CTA, after the operand loads are issued),
//                kHasFast + fast() (branch-free main path for the whole pack, flagged elements re-done),
//                kWarpSplit (the same pair split by a per-warp vote, rare path called), kRegionSplit + in_region()
//                + region() (one vote per tile: a warp entirely inside a cheap closed region runs only that),
//                kFastBatch / kBatchCall (whole-tile batch: an A/B option, no functor uses it by default).
//   k_map_rows : inner axis contiguous for OUT, inputs contiguous or broadcast (stride 0)
//                along it; outer axes arbitrary.  A CTA owns a column tile x a group of
//                consecutive rows; an operand that does not vary with the row (row-vector
//                broadcast) is staged ONCE in registers, a column-vector operand is one
//                broadcast load per row.  This replaces the reference's one-C-call-per-row
//                loop (src/utils/ptr-iterate-but-inner.lisp:35-64).
//   k_map_nd   : any strides (0 / negative / transposed), rank <= 8: per-element index
//                decomposition (no division when rank == 1, which is the BMAS strided call).
//
// A functor F provides: In (input element type), Out, kArity (1|2) and
//   __device__ Out operator()(In a [, In b]).
// OUT may alias X and/or Y exactly: every thread reads the elements it is going to
// overwrite before it writes them, and no thread reads another thread's elements.
#pragma once
#include <utility>
#include "common.cuh"

namespace nuk {

template <typename T> struct Operand {
  const T *ptr;   // nullptr => scalar by value
  Scalar64 val;
};

template <class F> struct PackOf {
  using In = typename F::In;
  using Out = typename F::Out;
  static constexpr int kWide = sizeof(In) > sizeof(Out) ? sizeof(In) : sizeof(Out);
  static constexpr int E = 16 / kWide;  // elements per pack
  using PIn = Pack<In, E>;
  using POut = Pack<Out, E>;
};

// packs per thread in the flat kernel: 4 unless the functor says otherwise (F::kMapUnroll);
// exp_f32 runs best with 2 (5.9 vs 5.2 TB/s), log/sin/tanh with 4 (profiles/membench_r01.txt)
template <class F, class = void> struct MapUnroll { static constexpr int value = 4; };
template <class F> struct MapUnroll<F, std::void_t<decltype(F::kMapUnroll)>> {
  static constexpr int value = F::kMapUnroll;
};

// A functor that reads lookup tables declares kSmemBytes and stage(void *smem): every CTA copies the
// tables into shared memory once (cooperatively) before its first element.
template <class F, class = void> struct HasTables { static constexpr bool value = false; };
template <class F> struct HasTables<F, std::void_t<decltype(F::kSmemBytes)>> { static constexpr bool value = true; };
template <class F> NUK_D void stage_tables(F &f) {
  if constexpr (HasTables<F>::value) {
    __shared__ __align__(16) unsigned char smem[F::kSmemBytes];
    f.stage(smem);
    __syncthreads();
  }
}

// minimum resident CTAs per SM the flat kernel is compiled for (caps registers): F::kMapMinBlocks
template <class F, class = void> struct MapMinBlocks { static constexpr int value = 0; };   // 0 = unspecified (a 1 here made ptxas spend 55 registers on exp_f32 instead of 32)
#if defined(NUK_MINB_OVERRIDE)   // register-cap experiments: one value for every functor of the translation unit
template <class F> struct MapMinBlocks<F, std::void_t<decltype(F::kArity)>> { static constexpr int value = NUK_MINB_OVERRIDE; };
#else
template <class F> struct MapMinBlocks<F, std::void_t<decltype(F::kMapMinBlocks)>> {
  static constexpr int value = F::kMapMinBlocks;
};
#endif

// A functor may split itself into a BRANCH-FREE main path `fast(a [, b], bool &slow)` that is valid for
// ordinary arguments and raises `slow` otherwise, and the complete operator() that handles everything.
// The flat kernel then evaluates fast() for all the elements of a pack -- straight-line code, so the
// compiler interleaves the independent dependency chains of the elements (the per-element special-case
// branches of operator() pin instruction-level parallelism at 1 and leave the kernel latency-bound at
// the 40-50 % occupancy these register-heavy functions get) -- and re-does flagged elements afterwards.
#ifndef NUK_ROWS_FAST
#define NUK_ROWS_FAST 1   // the rows kernel takes the split too (A/B: tools/time_rows.py)
#endif
template <class F, class = void> struct HasFast { static constexpr bool value = false; };
template <class F> struct HasFast<F, std::void_t<decltype(F::kHasFast)>> { static constexpr bool value = F::kHasFast; };
// ... and may keep the per-element form for its scalar-operand tiles (kScalarSplit = false: pow f32 spills there)
template <class F, class = void> struct ScalarSplit { static constexpr bool value = true; };
template <class F> struct ScalarSplit<F, std::void_t<decltype(F::kScalarSplit)>> { static constexpr bool value = F::kScalarSplit; };
// kWarpSplit: the same fast() / operator() pair, split per WARP instead of per pack: every element evaluates the
// branch-free fast() and the warp votes; only if some lane raised `slow` does the whole warp call the complete
// function (a real CALL).  The branch on a vote is uniform, so the hot path carries no BSSY / BSYNC pair and no
// jump to a join point -- two instructions fewer per element than the per-element `if (rare)` form (tanh f32:
// 43 -> 41), without keeping a pack's inputs and flags alive as the pack-level split does.  Only used in the
// all-vector tile of the flat kernel, where all 32 lanes of a warp are converged by construction.
template <class F, class = void> struct WarpSplit { static constexpr bool value = false; };
#if defined(NUK_WARP_SPLIT_ALL)   // A/B: every split functor of the translation unit takes the per-warp form
template <class F> struct WarpSplit<F, std::void_t<decltype(F::kHasFast)>> { static constexpr bool value = F::kHasFast; };
#else
template <class F> struct WarpSplit<F, std::void_t<decltype(F::kWarpSplit)>> { static constexpr bool value = F::kWarpSplit; };
#endif
// kFastBatch: the pack-level split over the WHOLE thread tile: fast() for all UNROLL x E elements first, one test of
// the collected flags, flagged elements re-done, then the stores.  The table-driven double functors wait on two
// dependent shared-memory reads per element (ncu: 7-10 warps per issue stalled on the short scoreboard, issue slots
// 40-58 % busy for sinh / cosh); with every element of the tile in one straight-line block the compiler overlaps
// those reads across elements instead of across two.
template <class F, class = void> struct FastBatch { static constexpr bool value = false; };
template <class F> struct FastBatch<F, std::void_t<decltype(F::kFastBatch)>> { static constexpr bool value = F::kFastBatch; };
// kBatchCall: the re-do of a flagged element CALLS the complete function (one copy per kernel) instead of inlining it
template <class F, class = void> struct BatchCall { static constexpr bool value = false; };
template <class F> struct BatchCall<F, std::void_t<decltype(F::kBatchCall)>> { static constexpr bool value = F::kBatchCall; };
// kRegionSplit: a functor whose function has a cheap closed region (asinh on |x| <= 1, acosh on (1, 3]: a polynomial
// instead of the logarithm path) provides in_region(a) and region(a) next to the fast() / operator() pair of its
// GENERAL path.  The thread tests its whole tile, the warp votes ONCE, and a warp whose 32 x UNROLL x E arguments are
// all inside runs region() for them -- straight-line code, nothing else; any other warp runs the general pack-level
// split unchanged, where fast() flags the in-region arguments too and the complete function (which branches to the
// same polynomial) re-does them: an element's result never depends on the path its neighbours forced.
template <class F, class = void> struct RegionSplit { static constexpr bool value = false; };
template <class F> struct RegionSplit<F, std::void_t<decltype(F::kRegionSplit)>> { static constexpr bool value = F::kRegionSplit; };
template <class F> __device__ __noinline__ typename F::Out apply_called(F f, typename F::In a, typename F::In b) {
  if constexpr (F::kArity == 2) return f(a, b);
  else return f(a);
}
template <class F> NUK_D typename F::Out apply_fast(const F &f, typename F::In a, typename F::In b, bool &slow) {
  if constexpr (F::kArity == 2) return f.fast(a, b, slow);
  else return f.fast(a, slow);
}

template <class F> NUK_D typename F::Out apply(const F &f, typename F::In a, typename F::In b) {
  if constexpr (F::kArity == 2) return f(a, b);
  else return f(a);
}
// The complete function where it is the RARE path of a split functor (re-doing a flagged element): called, not
// inlined -- inlined once per element of a pack the special-case code is most of the kernel and pushes the hot
// paths of neighbouring elements kilobytes apart in the instruction cache (NUK_RARE_CALL, A/B in profiles/).
#ifndef NUK_RARE_CALL
#define NUK_RARE_CALL 0
#endif
#if NUK_RARE_CALL
template <class F> __device__ __noinline__ typename F::Out apply_rare(F f, typename F::In a, typename F::In b) {
  if constexpr (F::kArity == 2) return f(a, b);
  else return f(a);
}
#else
template <class F> NUK_D typename F::Out apply_rare(const F &f, typename F::In a, typename F::In b) { return apply(f, a, b); }
#endif

// ------------------------------------------------------------------ flat kernel
// MODE: 2 = one kernel for all-vector and scalar-operand tiles (default); a split functor is instantiated twice,
// 0 = all-vector tiles only, 1 = scalar-operand tiles only, so that each gets its own register allocation (with
// both branches in one kernel the split of the scalar branch spilled and slowed the vector path: pow f64 1.13 ->
// 1.41 ms, profiles/fast_split_ab_r01.txt).
template <class F, class = void> struct ScalarMinBlocks { static constexpr int value = MapMinBlocks<F>::value; };
template <class F> struct ScalarMinBlocks<F, std::void_t<decltype(F::kScalarMinBlocks)>> {
  static constexpr int value = F::kScalarMinBlocks;   // register cap of the MODE 1 instantiation
};
template <class F, int UNROLL, int MODE = 2>
__global__ void __launch_bounds__(kThreads, MODE == 1 ? ScalarMinBlocks<F>::value : MapMinBlocks<F>::value)
k_map_flat(size_t n, const typename F::In *x, const typename F::In *y, typename F::Out *out,
           int x_scalar_arg, int y_scalar_arg, Scalar64 xv, Scalar64 yv, F f) {
  // MODE 0 is only launched with all-vector operands: the scalar bookkeeping folds away at compile time (it was
  // a third of the ~33-instruction per-thread prologue, i.e. ~0.7 instructions per element of a 16-element thread)
  const int x_scalar = (MODE == 0) ? 0 : x_scalar_arg, y_scalar = (MODE == 0) ? 0 : y_scalar_arg;
  // Programmatic dependent launch (launch_flat below): this grid may have been scheduled while its predecessor
  // in the stream was still draining.  Let OUR successor be scheduled as early, then wait for the predecessor to
  // complete (and its writes to be visible) before touching memory.  Both are no-ops for an ordinary launch.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  using In = typename F::In;
  using Out = typename F::Out;
  using P = PackOf<F>;
  constexpr int E = P::E;
  using PIn = typename P::PIn;
  using POut = typename P::POut;

  In xs = In(), ys = In();
  if (x_scalar) xs = x ? x[0] : scalar_as<In>(xv);
  if (F::kArity == 2 && y_scalar) ys = y ? y[0] : scalar_as<In>(yv);

  const size_t npack = n / E;
  constexpr size_t kTile = (size_t)kThreads * UNROLL;
  const size_t nfull = npack / kTile;
  const PIn *xp = reinterpret_cast<const PIn *>(x);
  const PIn *yp = reinterpret_cast<const PIn *>(y);
  POut *op = reinterpret_cast<POut *>(out);

  // ONE tile per CTA (grid = nfull + 1): on B200 the hardware CTA scheduler spreads the tiles over
  // the HBM channels better than a persistent grid-stride loop does -- 6.8 vs 5.9 TB/s for a unary
  // f32 map, profiles/membench_r01.txt -- so the "grid-stride" loop degenerates to one iteration.
  if (MODE != 1 && blockIdx.x < nfull && (MODE == 0 || (!x_scalar && (F::kArity == 1 || !y_scalar)))) {
    // all-vector tile (the hot case): no per-element scalar selects
    const size_t base = (size_t)blockIdx.x * kTile + threadIdx.x;
    PIn a[UNROLL], b[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) a[u] = ld_stream(xp + base + (size_t)u * kThreads);
    if (F::kArity == 2) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) b[u] = ld_stream(yp + base + (size_t)u * kThreads);
    }
    stage_tables(f);   // after the operand loads are in flight: the table copy hides under their latency
    if constexpr (RegionSplit<F>::value) {
      bool inside = true;
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
        for (int e = 0; e < E; ++e) inside &= f.in_region(a[u].v[e]);
      }
      if (__all_sync(0xffffffffu, inside)) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          POut r;
#pragma unroll
          for (int e = 0; e < E; ++e) r.v[e] = f.region(a[u].v[e]);
          st_stream(op + base + (size_t)u * kThreads, r);
        }
        return;
      }
    }
    if constexpr (FastBatch<F>::value) {
      POut r[UNROLL];
      uint32_t flags = 0;
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          bool s1 = false;
          r[u].v[e] = apply_fast(f, a[u].v[e], F::kArity == 2 ? b[u].v[e] : In(), s1);
          flags |= (uint32_t)s1 << (u * E + e);
        }
      }
      if (flags) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
          for (int e = 0; e < E; ++e)
            if ((flags >> (u * E + e)) & 1u) {
              if constexpr (BatchCall<F>::value) r[u].v[e] = apply_called(f, a[u].v[e], F::kArity == 2 ? b[u].v[e] : In());
              else r[u].v[e] = apply_rare(f, a[u].v[e], F::kArity == 2 ? b[u].v[e] : In());
            }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) st_stream(op + base + (size_t)u * kThreads, r[u]);
      return;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      POut r;
      if constexpr (Word4<F>::ok) {
        uint32_t wa[4], wb[4], wr[4];
        memcpy(wa, &a[u], 16);
        memcpy(wb, &b[u], 16);
#pragma unroll
        for (int w = 0; w < 4; ++w) wr[w] = Word4<F>::op(wa[w], wb[w]);
        memcpy(&r, wr, 16);
      } else if constexpr (WarpSplit<F>::value) {
#if defined(NUK_WARP_PACK)   // A/B: one vote per pack instead of one per element
        bool slow = false;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          bool s1 = false;
          r.v[e] = apply_fast(f, a[u].v[e], F::kArity == 2 ? b[u].v[e] : In(), s1);
          slow |= s1;
        }
        if (__any_sync(0xffffffffu, slow)) {
#pragma unroll
          for (int e = 0; e < E; ++e) r.v[e] = apply_called(f, a[u].v[e], F::kArity == 2 ? b[u].v[e] : In());
        }
#else
#pragma unroll
        for (int e = 0; e < E; ++e) {
          bool slow = false;
          r.v[e] = apply_fast(f, a[u].v[e], F::kArity == 2 ? b[u].v[e] : In(), slow);
          // every lane of a flagged warp takes the call (restricting it to the flagged lanes puts a BSSY / BSYNC pair on
          // the hot path: tanh f32 43 -> 46 instructions per element; the complete function returns fast()'s bits there)
          if (__any_sync(0xffffffffu, slow)) r.v[e] = apply_called(f, a[u].v[e], F::kArity == 2 ? b[u].v[e] : In());
        }
#endif
      } else if constexpr (HasFast<F>::value) {
        bool slow[E], any = false;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          slow[e] = false;
          r.v[e] = apply_fast(f, a[u].v[e], F::kArity == 2 ? b[u].v[e] : In(), slow[e]);
          any |= slow[e];
        }
        if (any) {
#pragma unroll
          for (int e = 0; e < E; ++e)
            if (slow[e]) r.v[e] = apply_rare(f, a[u].v[e], F::kArity == 2 ? b[u].v[e] : In());
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) r.v[e] = apply(f, a[u].v[e], F::kArity == 2 ? b[u].v[e] : In());
      }
      st_stream(op + base + (size_t)u * kThreads, r);
    }
  } else if (MODE != 0 && blockIdx.x < nfull) {
    // a scalar operand (uniform across the grid): same tile with broadcast values
    stage_tables(f);
    const size_t base = (size_t)blockIdx.x * kTile + threadIdx.x;
    PIn a[UNROLL], b[UNROLL];
    if (!x_scalar) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) a[u] = ld_stream(xp + base + (size_t)u * kThreads);
    }
    if (F::kArity == 2 && !y_scalar) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) b[u] = ld_stream(yp + base + (size_t)u * kThreads);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      POut r;
      if constexpr (Word4<F>::ok) {
        // 8-bit operands: four elements per 32-bit SIMD-in-a-word instruction
        uint32_t wa[4], wb[4], wr[4];
        if (x_scalar) { wa[0] = wa[1] = wa[2] = wa[3] = (uint32_t)(uint8_t)xs * 0x01010101u; }
        else memcpy(wa, &a[u], 16);
        if (F::kArity == 2 && !y_scalar) memcpy(wb, &b[u], 16);
        else { wb[0] = wb[1] = wb[2] = wb[3] = (uint32_t)(uint8_t)ys * 0x01010101u; }
#pragma unroll
        for (int w = 0; w < 4; ++w) wr[w] = Word4<F>::op(wa[w], wb[w]);
        memcpy(&r, wr, 16);
      } else if constexpr (HasFast<F>::value && MODE == 1 && ScalarSplit<F>::value) {
        In av[E], bv[E];
        bool slow[E], any = false;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          av[e] = x_scalar ? xs : a[u].v[e];
          bv[e] = (F::kArity == 2 && !y_scalar) ? b[u].v[e] : ys;
          slow[e] = false;
          r.v[e] = apply_fast(f, av[e], bv[e], slow[e]);
          any |= slow[e];
        }
        if (any) {
#pragma unroll
          for (int e = 0; e < E; ++e)
            if (slow[e]) r.v[e] = apply_rare(f, av[e], bv[e]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e)
          r.v[e] = apply(f, x_scalar ? xs : a[u].v[e], (F::kArity == 2 && !y_scalar) ? b[u].v[e] : ys);
      }
      st_stream(op + base + (size_t)u * kThreads, r);
    }
  }
  // remainder packs and the element tail: the last CTA
  if (blockIdx.x == nfull) {
    stage_tables(f);
    for (size_t p = nfull * kTile + threadIdx.x; p < npack; p += kThreads) {
      PIn a, b;
      if (!x_scalar) a = ld_stream(xp + p);
      if (F::kArity == 2 && !y_scalar) b = ld_stream(yp + p);
      POut r;
#pragma unroll
      for (int e = 0; e < E; ++e)
        r.v[e] = apply(f, x_scalar ? xs : a.v[e], (F::kArity == 2 && !y_scalar) ? b.v[e] : ys);
      st_stream(op + p, r);
    }
    const size_t i = npack * E + threadIdx.x;
    if (i < n) {
      const In a = x_scalar ? xs : x[i];
      const In b = (F::kArity == 2 && !y_scalar) ? y[i] : ys;
      out[i] = apply(f, a, b);
    }
  }
}

// ------------------------------------------------------------------ flat kernel, broadcast operands
// OUT is one contiguous block of nrows x ncols elements (the usual case: :out preallocated, or a fresh
// result); each input is either laid out like OUT (kind 0), a ROW vector repeated down the rows (kind 1:
// index = column), a COLUMN vector repeated along the rows (kind 2: index = row * stride) or a scalar
// (kind 3).  Same one-tile-per-CTA streaming structure as k_map_flat -- the tile walks OUT's flat index and
// the broadcast operands are addressed by (row, column) recovered with ONE division per thread (ncols >=
// kThreads * E, so consecutive packs of a thread wrap at most one row).  This is the C3 shape of
// BASELINE.json: (4096 x 4096) * (4096) moved at 0.62 of the HBM peak through the row-looping k_map_rows
// (a 64-bit division chain per CTA, loads of row r+1 stuck behind the stores of row r) while the flat
// kernel moves the same bytes at 0.88 - 1.03 (profiles/time_c3_r02.txt).
template <class F, int UNROLL>
__global__ void __launch_bounds__(kThreads, MapMinBlocks<F>::value)
k_map_flat2(size_t n, int64_t ncols, const typename F::In *x, const typename F::In *y, typename F::Out *out,
            int xk, int yk, int64_t xrs, int64_t yrs, Scalar64 xv, Scalar64 yv, F f) {
  using In = typename F::In;
  using Out = typename F::Out;
  using P = PackOf<F>;
  constexpr int E = P::E;
  using PIn = typename P::PIn;
  using POut = typename P::POut;
  const size_t npack = n / E;                      // ncols % E == 0, so n % E == 0: no element tail
  constexpr size_t kTile = (size_t)kThreads * UNROLL;
  const size_t nfull = npack / kTile;
  POut *op = reinterpret_cast<POut *>(out);
  const In xs = scalar_as<In>(xv), ys = scalar_as<In>(yv);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch, as k_map_flat
  asm volatile("griddepcontrol.wait;" ::: "memory");

  auto load = [&](const In *ptr, int kind, int64_t rs, In sv, size_t pk, int64_t r, int64_t c) -> PIn {
    PIn v;
    if (kind == 0) return ld_stream(reinterpret_cast<const PIn *>(ptr) + pk);
    if (kind == 1) return ld_keep(reinterpret_cast<const PIn *>(ptr + c));
    const In s = (kind == 2) ? ptr[r * rs] : sv;
#pragma unroll
    for (int e = 0; e < E; ++e) v.v[e] = s;
    return v;
  };
  auto compute = [&](const PIn &a, const PIn &b) -> POut {
    POut r;
    if constexpr (Word4<F>::ok) {
      uint32_t wa[4], wb[4], wr[4];
      memcpy(wa, &a, 16);
      memcpy(wb, &b, 16);
#pragma unroll
      for (int w = 0; w < 4; ++w) wr[w] = Word4<F>::op(wa[w], wb[w]);
      memcpy(&r, wr, 16);
    // (a region functor's fast() is its GENERAL path and flags in-region arguments: without the tile vote of k_map_flat
    // the complete function, which branches to the region itself, is the better per-element form -- here and in k_map_rows)
    } else if constexpr (HasFast<F>::value && !RegionSplit<F>::value) {
      bool slow[E], any = false;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        slow[e] = false;
        r.v[e] = apply_fast(f, a.v[e], F::kArity == 2 ? b.v[e] : In(), slow[e]);
        any |= slow[e];
      }
      if (any) {
#pragma unroll
        for (int e = 0; e < E; ++e)
          if (slow[e]) r.v[e] = apply_rare(f, a.v[e], F::kArity == 2 ? b.v[e] : In());
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) r.v[e] = apply(f, a.v[e], F::kArity == 2 ? b.v[e] : In());
    }
    return r;
  };

  // A column-vector operand is ONE dependent load per pack from a vector nobody else keeps warm: with HBM busy
  // absorbing the write stream a cold line took microseconds to arrive and every CTA that needed it stalled
  // ((8192 x 1) + (1 x 8192): 54 us cold vs 42 us with the vector in L2, profiles/time_c3_r02.txt).  Each CTA
  // therefore asks L2 for the line it will NOT need for another ~2048 CTAs (and the first CTAs for the head).
  if (threadIdx.x == 0 && (xk == 2 || (F::kArity == 2 && yk == 2))) {
    const int64_t nrows = (int64_t)(n / (size_t)ncols);
    const int64_t ahead = (int64_t)((2048 * kTile * E) / (size_t)ncols) + 32;
    const int64_t r0 = (int64_t)(((size_t)blockIdx.x * kTile * E) / (size_t)ncols);
    const int64_t want[2] = {r0 + ahead, (int64_t)blockIdx.x};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (want[k] < nrows && (k == 0 || want[k] < ahead)) {
        if (xk == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + want[k] * xrs));
        if (F::kArity == 2 && yk == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(y + want[k] * yrs));
      }
    }
  }
  if (blockIdx.x < nfull) {
    const size_t base = (size_t)blockIdx.x * kTile + threadIdx.x;
    int64_t r = (int64_t)((base * E) / (size_t)ncols);
    int64_t c = (int64_t)(base * E) - r * ncols;
    PIn a[UNROLL], b[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      a[u] = load(x, xk, xrs, xs, base + (size_t)u * kThreads, r, c);
      if (F::kArity == 2) b[u] = load(y, yk, yrs, ys, base + (size_t)u * kThreads, r, c);
      c += (int64_t)kThreads * E;
      if (c >= ncols) { c -= ncols; ++r; }
    }
    stage_tables(f);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) st_stream(op + base + (size_t)u * kThreads, compute(a[u], b[u]));
  } else {
    stage_tables(f);
    for (size_t pk = nfull * kTile + threadIdx.x; pk < npack; pk += kThreads) {
      const int64_t r = (int64_t)((pk * E) / (size_t)ncols);
      const int64_t c = (int64_t)(pk * E) - r * ncols;
      const PIn a = load(x, xk, xrs, xs, pk, r, c);
      PIn b = a;
      if (F::kArity == 2) b = load(y, yk, yrs, ys, pk, r, c);
      st_stream(op + pk, compute(a, b));
    }
  }
}

// ------------------------------------------------------------------ rows kernel
struct RowsDesc {
  int64_t ncols;                 // inner (contiguous) extent
  int64_t nrows;                 // extent of the last outer axis (rows a CTA may group)
  int64_t row_stride[3];         // out, x, y stride of that axis (elements)
  int col_stride[3];             // 1 or 0 for x, y along the inner axis (out is always 1)
  int nouter;                    // remaining outer axes (rank - 2), slowest first
  int64_t odims[NUK_MAX_RANK];
  int64_t ostride[3][NUK_MAX_RANK];
  int rows_per_cta;
};

// the rows kernel takes the split too (tools/time_rows.py: (8192 x 8192) ^ (8192) f32 0.314 -> 0.244 ms, f64 0.499 ->
// 0.395 ms); a functor may override its unroll / register cap (kRowsUnroll, kRowsMinBlocks: atan2 f64 0.596 -> 0.458 ms)
template <class F, class = void> struct RowsTuning {   // default: 4 packs per thread, registers uncapped
  static constexpr int kUnroll = 4;
  static constexpr int kMinBlocks = 0;
};
template <class F> struct RowsTuning<F, std::void_t<decltype(F::kRowsUnroll)>> {
  static constexpr int kUnroll = F::kRowsUnroll;
  static constexpr int kMinBlocks = F::kRowsMinBlocks;
};
template <class F, int UNROLL>
__global__ void __launch_bounds__(kThreads, RowsTuning<F>::kMinBlocks)
k_map_rows(RowsDesc d, const typename F::In *x, const typename F::In *y, typename F::Out *out,
           Scalar64 xv, Scalar64 yv, F f) {
  using In = typename F::In;
  using Out = typename F::Out;
  using P = PackOf<F>;
  constexpr int E = P::E;
  using PIn = typename P::PIn;
  using POut = typename P::POut;

  stage_tables(f);
  // blockIdx.y -> column tile; blockIdx.x -> (outer index, row group)
  const int64_t row_groups = (d.nrows + d.rows_per_cta - 1) / d.rows_per_cta;
  int64_t g = blockIdx.x;
  const int64_t rg = g % row_groups;
  g /= row_groups;
  int64_t off_o = 0, off_x = 0, off_y = 0;
  for (int a = d.nouter - 1; a >= 0; --a) {
    const int64_t i = g % d.odims[a];
    g /= d.odims[a];
    off_o += i * d.ostride[0][a];
    off_x += i * d.ostride[1][a];
    off_y += i * d.ostride[2][a];
  }
  const int64_t r0 = rg * d.rows_per_cta;
  const int64_t r1 = (r0 + d.rows_per_cta < d.nrows) ? r0 + d.rows_per_cta : d.nrows;

  const int64_t c0 = ((int64_t)blockIdx.y * kThreads * UNROLL + threadIdx.x) * E;  // first column of pack u=0
  const bool x_rowinv = (d.row_stride[1] == 0), y_rowinv = (d.row_stride[2] == 0);
  const bool x_col = d.col_stride[1] != 0, y_col = (F::kArity == 2) && d.col_stride[2] != 0;

  const In *xb = x ? x + off_x : nullptr;
  const In *yb = y ? y + off_y : nullptr;
  Out *ob = out + off_o;

  // packs fully inside the row use the vector path; the ragged right edge goes scalar
  bool full[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) full[u] = c0 + (int64_t)u * kThreads * E + E <= d.ncols;

  PIn ha[UNROLL], hb[UNROLL];  // hoisted row-invariant operands
  if (x_col && x_rowinv) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (full[u]) ha[u] = ld_keep(reinterpret_cast<const PIn *>(xb + c0 + (int64_t)u * kThreads * E));
  }
  if (y_col && y_rowinv) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (full[u]) hb[u] = ld_keep(reinterpret_cast<const PIn *>(yb + c0 + (int64_t)u * kThreads * E));
  }

  for (int64_t r = r0; r < r1; ++r) {
    const In *xr = xb ? xb + r * d.row_stride[1] : nullptr;
    const In *yr = yb ? yb + r * d.row_stride[2] : nullptr;
    Out *orow = ob + r * d.row_stride[0];
    In xs = In(), ys = In();
    if (!x_col) xs = xr ? xr[0] : scalar_as<In>(xv);
    if (F::kArity == 2 && !y_col) ys = yr ? yr[0] : scalar_as<In>(yv);
    PIn a[UNROLL], b[UNROLL];
    if (x_col && !x_rowinv) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (full[u]) a[u] = ld_stream(reinterpret_cast<const PIn *>(xr + c0 + (int64_t)u * kThreads * E));
    }
    if (y_col && !y_rowinv) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (full[u]) b[u] = ld_stream(reinterpret_cast<const PIn *>(yr + c0 + (int64_t)u * kThreads * E));
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t c = c0 + (int64_t)u * kThreads * E;
      if (full[u]) {
        POut rr;
        if constexpr (HasFast<F>::value && !RegionSplit<F>::value && NUK_ROWS_FAST) {
          In av[E], bv[E];
          bool slow[E], any = false;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            av[e] = !x_col ? xs : (x_rowinv ? ha[u].v[e] : a[u].v[e]);
            bv[e] = (F::kArity == 2) ? (!y_col ? ys : (y_rowinv ? hb[u].v[e] : b[u].v[e])) : ys;
            slow[e] = false;
            rr.v[e] = apply_fast(f, av[e], bv[e], slow[e]);
            any |= slow[e];
          }
          if (any) {
#pragma unroll
            for (int e = 0; e < E; ++e)
              if (slow[e]) rr.v[e] = apply_rare(f, av[e], bv[e]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const In av = !x_col ? xs : (x_rowinv ? ha[u].v[e] : a[u].v[e]);
            const In bv = (F::kArity == 2) ? (!y_col ? ys : (y_rowinv ? hb[u].v[e] : b[u].v[e])) : ys;
            rr.v[e] = apply(f, av, bv);
          }
        }
        st_stream(reinterpret_cast<POut *>(orow + c), rr);
      } else {
        for (int e = 0; e < E && c + e < d.ncols; ++e) {
          const In av = !x_col ? xs : xr[c + e];
          const In bv = (F::kArity == 2 && y_col) ? yr[c + e] : ys;
          orow[c + e] = apply(f, av, bv);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ generic nd kernel
template <class F, int UNROLL>
__global__ void __launch_bounds__(kThreads)
k_map_nd(size_t n, NdDesc d, const typename F::In *x, const typename F::In *y,
         typename F::Out *out, Scalar64 xv, Scalar64 yv, F f) {
  using In = typename F::In;
  stage_tables(f);
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i0 = (size_t)blockIdx.x * kThreads + threadIdx.x; i0 < n; i0 += stride * UNROLL) {
    In a[UNROLL], b[UNROLL];
    int64_t oo[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      if (i < n) {
        int64_t ox = 0, oy = 0, o = 0;
        if (d.rank <= 1) {
          o = (int64_t)i * d.stride[0][0];
          ox = (int64_t)i * d.stride[1][0];
          oy = (int64_t)i * d.stride[2][0];
        } else {
          size_t rem = i;
          for (int ax = d.rank - 1; ax > 0; --ax) {
            const size_t q = rem / (size_t)d.dims[ax];
            const int64_t k = (int64_t)(rem - q * (size_t)d.dims[ax]);
            rem = q;
            o += k * d.stride[0][ax];
            ox += k * d.stride[1][ax];
            oy += k * d.stride[2][ax];
          }
          o += (int64_t)rem * d.stride[0][0];
          ox += (int64_t)rem * d.stride[1][0];
          oy += (int64_t)rem * d.stride[2][0];
        }
        oo[u] = o;
        a[u] = x ? x[ox] : scalar_as<In>(xv);
        if (F::kArity == 2) b[u] = y ? y[oy] : scalar_as<In>(yv);
        else b[u] = In();
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      if (i < n) out[oo[u]] = apply(f, a[u], b[u]);
    }
  }
}

// ------------------------------------------------------------------ launcher
// Problem statement shared by every map entry point (BMAS strided call == rank-1 problem).
struct MapProblem {
  int rank;                          // already coalesced over (out, x, y)
  int64_t dims[NUK_MAX_RANK];
  int64_t so[NUK_MAX_RANK], sx[NUK_MAX_RANK], sy[NUK_MAX_RANK];
  const void *x, *y;                 // nullptr => scalar by value (xv / yv)
  void *out;
  Scalar64 xv, yv;
  cudaStream_t stream;
};

int build_problem(MapProblem *p, int arity, int rank, const int64_t *dims, const int64_t *sx,
                  const int64_t *sy, const int64_t *so);

// Launch of the flat kernel with programmatic stream serialization: back-to-back small maps (C1: 10^6 doubles, ~3 us
// of work each) are separated by ~1.1 us of launch latency in stream order; with the attribute the next grid is
// scheduled while this one drains and blocks in griddepcontrol.wait until this one has completed -- the same
// ordering, without the bubble (profiles/c1_latency_r02.txt).  Option pdl = 0 turns it off.
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kernel)(KArgs...), const char *name, size_t grid, cudaStream_t stream, Args &&...args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = opt(OPT_PDL) != 0 ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
  if (e != cudaSuccess) {
    set_error(NUK_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e));
    cudaGetLastError();
    return NUK_ERR_CUDA;
  }
  return post_launch(name);
}
template <class F, int FU, int MODE>
int launch_flat(size_t grid, cudaStream_t stream, size_t n, const typename F::In *x, const typename F::In *y,
                typename F::Out *out, int xs, int ys, const Scalar64 &xv, const Scalar64 &yv, const F &f) {
  return launch_pdl(k_map_flat<F, FU, MODE>, "k_map_flat", grid, stream, n, x, y, out, xs, ys, xv, yv, f);
}

template <class F> int launch_map(const MapProblem &p, F f = F()) {
  using In = typename F::In;
  using Out = typename F::Out;
  using P = PackOf<F>;
  constexpr int E = P::E;
  constexpr int UNROLL = 4;
  const DeviceInfo *dev = device_info();
  if (!dev) return nuk_last_status();

  size_t n = 1;
  for (int a = 0; a < p.rank; ++a) n *= (size_t)p.dims[a];
  if (n == 0) return NUK_OK;
  const In *x = static_cast<const In *>(p.x);
  const In *y = static_cast<const In *>(p.y);
  Out *out = static_cast<Out *>(p.out);
  const bool binary = F::kArity == 2;
  const int last = p.rank - 1;

  auto vec_ok_in = [&](const In *ptr, int64_t inner_stride) {
    return ptr == nullptr || inner_stride == 0 || aligned_to(ptr, sizeof(typename P::PIn));
  };

  // ---- flat: rank <= 1, out stride 1, inputs stride 1 or 0
  if (p.rank <= 1) {
    const int64_t so = p.rank ? p.so[0] : 1, sx = p.rank ? p.sx[0] : 0, sy = p.rank ? p.sy[0] : 0;
    const bool xs = (x == nullptr) || sx == 0 || p.rank == 0;
    const bool ys = !binary || (y == nullptr) || sy == 0 || p.rank == 0;
    if (so == 1 && (xs || sx == 1) && (ys || sy == 1) && aligned_to(out, sizeof(typename P::POut)) &&
        (xs || vec_ok_in(x, 1)) && (ys || vec_ok_in(y, 1))) {
      constexpr int FU = MapUnroll<F>::value;    // per-functor unroll (profiles/membench_r01.txt)
      const size_t nfull = n / E / ((size_t)kThreads * FU);
      if (nfull + 1 > 0x7fffffffu) {
        set_error(NUK_ERR_INVALID, "array too large for one launch (%zu tiles)", nfull);
        return NUK_ERR_INVALID;
      }
      if constexpr ((HasFast<F>::value && F::kArity == 2) || F::kArity == 1) {
        if (xs || (binary && ys)) return launch_flat<F, FU, 1>(nfull + 1, p.stream, n, x, y, out, xs ? 1 : 0, ys ? 1 : 0, p.xv, p.yv, f);
        return launch_flat<F, FU, 0>(nfull + 1, p.stream, n, x, y, out, 0, 0, p.xv, p.yv, f);
      } else {
        return launch_flat<F, FU, 2>(nfull + 1, p.stream, n, x, y, out, xs ? 1 : 0, ys ? 1 : 0, p.xv, p.yv, f);
      }
    }
  }
  // ---- flat with broadcast operands: rank 2, OUT one contiguous block, inputs laid out like OUT / row vector /
  //      column vector / scalar (k_map_flat2)
  if (p.rank == 2 && p.so[1] == 1 && p.so[0] == p.dims[1] && p.dims[1] >= (int64_t)kThreads * E && p.dims[1] % E == 0 &&
      aligned_to(out, sizeof(typename P::POut)) && opt(OPT_ROWS_RPC) == 0) {
    const int64_t ncols = p.dims[1];
    auto kind_of = [&](const In *ptr, int64_t s0, int64_t s1) -> int {
      if (ptr == nullptr || (s0 == 0 && s1 == 0)) return 3;
      if (s1 == 1 && s0 == ncols) return aligned_to(ptr, sizeof(typename P::PIn)) ? 0 : -1;
      if (s1 == 1 && s0 == 0) return aligned_to(ptr, sizeof(typename P::PIn)) ? 1 : -1;
      if (s1 == 0) return 2;
      return -1;
    };
    const int xk = kind_of(x, p.sx[0], p.sx[1]);
    const int yk = binary ? kind_of(y, p.sy[0], p.sy[1]) : 3;
    constexpr int FU = MapUnroll<F>::value;
    const size_t nfull = n / E / ((size_t)kThreads * FU);
    if (xk >= 0 && yk >= 0 && nfull + 1 <= 0x7fffffffu) {
      Scalar64 xv = p.xv, yv = p.yv;
      // a (0, 0)-strided operand that lives in device memory is a one-element array: kind 2 with stride 0
      const int xk2 = (xk == 3 && x != nullptr) ? 2 : xk, yk2 = (yk == 3 && binary && y != nullptr) ? 2 : yk;
      return launch_pdl(k_map_flat2<F, FU>, "k_map_flat2", nfull + 1, p.stream, n, ncols, x, y, out, xk2, yk2,
                        (int64_t)(xk2 == 2 ? p.sx[0] : 0), (int64_t)(yk2 == 2 ? p.sy[0] : 0), xv, yv, f);
    }
  }
  // ---- rows: rank >= 2, inner axis contiguous for out, inputs 1/0 along it
  if (p.rank >= 2 && p.so[last] == 1 && (x == nullptr || p.sx[last] == 1 || p.sx[last] == 0) &&
      (!binary || y == nullptr || p.sy[last] == 1 || p.sy[last] == 0) && p.dims[last] >= 64) {
    // vector path needs every row start 16B-aligned for every vector operand
    bool ok = aligned_to(out, sizeof(typename P::POut));
    auto rows_aligned = [&](const int64_t *st, size_t elt) {
      for (int a = 0; a < last; ++a)
        if (((st[a] * (int64_t)elt) & 15) != 0) return false;
      return true;
    };
    ok = ok && rows_aligned(p.so, sizeof(Out));
    if (x && p.sx[last] == 1) ok = ok && aligned_to(x, sizeof(typename P::PIn)) && rows_aligned(p.sx, sizeof(In));
    if (binary && y && p.sy[last] == 1)
      ok = ok && aligned_to(y, sizeof(typename P::PIn)) && rows_aligned(p.sy, sizeof(In));
    size_t outer = 1;
    for (int a = 0; a < last - 1; ++a) outer *= (size_t)p.dims[a];
    if (ok) {
      RowsDesc d;
      memset(&d, 0, sizeof(d));
      d.ncols = p.dims[last];
      d.nrows = p.dims[last - 1];
      d.row_stride[0] = p.so[last - 1];
      d.row_stride[1] = x ? p.sx[last - 1] : 0;
      d.row_stride[2] = (binary && y) ? p.sy[last - 1] : 0;
      d.col_stride[0] = 1;
      d.col_stride[1] = (x && p.sx[last] == 1) ? 1 : 0;
      d.col_stride[2] = (binary && y && p.sy[last] == 1) ? 1 : 0;
      d.nouter = last - 1;
      for (int a = 0; a < d.nouter; ++a) {
        d.odims[a] = p.dims[a];
        d.ostride[0][a] = p.so[a];
        d.ostride[1][a] = x ? p.sx[a] : 0;
        d.ostride[2][a] = (binary && y) ? p.sy[a] : 0;
      }
      constexpr int RU = RowsTuning<F>::kUnroll;
      const int64_t col_tile = (int64_t)kThreads * RU * E;
      const int64_t col_tiles = (d.ncols + col_tile - 1) / col_tile;
      // Group rows so that the grid is a whole number of WAVES: `cap` CTAs are resident at once (occupancy x
      // SMs); a grid of 1.15 x cap runs as two waves and spends as long on the 15 % as on the first 100 %
      // (C3b at 4096^2 was 1366 CTAs on 1184 slots: 59 % of the HBM peak).  rows_per_cta = ceil(tiles / (k cap))
      // for the smallest k that keeps the group <= 32 rows: the last wave is as full as the others.
      static int occ_cached = 0;    // per kernel instantiation; the same on every B200
      if (occ_cached <= 0) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_map_rows<F, RU>, kThreads, 0) != cudaSuccess || occ <= 0) {
          cudaGetLastError();
          occ = 8;
        }
        occ_cached = occ;
      }
      const int64_t per_sm = opt(OPT_ROWS_CTAS_PER_SM) > 0 ? opt(OPT_ROWS_CTAS_PER_SM) : occ_cached;
      const int64_t cap = (int64_t)dev->sm_count * per_sm;
      const int64_t tiles = col_tiles * d.nrows * (int64_t)outer;
      int64_t rpc = 1;
      for (int64_t k = 1; tiles > cap; ++k) {
        rpc = (tiles + k * cap - 1) / (k * cap);
        if (rpc <= 32) break;
      }
      if (opt(OPT_ROWS_RPC) > 0) rpc = opt(OPT_ROWS_RPC);
      if (rpc > d.nrows) rpc = d.nrows;
      d.rows_per_cta = (int)rpc;
      const int64_t row_groups = (d.nrows + rpc - 1) / rpc;
      const size_t gy = (size_t)row_groups * outer;
      if (col_tiles <= 65535 && gy <= 0x7fffffffu) {
        dim3 grid((unsigned)gy, (unsigned)col_tiles);
        k_map_rows<F, RU><<<grid, kThreads, 0, p.stream>>>(d, x, y, out, p.xv, p.yv, f);
        return post_launch("k_map_rows");
      }
    }
  }
  // ---- generic
  {
    NdDesc d;
    memset(&d, 0, sizeof(d));
    d.rank = p.rank;
    for (int a = 0; a < p.rank; ++a) {
      d.dims[a] = p.dims[a];
      d.stride[0][a] = p.so[a];
      d.stride[1][a] = x ? p.sx[a] : 0;
      d.stride[2][a] = (binary && y) ? p.sy[a] : 0;
    }
    const size_t blocks = (n + (size_t)kThreads * UNROLL - 1) / ((size_t)kThreads * UNROLL);
    const int grid = grid_for(blocks, (int)opt(OPT_GRID_MULT));
    k_map_nd<F, UNROLL><<<grid, kThreads, 0, p.stream>>>(n, d, x, y, out, p.xv, p.yv, f);
    return post_launch("k_map_nd");
  }
}

}  // namespace nuk
