// trans_f32.cu -- single-float transcendental maps (nu:sin ... nu:exp nu:log nu:expt nu::atan2),
// table src/transcendental/package.lisp:63-84.  Kernels are the generic map kernels of map.cuh
// instantiated with the functors below; the math is in nukmath_f32.cuh / nukmath_f64.cuh.
#include "dispatch.cuh"
#include "nukmath_f32.cuh"
#include "nukmath_f64.cuh"
#include "tabstage.cuh"

namespace nuk {

#define NUK_F32_UNARY_U(Name, FN, UNR)                  \
  struct Name {                                         \
    using In = float; using Out = float;                \
    static constexpr int kArity = 1;                    \
    static constexpr int kMapUnroll = UNR;              \
    NUK_D float operator()(float a) const { return nukmath::FN(a); } \
  };
#define NUK_F32_UNARY(Name, FN) NUK_F32_UNARY_U(Name, FN, 4)
// with a branch-free main path FN_fast (map.cuh:HasFast); MINB caps registers (resident CTAs per SM)
#define NUK_F32_UNARY_FAST(Name, FN, UNR, MINB, ON)     \
  struct Name {                                         \
    using In = float; using Out = float;                \
    static constexpr int kArity = 1;                    \
    static constexpr int kMapUnroll = UNR;              \
    static constexpr int kMapMinBlocks = MINB;          \
    static constexpr bool kHasFast = (ON) == 1;         \
    static constexpr bool kWarpSplit = (ON) == 2;       /* 2: per-warp vote + called rare path (map.cuh) */ \
    NUK_D float operator()(float a) const { return nukmath::FN(a); } \
    NUK_D float fast(float a, bool &slow) const { return nukmath::FN##_fast(a, slow); } \
  };
#ifndef NUK_F32_MINB
#define NUK_F32_MINB 5
#endif
#ifndef NUK_FAST_TRIG32
#define NUK_FAST_TRIG32 2   // 0 per-element branches, 1 pack-level split (slower: profiles/fast_split_ab_r01.txt), 2 per-warp vote + called rare path (profiles/ab_warp_r02.txt: sin 0.955 -> 0.979, cos 0.962 -> 0.999 of the measured HBM peak)
#endif
#ifndef NUK_FAST_HYP32
#define NUK_FAST_HYP32 2    // sinh cosh: per-warp vote (profiles/ab_invtrig_r02.txt: sinh 0.783 pack split, 0.824 per-element branches, 0.859 warp vote)
#endif
#ifndef NUK_FAST_ATAN2F
#define NUK_FAST_ATAN2F 2   // 0 per-element branch (round 1: the pack split measured slower, profiles/fast_split_ab_r01.txt), 1 pack split, 2 per-warp vote
#endif
// measured on B200 (profiles/time_unroll_r01.txt, 2^28 elements): 4 packs per thread beat 1 by 15-30 % for
// every unary below (tan 5.2 vs 4.2 TB/s, cosh 5.2 vs 4.0), 2 packs are best for the binary ones
#ifndef NUK_F32_UNROLL
#define NUK_F32_UNROLL 4
#endif
#define NUK_F32_VIA_F64(Name, FN) NUK_F32_UNARY_U(Name, FN, NUK_F32_UNROLL)
#define NUK_F32_BINARY(Name, FN)                        \
  struct Name {                                         \
    using In = float; using Out = float;                \
    static constexpr int kArity = 2;                    \
    static constexpr int kMapUnroll = 2;                \
    NUK_D float operator()(float a, float b) const { return nukmath::FN(a, b); } \
  };
#ifndef NUK_FAST_TANH32
#define NUK_FAST_TANH32 2
#endif
#ifndef NUK_TANH32_MINB
#define NUK_TANH32_MINB 0
#endif
#ifndef NUK_TRIG32_MINB
#define NUK_TRIG32_MINB 0   // the 32-register cap of round 1 makes the warp-split form spill around the call
#endif
// packs per thread (profiles/ab_unroll_r02.txt): sin / cos 8 (32 elements per thread: 0.335 -> 0.312 ms, 0.336 -> 0.309 ms
// per 2^28 -- 105 / 106 % of the measured HBM peak), tanh 4 (8: 0.360 -> 0.368 ms)
#ifndef NUK_TRIG32_UNROLL
#define NUK_TRIG32_UNROLL 8
#endif
#ifndef NUK_TANH32_UNROLL
#define NUK_TANH32_UNROLL 4
#endif
#ifndef NUK_TAN32_UNROLL
#define NUK_TAN32_UNROLL 8
#endif
NUK_F32_UNARY_FAST(SinF, sin_f32, NUK_TRIG32_UNROLL, NUK_TRIG32_MINB, NUK_FAST_TRIG32) NUK_F32_UNARY_FAST(CosF, cos_f32, NUK_TRIG32_UNROLL, NUK_TRIG32_MINB, NUK_FAST_TRIG32)
NUK_F32_UNARY_FAST(TanhF, tanh_f32, NUK_TANH32_UNROLL, NUK_TANH32_MINB, NUK_FAST_TANH32)
NUK_F32_UNARY(LogF, log_f32)
NUK_F32_UNARY_U(ExpF, exp_f32, 2)   // 5.9 TB/s with 2 packs per thread vs 5.2 with 4 (profiles/membench_r01.txt)
#ifndef NUK_FAST_ATAN32
#define NUK_FAST_ATAN32 2   // asin acos atan: branch-free main path, per-warp vote, called rare path
#endif
#ifndef NUK_ATAN32_UNROLL
#define NUK_ATAN32_UNROLL 8 // profiles/ab_invtrig_r02.txt: asin 1.00 -> 1.06, acos 0.94 -> 0.99, atan 0.93 -> 0.98 of the measured peak with 8 packs per thread
#endif
NUK_F32_UNARY_FAST(TanF, tan_f32, NUK_TAN32_UNROLL, 0, NUK_FAST_TRIG32) NUK_F32_UNARY_FAST(AsinF, asin_f32, NUK_ATAN32_UNROLL, 0, NUK_FAST_ATAN32) NUK_F32_UNARY_FAST(AcosF, acos_f32, NUK_ATAN32_UNROLL, 0, NUK_FAST_ATAN32)
NUK_F32_UNARY_FAST(AtanF, atan_f32, NUK_ATAN32_UNROLL, 0, NUK_FAST_ATAN32) NUK_F32_UNARY_FAST(SinhF, sinh_f32, NUK_F32_UNROLL, 0, NUK_FAST_HYP32) NUK_F32_UNARY_FAST(CoshF, cosh_f32, NUK_F32_UNROLL, 0, NUK_FAST_HYP32)
// atanh, and asinh / acosh outside their polynomial regions: double arithmetic on the 768 bytes of single-float pow tables (nukmath_tab.cuh)
#ifndef NUK_HYPINV32_MINB
#define NUK_HYPINV32_MINB 0
#endif
#ifndef NUK_HYPINV32_UNROLL
#define NUK_HYPINV32_UNROLL 2
#endif
#define NUK_F32_UNARY_TAB(Name, FN, UNR, ON)            \
  struct Name : TabPowF {                               \
    using In = float; using Out = float;                \
    static constexpr int kArity = 1;                    \
    static constexpr int kMapUnroll = UNR;              \
    static constexpr int kMapMinBlocks = NUK_HYPINV32_MINB; \
    static constexpr bool kHasFast = (ON) == 1;         \
    static constexpr bool kWarpSplit = (ON) == 2;       \
    NUK_D float operator()(float a) const { return nukmath::FN(a, tab); } \
    NUK_D float fast(float a, bool &slow) const { return nukmath::FN##_fast(a, tab, slow); } \
  };
// asinh on |x| <= 1 and acosh on (1, 3] are single-float polynomials (nukmath_f32.cuh:asinh_small_f32 / acosh_near_f32:
// ~20 / ~31 instructions per element, HBM-bound) for a warp whose whole tile lies inside (map.cuh:kRegionSplit); every
// other warp runs the double-table logarithm path as the pack-level split it was before the regions existed
#ifndef NUK_ASINH32_UNROLL
#define NUK_ASINH32_UNROLL 4
#endif
#define NUK_F32_UNARY_REGION(Name, FN, POLY, INSIDE)    \
  struct Name : TabPowF {                               \
    using In = float; using Out = float;                \
    static constexpr int kArity = 1;                    \
    static constexpr int kMapUnroll = NUK_ASINH32_UNROLL; \
    static constexpr int kMapMinBlocks = NUK_HYPINV32_MINB; \
    static constexpr bool kHasFast = true;              \
    static constexpr bool kRegionSplit = true;          \
    NUK_D bool in_region(float a) const { return INSIDE; } \
    NUK_D float region(float a) const { return nukmath::POLY(a); } \
    NUK_D float operator()(float a) const { return nukmath::FN(a, tab); } \
    NUK_D float fast(float a, bool &slow) const { return nukmath::FN##_general(a, tab, slow); } \
  };
NUK_F32_UNARY_REGION(AsinhF, asinh_f32, asinh_small_f32, fabsf(a) <= 1.0f)
NUK_F32_UNARY_REGION(AcoshF, acosh_f32, acosh_near_f32, a > 1.0f && a <= 3.0f)
NUK_F32_UNARY_TAB(AtanhF, atanh_f32, NUK_HYPINV32_UNROLL, 1)
struct Atan2F {
  using In = float; using Out = float;
  static constexpr int kArity = 2;
  static constexpr int kMapUnroll = 2;
  static constexpr int kMapMinBlocks = 0;
  static constexpr bool kHasFast = NUK_FAST_ATAN2F == 1;
  static constexpr bool kWarpSplit = NUK_FAST_ATAN2F == 2;
  NUK_D float operator()(float a, float b) const { return nukmath::atan2_f32(a, b); }
  NUK_D float fast(float a, float b, bool &slow) const { return nukmath::atan2_f32_fast(a, b, slow); }
};
#ifndef NUK_POWF_MINB
#define NUK_POWF_MINB 5    // 48 registers: 0.59 ms per 2^28 (83 %); uncapped (88 registers, 2 CTAs/SM) 0.89 ms.  Re-checked after the table re-layout (profiles/ab_powf_r02.txt): 0.869 of the measured peak; caps of 4 / 6 0.794 / 0.743, 4 packs 0.738, 1 pack 0.625
#endif
#ifndef NUK_POWF_UNROLL
#define NUK_POWF_UNROLL 2
#endif
struct PowF : TabPowF {   // 768 bytes of tables in shared memory (nukmath_tab.cuh:pow_f32)
  using In = float; using Out = float;
  static constexpr int kArity = 2;
  static constexpr int kMapUnroll = NUK_POWF_UNROLL;
  static constexpr int kMapMinBlocks = NUK_POWF_MINB;
  static constexpr bool kHasFast = true;
#ifndef NUK_POWF_SCALAR_SPLIT
#define NUK_POWF_SCALAR_SPLIT false
#endif
#ifndef NUK_POWF_SCALAR_MINB
#define NUK_POWF_SCALAR_MINB 0   // uncapped, per-element: 0.68 ms; split at 64 / 80 registers 0.71 / 0.79 ms
#endif
  static constexpr bool kScalarSplit = NUK_POWF_SCALAR_SPLIT;   // nu:expt x 2.5: 0.70 ms per 2^28 per-element, 0.98 ms split at 48 registers (spills)
  static constexpr int kScalarMinBlocks = NUK_POWF_SCALAR_MINB;
  NUK_D float operator()(float a, float b) const { return nukmath::pow_f32(a, b, tab); }
  NUK_D float fast(float a, float b, bool &slow) const { return nukmath::pow_f32_fast(a, b, tab, slow); }
};

// nu:log x y = log(x) / log(y) in one pass; IEEE division of the two rounded logs, so the bits are those of
// the reference's log, log, divide sequence (src/transcendental/two-arg-fn-float.lisp:531-547)
struct LogbF {
  using In = float; using Out = float;
  static constexpr int kArity = 2;
  static constexpr int kMapUnroll = 2;
  static constexpr bool kHasFast = true;
  NUK_D float operator()(float a, float b) const { return nukmath::log_f32(a) / nukmath::log_f32(b); }
  NUK_D float fast(float a, float b, bool &slow) const {
    bool sa, sb;
    const float la = nukmath::log_f32_fast(a, sa), lb = nukmath::log_f32_fast(b, sb);
    slow = sa | sb;
    return la / lb;
  }
};

int launch_trans_f32(int op, int arity, const MapProblem &p) {
  if (arity == 2) {
    if (op == NUK_LOGB) return launch_map<LogbF>(p);
    if (op == NUK_POW) return launch_map<PowF>(p);
    if (op == NUK_ATAN2) return launch_map<Atan2F>(p);
  } else {
    switch (op) {
      case NUK_SIN: return launch_map<SinF>(p);
      case NUK_COS: return launch_map<CosF>(p);
      case NUK_TAN: return launch_map<TanF>(p);
      case NUK_ASIN: return launch_map<AsinF>(p);
      case NUK_ACOS: return launch_map<AcosF>(p);
      case NUK_ATAN: return launch_map<AtanF>(p);
      case NUK_SINH: return launch_map<SinhF>(p);
      case NUK_COSH: return launch_map<CoshF>(p);
      case NUK_TANH: return launch_map<TanhF>(p);
      case NUK_ASINH: return launch_map<AsinhF>(p);
      case NUK_ACOSH: return launch_map<AcoshF>(p);
      case NUK_ATANH: return launch_map<AtanhF>(p);
      case NUK_EXP: return launch_map<ExpF>(p);
      case NUK_LOG: return launch_map<LogF>(p);
    }
  }
  set_error(NUK_ERR_UNSUPPORTED, "no single-float transcendental kernel for op %d (arity %d)", op, arity);
  return NUK_ERR_UNSUPPORTED;
}

}  // namespace nuk

// ---- self-test of the seed primitives: rcp_normal / sqrt_normal (MUFU seed + FMA step, no range
// guard) against the IEEE instructions rcp.rn.f32 / sqrt.rn.f32 for every bit pattern in [lo, hi).
// The host compile of the math headers uses 1.0f/d and sqrtf, so "device seed == IEEE" is what
// makes the host ULP sweeps carry over to the device (tests/test_transcendental_gpu.py).
namespace nuk {
__global__ void k_selftest_seeds(uint32_t lo, uint64_t count, unsigned long long *mismatches) {
  unsigned long long bad = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
    const float x = __uint_as_float(lo + (uint32_t)i);
    if (__float_as_uint(nukmath::rcp_normal(x)) != __float_as_uint(__frcp_rn(x))) ++bad;
    if (__float_as_uint(nukmath::sqrt_normal(x)) != __float_as_uint(__fsqrt_rn(x))) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}
}  // namespace nuk

NUK_API int nuk_selftest_seeds(uint32_t lo_bits, uint32_t hi_bits, unsigned long long *mismatches, nuk_stream_t s) {
  using namespace nuk;
  if (!mismatches || hi_bits < lo_bits) {
    set_error(NUK_ERR_INVALID, "nuk_selftest_seeds: bad arguments");
    return NUK_ERR_INVALID;
  }
  if (!device_info()) return nuk_last_status();
  void *d = nullptr;
  NUK_TRY(scratch((cudaStream_t)s, sizeof(unsigned long long), &d));
  NUK_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), (cudaStream_t)s));
  k_selftest_seeds<<<148 * 8, 256, 0, (cudaStream_t)s>>>(lo_bits, (uint64_t)hi_bits - lo_bits, (unsigned long long *)d);
  NUK_CUDA(cudaGetLastError());
  count_launch();
  NUK_CUDA(cudaMemcpyAsync(mismatches, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, (cudaStream_t)s));
  NUK_CUDA(cudaStreamSynchronize((cudaStream_t)s));
  return NUK_OK;
}
