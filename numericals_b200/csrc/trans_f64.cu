// trans_f64.cu -- double-float transcendental maps, table src/transcendental/package.lisp:63-84.
#include "dispatch.cuh"
#include "nukmath_f64.cuh"
#include "tabstage.cuh"

namespace nuk {

#define NUK_F64_UNARY(Name, FN)                         \
  struct Name {                                         \
    using In = double; using Out = double;              \
    static constexpr int kArity = 1;                    \
    static constexpr int kMapUnroll = 2;                \
    NUK_D double operator()(double a) const { return nukmath::FN(a); } \
  };
#define NUK_F64_BINARY(Name, FN)                        \
  struct Name {                                         \
    using In = double; using Out = double;              \
    static constexpr int kArity = 2;                    \
    static constexpr int kMapUnroll = 1;                \
    NUK_D double operator()(double a, double b) const { return nukmath::FN(a, b); } \
  };
#ifndef NUK_F64_UNROLL
#define NUK_F64_UNROLL 2
#endif
#define NUK_F64_UNARY_FAST(Name, FN, MINB, ON, UNR, BATCH) \
  struct Name {                                         \
    using In = double; using Out = double;              \
    static constexpr int kArity = 1;                    \
    static constexpr int kMapUnroll = UNR;              \
    static constexpr int kMapMinBlocks = MINB;          \
    static constexpr bool kHasFast = ON;                \
    static constexpr bool kFastBatch = (ON) && (BATCH); /* whole-tile batch (map.cuh:kFastBatch) */ \
    NUK_D double operator()(double a) const { return nukmath::FN(a); } \
    NUK_D double fast(double a, bool &slow) const { return nukmath::FN##_fast(a, slow); } \
  };
#ifndef NUK_FAST_TRIG64
#define NUK_FAST_TRIG64 true
#endif
#ifndef NUK_FAST_INV64
#define NUK_FAST_INV64 true
#endif
#ifndef NUK_FAST_HYP64
#define NUK_FAST_HYP64 true
#endif
// which functors take the split, and their register caps, were chosen by same-box A/B runs
// (profiles/fast_split_ab_r01.txt): atan2 +16, atan +5, asin/acos/tan/cosh +1 points of HBM peak; sin, cos, sinh
// and the single-float trig/atan2 kernels were faster without it
#ifndef NUK_TRIG_MINB
#define NUK_TRIG_MINB 4
#endif
#ifndef NUK_SINCOS64_UNROLL
#define NUK_SINCOS64_UNROLL 4
#endif
// packs per thread / register caps from same-box A/B runs (profiles/ab_f64b_r02.txt, ab_f64c_r02.txt): atan 4 packs
// uncapped 0.683 -> 0.718, tan 4 packs 0.540 -> 0.580, asin / acos 3 packs 0.745 -> 0.758 / 0.744 -> 0.770,
// atan2 2 packs at 64 registers 0.829 -> 0.934 of the measured peak
#ifndef NUK_ATAN64_UNROLL
#define NUK_ATAN64_UNROLL 4
#endif
#ifndef NUK_ATAN64_BATCH
#define NUK_ATAN64_BATCH 0
#endif
#ifndef NUK_ATAN64_MINB
#define NUK_ATAN64_MINB 0
#endif
#ifndef NUK_TAN64_UNROLL
#define NUK_TAN64_UNROLL 4
#endif
#ifndef NUK_TAN64_MINB
#define NUK_TAN64_MINB 5   // profiles/ab_tan64_r02.txt: 0.669 of the measured peak; caps of 3 / 4 / 6 / none 0.676 / 0.664 / 0.610 / 0.658
#endif
#ifndef NUK_ASIN64_UNROLL
#define NUK_ASIN64_UNROLL 3
#endif
#ifndef NUK_ASIN64_MINB
#define NUK_ASIN64_MINB 8   // 32 registers: asin 0.834 -> 0.871 of the measured peak; acos is best uncapped (0.868; 0.839 at 8) -- profiles/ab_asin64_r02.txt
#endif
#ifndef NUK_ACOS64_MINB
#define NUK_ACOS64_MINB 0
#endif
// cos capped at 32 registers: 74 -> 86 % (profiles/minblocks_sweep_r01.txt)
#ifndef NUK_SINCOS64_FAST
#define NUK_SINCOS64_FAST false   // the pack-level split is still slower at 4 packs per thread (profiles/ab_sincos64_r02.txt: sin 0.840 -> 0.831, cos 0.879 -> 0.817; register caps of 5 / 6 change nothing or lose)
#endif
#ifndef NUK_SIN64_MINB
#define NUK_SIN64_MINB 0
#endif
#ifndef NUK_COS64_MINB
#define NUK_COS64_MINB 8
#endif
NUK_F64_UNARY_FAST(SinD, sin_f64, NUK_SIN64_MINB, NUK_SINCOS64_FAST, NUK_SINCOS64_UNROLL, 0)   // 4 packs per thread: sin 0.811 -> 0.840, cos 0.863 -> 0.879 (profiles/ab_f64b_r02.txt)
 NUK_F64_UNARY_FAST(CosD, cos_f64, NUK_COS64_MINB, NUK_SINCOS64_FAST, NUK_SINCOS64_UNROLL, 0)
NUK_F64_UNARY_FAST(TanD, tan_f64, NUK_TAN64_MINB, NUK_FAST_TRIG64, NUK_TAN64_UNROLL, 0)
NUK_F64_UNARY_FAST(AsinD, asin_f64, NUK_ASIN64_MINB, NUK_FAST_INV64, NUK_ASIN64_UNROLL, 0) NUK_F64_UNARY_FAST(AcosD, acos_f64, NUK_ACOS64_MINB, NUK_FAST_INV64, NUK_ASIN64_UNROLL, 0)
NUK_F64_UNARY_FAST(AtanD, atan_f64, NUK_ATAN64_MINB, NUK_FAST_INV64, NUK_ATAN64_UNROLL, NUK_ATAN64_BATCH)
#ifndef NUK_ATAN2_64_UNROLL
#define NUK_ATAN2_64_UNROLL 2
#endif
#ifndef NUK_ATAN2_64_MINB
#define NUK_ATAN2_64_MINB 4
#endif
struct Atan2D {
  using In = double; using Out = double;
  static constexpr int kArity = 2;
  static constexpr int kMapUnroll = NUK_ATAN2_64_UNROLL;
  static constexpr int kMapMinBlocks = NUK_ATAN2_64_MINB;
  static constexpr int kRowsUnroll = 1, kRowsMinBlocks = 5;
  static constexpr bool kHasFast = NUK_FAST_INV64;
  NUK_D double operator()(double a, double b) const { return nukmath::atan2_f64(a, b); }
  NUK_D double fast(double a, double b, bool &slow) const { return nukmath::atan2_f64_fast(a, b, slow); }
};
// table-driven functors (tabstage.cuh)
#ifndef NUK_FAST_SPLIT
#define NUK_FAST_SPLIT true
#endif
#ifndef NUK_POW_UNROLL
#define NUK_POW_UNROLL 4
#endif
#ifndef NUK_POW_MINB
#define NUK_POW_MINB 3     // 80 registers: 1.13 ms per 2^28 (87 %); 64 registers spill (1.46 ms), uncapped 90 regs 1.26 ms
#endif
#ifndef NUK_TANH_MINB
#define NUK_TANH_MINB 4    // 64 registers: 0.73 ms per 2^28 (90 %)
#endif
struct PowD : TabAll {
  using In = double; using Out = double;
  static constexpr int kArity = 2;
  static constexpr int kMapUnroll = NUK_POW_UNROLL;
  static constexpr int kMapMinBlocks = NUK_POW_MINB;
  static constexpr bool kHasFast = NUK_FAST_SPLIT;
  NUK_D double operator()(double a, double b) const { return nukmath::pow_f64(a, b, tab); }
  NUK_D double fast(double a, double b, bool &slow) const { return nukmath::pow_f64_fast(a, b, tab, slow); }
};
#define NUK_F64_TAB_UNARY(Name, FN, Stage, MINB)        \
  struct Name : Stage {                                 \
    using In = double; using Out = double;              \
    static constexpr int kArity = 1;                    \
    static constexpr int kMapUnroll = 4;                \
    static constexpr int kMapMinBlocks = MINB;          \
    NUK_D double operator()(double a) const { return nukmath::FN(a, tab); } \
  };
NUK_F64_TAB_UNARY(ExpD, exp_f64, TabExp, 6) NUK_F64_TAB_UNARY(LogD, log_f64, TabLog, 8)
#define NUK_F64_TAB_UNARY_FAST(Name, FN, Stage, MINB)    \
  struct Name : Stage {                                 \
    using In = double; using Out = double;              \
    static constexpr int kArity = 1;                    \
    static constexpr int kMapUnroll = 4;                \
    static constexpr int kMapMinBlocks = NUK_FAST_HYP64 ? MINB : 0; \
    static constexpr bool kHasFast = NUK_FAST_HYP64;    \
    NUK_D double operator()(double a) const { return nukmath::FN(a, tab); } \
    NUK_D double fast(double a, bool &slow) const { return nukmath::FN##_fast(a, tab, slow); } \
  };
#ifndef NUK_FAST_IHYP64
#define NUK_FAST_IHYP64 true
#endif
#ifndef NUK_IHYP_MINB
#define NUK_IHYP_MINB 4
#endif
#define NUK_F64_TAB_UNARY_FAST2(Name, FN, Stage, MINB, ON) \
  struct Name : Stage {                                 \
    using In = double; using Out = double;              \
    static constexpr int kArity = 1;                    \
    static constexpr int kMapUnroll = 4;                \
    static constexpr int kMapMinBlocks = (ON) ? MINB : 0; \
    static constexpr bool kHasFast = ON;                \
    NUK_D double operator()(double a) const { return nukmath::FN(a, tab); } \
    NUK_D double fast(double a, bool &slow) const { return nukmath::FN##_fast(a, tab, slow); } \
  };
// asinh on |x| <= 1 and acosh on (1, 3] are polynomials (nukmath_tab.cuh:asinh_small_f64 / acosh_near_f64) for a warp whose
// whole tile lies inside (map.cuh:kRegionSplit); every other warp runs the logarithm path as the pack-level split it was
#ifndef NUK_IHYP64_UNROLL
#define NUK_IHYP64_UNROLL 4
#endif
#define NUK_F64_TAB_UNARY_REGION(Name, FN, POLY, INSIDE) \
  struct Name : TabLog {                                \
    using In = double; using Out = double;              \
    static constexpr int kArity = 1;                    \
    static constexpr int kMapUnroll = NUK_IHYP64_UNROLL; \
    static constexpr int kMapMinBlocks = NUK_IHYP_MINB; \
    static constexpr bool kHasFast = true;              \
    static constexpr bool kRegionSplit = true;          \
    NUK_D bool in_region(double a) const { return INSIDE; } \
    NUK_D double region(double a) const { return nukmath::POLY(a); } \
    NUK_D double operator()(double a) const { return nukmath::FN(a, tab); } \
    NUK_D double fast(double a, bool &slow) const { return nukmath::FN##_general(a, tab, slow); } \
  };
NUK_F64_TAB_UNARY_REGION(AsinhD, asinh_f64, asinh_small_f64, fabs(a) <= 1.0)
NUK_F64_TAB_UNARY_REGION(AcoshD, acosh_f64, acosh_near_f64, a > 1.0 && a <= 3.0)
NUK_F64_TAB_UNARY_FAST2(AtanhD, atanh_f64, TabLog, NUK_IHYP_MINB, NUK_FAST_IHYP64)
// sinh / cosh: ~50-64 instructions per element and two per-lane indexed table reads each (the 32-entry table: 4
// shared-memory wavefronts per read)
#ifndef NUK_HYP64_BATCH
#define NUK_HYP64_BATCH 0   // profiles/ab_f64_r02.txt: the whole-tile batch is slower here (0.63 vs 0.73 of the measured peak)
#endif
#ifndef NUK_HYP64_MINB
#define NUK_HYP64_MINB 0
#endif
#ifndef NUK_HYP64_UNROLL
#define NUK_HYP64_UNROLL 4
#endif
#define NUK_F64_TAB_UNARY_BATCH(Name, FN, Stage)        \
  struct Name : Stage {                                 \
    using In = double; using Out = double;              \
    static constexpr int kArity = 1;                    \
    static constexpr int kMapUnroll = NUK_HYP64_UNROLL; \
    static constexpr int kMapMinBlocks = NUK_HYP64_MINB; \
    static constexpr bool kHasFast = true;              \
    static constexpr bool kFastBatch = NUK_HYP64_BATCH; \
    NUK_D double operator()(double a) const { return nukmath::FN(a, tab); } \
    NUK_D double fast(double a, bool &slow) const { return nukmath::FN##_fast(a, tab, slow); } \
  };
NUK_F64_TAB_UNARY_BATCH(SinhD, sinh_f64, TabExp32) NUK_F64_TAB_UNARY_BATCH(CoshD, cosh_f64, TabExp32)
struct TanhD : TabExp {
  using In = double; using Out = double;
  static constexpr int kArity = 1;
  static constexpr int kMapUnroll = 4;
  static constexpr int kMapMinBlocks = NUK_TANH_MINB;
  static constexpr bool kHasFast = NUK_FAST_SPLIT;
  NUK_D double operator()(double a) const { return nukmath::tanh_f64(a, tab); }
  NUK_D double fast(double a, bool &slow) const { return nukmath::tanh_f64_fast(a, tab, slow); }
};

// nu:log x y in one pass (see LogbF); both logs from the shared-memory table
struct LogbD : TabLog {
  using In = double; using Out = double;
  static constexpr int kArity = 2;
  static constexpr int kMapUnroll = 2;
  static constexpr bool kHasFast = true;
  NUK_D double operator()(double a, double b) const { return nukmath::log_f64(a, tab) / nukmath::log_f64(b, tab); }
  NUK_D double fast(double a, double b, bool &slow) const {
    bool sa, sb;
    const double la = nukmath::log_f64_fast(a, tab, sa), lb = nukmath::log_f64_fast(b, tab, sb);
    slow = sa | sb;
    return la / lb;
  }
};

int launch_trans_f64(int op, int arity, const MapProblem &p) {
  if (arity == 2) {
    if (op == NUK_LOGB) return launch_map<LogbD>(p);
    if (op == NUK_POW) return launch_map<PowD>(p);
    if (op == NUK_ATAN2) return launch_map<Atan2D>(p);
  } else {
    switch (op) {
      case NUK_SIN: return launch_map<SinD>(p);
      case NUK_COS: return launch_map<CosD>(p);
      case NUK_TAN: return launch_map<TanD>(p);
      case NUK_ASIN: return launch_map<AsinD>(p);
      case NUK_ACOS: return launch_map<AcosD>(p);
      case NUK_ATAN: return launch_map<AtanD>(p);
      case NUK_SINH: return launch_map<SinhD>(p);
      case NUK_COSH: return launch_map<CoshD>(p);
      case NUK_TANH: return launch_map<TanhD>(p);
      case NUK_ASINH: return launch_map<AsinhD>(p);
      case NUK_ACOSH: return launch_map<AcoshD>(p);
      case NUK_ATANH: return launch_map<AtanhD>(p);
      case NUK_EXP: return launch_map<ExpD>(p);
      case NUK_LOG: return launch_map<LogD>(p);
    }
  }
  set_error(NUK_ERR_UNSUPPORTED, "no double-float transcendental kernel for op %d (arity %d)", op, arity);
  return NUK_ERR_UNSUPPORTED;
}

}  // namespace nuk
