// ops.cuh -- elementwise functors (arithmetic, min/max, compare, abs/rounding, casts, bitwise).
// Compiled with -fmad=false: a*b+c chains are never contracted, so add/sub/mul/div are
// bit-identical to the CPU (IEEE-754 RN; nvcc's default -prec-div=true -prec-sqrt=true, no FTZ).
#pragma once
#include "common.cuh"

namespace nuk {

template <typename T> struct UnsignedOf { using type = typename std::make_unsigned<T>::type; };

// integer arithmetic wraps: done in the unsigned type (>= 32 bit) of the same width
template <typename T, bool IsFloat = std::is_floating_point<T>::value> struct WrapT {
  using type = T;  // floats accumulate in their own type
};
template <typename T> struct WrapT<T, false> {
  using U = typename std::make_unsigned<T>::type;
  using type = typename std::conditional<(sizeof(T) < 4), uint32_t, U>::type;
};

#define NUK_BINARY_FUNCTOR(Name, EXPR_FLOAT, EXPR_INT)                              \
  template <typename T> struct Name {                                               \
    using In = T;                                                                   \
    using Out = T;                                                                  \
    static constexpr int kArity = 2;                                                \
    NUK_D T operator()(T a, T b) const {                                            \
      if constexpr (std::is_floating_point<T>::value) { return EXPR_FLOAT; }        \
      else {                                                                        \
        using W = typename WrapT<T>::type;                                          \
        const W ua = (W)a, ub = (W)b;                                               \
        (void)ua; (void)ub;                                                         \
        return EXPR_INT;                                                            \
      }                                                                             \
    }                                                                               \
  };
NUK_BINARY_FUNCTOR(AddOp, a + b, (T)(ua + ub))
NUK_BINARY_FUNCTOR(SubOp, a - b, (T)(ua - ub))
NUK_BINARY_FUNCTOR(MulOp, a * b, (T)(ua * ub))
// x86 maxps/minps semantics: second operand when unordered or equal
NUK_BINARY_FUNCTOR(MaxOp, (a > b) ? a : b, (a > b) ? a : b)
NUK_BINARY_FUNCTOR(MinOp, (a < b) ? a : b, (a < b) ? a : b)

template <typename T> struct DivOp {
  using In = T; using Out = T;
  static constexpr int kArity = 2;
  NUK_D T operator()(T a, T b) const { return a / b; }  // IEEE division (div.rn)
};

#define NUK_CMP_FUNCTOR(Name, OP)                                  \
  template <typename T> struct Name {                              \
    using In = T; using Out = uint8_t;                             \
    static constexpr int kArity = 2;                               \
    NUK_D uint8_t operator()(T a, T b) const { return (a OP b) ? 1 : 0; } \
  };
NUK_CMP_FUNCTOR(LtOp, <) NUK_CMP_FUNCTOR(LeOp, <=) NUK_CMP_FUNCTOR(EqOp, ==)
NUK_CMP_FUNCTOR(NeqOp, !=) NUK_CMP_FUNCTOR(GeOp, >=) NUK_CMP_FUNCTOR(GtOp, >)

// ---- 8-bit SIMD-in-a-word forms (four elements per instruction group): without them the byte
// ops are issue-bound at ~45% of the HBM roofline (profiles/sweep_r01_a.jsonl)
#define NUK_WORD4(FUNCTOR, EXPR)                                                 \
  template <> struct Word4<FUNCTOR> {                                            \
    static constexpr bool ok = true;                                             \
    NUK_D static uint32_t op(uint32_t a, uint32_t b) { return (EXPR); }          \
  };
NUK_WORD4(AddOp<int8_t>, __vadd4(a, b))
NUK_WORD4(SubOp<int8_t>, __vsub4(a, b))
NUK_WORD4(MaxOp<int8_t>, __vmaxs4(a, b))
NUK_WORD4(MinOp<int8_t>, __vmins4(a, b))
NUK_WORD4(MaxOp<uint8_t>, __vmaxu4(a, b))
NUK_WORD4(MinOp<uint8_t>, __vminu4(a, b))
NUK_WORD4(LtOp<int8_t>, __vcmplts4(a, b) & 0x01010101u)
NUK_WORD4(LeOp<int8_t>, __vcmples4(a, b) & 0x01010101u)
NUK_WORD4(GeOp<int8_t>, __vcmpges4(a, b) & 0x01010101u)
NUK_WORD4(GtOp<int8_t>, __vcmpgts4(a, b) & 0x01010101u)
NUK_WORD4(LtOp<uint8_t>, __vcmpltu4(a, b) & 0x01010101u)
NUK_WORD4(LeOp<uint8_t>, __vcmpleu4(a, b) & 0x01010101u)
NUK_WORD4(GeOp<uint8_t>, __vcmpgeu4(a, b) & 0x01010101u)
NUK_WORD4(GtOp<uint8_t>, __vcmpgtu4(a, b) & 0x01010101u)
NUK_WORD4(EqOp<uint8_t>, __vcmpeq4(a, b) & 0x01010101u)
NUK_WORD4(NeqOp<uint8_t>, __vcmpne4(a, b) & 0x01010101u)

template <typename T> struct AbsOp {
  using In = T; using Out = T;
  static constexpr int kArity = 1;
  NUK_D T operator()(T a) const {
    if constexpr (std::is_same<T, float>::value) return fabsf(a);
    else if constexpr (std::is_same<T, double>::value) return fabs(a);
    else if constexpr (std::is_unsigned<T>::value) return a;
    else {
      using W = typename WrapT<T>::type;
      return (T)(a < 0 ? (W)0 - (W)a : (W)a);
    }
  }
};
#define NUK_ROUND_FUNCTOR(Name, FN32, FN64)                                   \
  template <typename T> struct Name {                                         \
    using In = T; using Out = T;                                              \
    static constexpr int kArity = 1;                                          \
    NUK_D T operator()(T a) const {                                           \
      if constexpr (std::is_same<T, float>::value) return FN32(a);            \
      else return FN64(a);                                                    \
    }                                                                         \
  };
NUK_ROUND_FUNCTOR(RoundOp, rintf, rint)   // half-even = CL fround
NUK_ROUND_FUNCTOR(TruncOp, truncf, trunc)
NUK_ROUND_FUNCTOR(FloorOp, floorf, floor)
NUK_ROUND_FUNCTOR(CeilOp, ceilf, ceil)
NUK_ROUND_FUNCTOR(SqrtOp, sqrtf, sqrt)    // sqrt.rn: correctly rounded

template <typename TI, typename TO> struct CastOp {
  using In = TI; using Out = TO;
  static constexpr int kArity = 1;
  NUK_D TO operator()(TI a) const { return (TO)a; }  // cvt.rn for int->float, float->float
};

// bitwise ops run on machine words; BMAS counts bytes (two-arg-fn-logical.lisp:98)
#define NUK_BIT_FUNCTOR(Name, EXPR)                           \
  template <typename T> struct Name {                         \
    using In = T; using Out = T;                              \
    static constexpr int kArity = 2;                          \
    NUK_D T operator()(T a, T b) const { return (T)(EXPR); }  \
  };
NUK_BIT_FUNCTOR(AndOp, a & b) NUK_BIT_FUNCTOR(OrOp, a | b) NUK_BIT_FUNCTOR(XorOp, a ^ b)
NUK_BIT_FUNCTOR(AndNotOp, (~a) & b)
template <typename T> struct NotOp {
  using In = T; using Out = T;
  static constexpr int kArity = 1;
  NUK_D T operator()(T a) const { return (T)~a; }
};
struct SraOp {
  using In = int64_t; using Out = int64_t;
  static constexpr int kArity = 2;
  NUK_D int64_t operator()(int64_t a, int64_t b) const { return a >> (b & 63); }
};

}  // namespace nuk
