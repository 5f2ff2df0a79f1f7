// unary.cu -- nu:abs ffloor fceiling fround ftruncate sqrt copy lognot + the casts behind
// nu:copy / nu:astype.  Table: src/basic-math/package.lisp:112-129,175-177.
#include "dispatch.cuh"
#include "ops.cuh"

namespace nuk {

template <typename T> struct CopyOp {
  using In = T; using Out = T;
  static constexpr int kArity = 1;
  NUK_D T operator()(T a) const { return a; }
};

template <template <class> class Op> static int floats(int op, int dtype, const MapProblem &p) {
  if (dtype == NUK_F32) return launch_map<Op<float>>(p);
  if (dtype == NUK_F64) return launch_map<Op<double>>(p);
  set_error(NUK_ERR_UNSUPPORTED, "unary op %d: no CFUNCTION for non-float dtype %d", op, dtype);
  return NUK_ERR_UNSUPPORTED;
}
template <template <class> class Op> static int by_width(int dtype, const MapProblem &p) {
  switch (dtype_size(dtype)) {
    case 8: return launch_map<Op<uint64_t>>(p);
    case 4: return launch_map<Op<uint32_t>>(p);
    case 2: return launch_map<Op<uint16_t>>(p);
    case 1: return launch_map<Op<uint8_t>>(p);
  }
  set_error(NUK_ERR_UNSUPPORTED, "no kernel for dtype %d", dtype);
  return NUK_ERR_UNSUPPORTED;
}

int launch_unary(int op, int dtype, const MapProblem &p) {
  switch (op) {
    case NUK_COPY: return by_width<CopyOp>(dtype, p);
    case NUK_NOT:
      if (dtype == NUK_F32 || dtype == NUK_F64) break;
      return by_width<NotOp>(dtype, p);
    case NUK_ABS:
      switch (dtype) {
        case NUK_F32: return launch_map<AbsOp<float>>(p);
        case NUK_F64: return launch_map<AbsOp<double>>(p);
        // unsigned arrays take the signed symbol in the reference (package.lisp:112-115)
        case NUK_I64: case NUK_U64: return launch_map<AbsOp<int64_t>>(p);
        case NUK_I32: case NUK_U32: return launch_map<AbsOp<int32_t>>(p);
        case NUK_I16: case NUK_U16: return launch_map<AbsOp<int16_t>>(p);
        case NUK_I8: case NUK_U8: return launch_map<AbsOp<int8_t>>(p);
      }
      break;
    case NUK_ROUND: return floats<RoundOp>(op, dtype, p);
    case NUK_TRUNC: return floats<TruncOp>(op, dtype, p);
    case NUK_FLOOR: return floats<FloorOp>(op, dtype, p);
    case NUK_CEIL: return floats<CeilOp>(op, dtype, p);
    case NUK_SQRT: return floats<SqrtOp>(op, dtype, p);
  }
  set_error(NUK_ERR_UNSUPPORTED, "unary op %d: no kernel for dtype %d", op, dtype);
  return NUK_ERR_UNSUPPORTED;
}

int launch_op1(int op, int dtype, const MapProblem &p) {
  if (op >= NUK_SIN && op <= NUK_LOG) {
    if (dtype == NUK_F32) return launch_trans_f32(op, 1, p);
    if (dtype == NUK_F64) return launch_trans_f64(op, 1, p);
    set_error(NUK_ERR_UNSUPPORTED, "transcendental op %d needs a float dtype, got %d", op, dtype);
    return NUK_ERR_UNSUPPORTED;
  }
  return launch_unary(op, dtype, p);
}

template <typename TO> static int cast_to(int src, const MapProblem &p) {
  switch (src) {
    case NUK_F32: return launch_map<CastOp<float, TO>>(p);
    case NUK_F64: return launch_map<CastOp<double, TO>>(p);
    case NUK_I64: return launch_map<CastOp<int64_t, TO>>(p);
    case NUK_I32: return launch_map<CastOp<int32_t, TO>>(p);
    case NUK_I16: return launch_map<CastOp<int16_t, TO>>(p);
    case NUK_I8: return launch_map<CastOp<int8_t, TO>>(p);
    case NUK_U64: return launch_map<CastOp<uint64_t, TO>>(p);
    case NUK_U32: return launch_map<CastOp<uint32_t, TO>>(p);
    case NUK_U16: return launch_map<CastOp<uint16_t, TO>>(p);
    case NUK_U8: return launch_map<CastOp<uint8_t, TO>>(p);
  }
  set_error(NUK_ERR_UNSUPPORTED, "cast: unknown source dtype %d", src);
  return NUK_ERR_UNSUPPORTED;
}

// nu::to-single-float / nu::to-double-float rows of the table (package.lisp:124-129);
// same-type pairs are nu:copy.
int launch_cast(int src, int dst, const MapProblem &p) {
  if (src == dst) return launch_unary(NUK_COPY, src, p);
  if (dst == NUK_F32) return cast_to<float>(src, p);
  if (dst == NUK_F64) return cast_to<double>(src, p);
  set_error(NUK_ERR_UNSUPPORTED, "cast %d -> %d: the reference only casts to single/double-float", src, dst);
  return NUK_ERR_UNSUPPORTED;
}

}  // namespace nuk
