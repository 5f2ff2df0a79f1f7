// arith.cu -- nu:add subtract multiply divide two-arg-max two-arg-min + bitwise family.
// Table: src/basic-math/package.lisp:131-147,166-180.
#include "dispatch.cuh"
#include "ops.cuh"

namespace nuk {

template <template <class> class Op> static int by_type_all(int dtype, const MapProblem &p) {
  switch (dtype) {
    case NUK_F32: return launch_map<Op<float>>(p);
    case NUK_F64: return launch_map<Op<double>>(p);
    case NUK_I64: return launch_map<Op<int64_t>>(p);
    case NUK_I32: return launch_map<Op<int32_t>>(p);
    case NUK_I16: return launch_map<Op<int16_t>>(p);
    case NUK_I8: return launch_map<Op<int8_t>>(p);
    case NUK_U64: return launch_map<Op<uint64_t>>(p);
    case NUK_U32: return launch_map<Op<uint32_t>>(p);
    case NUK_U16: return launch_map<Op<uint16_t>>(p);
    case NUK_U8: return launch_map<Op<uint8_t>>(p);
  }
  set_error(NUK_ERR_UNSUPPORTED, "no kernel for dtype %d", dtype);
  return NUK_ERR_UNSUPPORTED;
}

// add/sub/mul wrap identically for signed and unsigned: one kernel per width
// (the reference itself routes unsigned add/sub to the signed symbols, package.lisp:131-136)
template <template <class> class Op> static int by_type_wrap(int dtype, const MapProblem &p) {
  switch (dtype) {
    case NUK_F32: return launch_map<Op<float>>(p);
    case NUK_F64: return launch_map<Op<double>>(p);
    case NUK_I64: case NUK_U64: return launch_map<Op<int64_t>>(p);
    case NUK_I32: case NUK_U32: return launch_map<Op<int32_t>>(p);
    case NUK_I16: case NUK_U16: return launch_map<Op<int16_t>>(p);
    case NUK_I8: case NUK_U8: return launch_map<Op<int8_t>>(p);
  }
  set_error(NUK_ERR_UNSUPPORTED, "no kernel for dtype %d", dtype);
  return NUK_ERR_UNSUPPORTED;
}

int launch_arith(int op, int dtype, const MapProblem &p) {
  switch (op) {
    case NUK_ADD: return by_type_wrap<AddOp>(dtype, p);
    case NUK_SUB: return by_type_wrap<SubOp>(dtype, p);
    case NUK_MUL: return by_type_wrap<MulOp>(dtype, p);
    case NUK_MAX: return by_type_all<MaxOp>(dtype, p);
    case NUK_MIN: return by_type_all<MinOp>(dtype, p);
    case NUK_DIV:
      // nu:divide has translations for single/double-float only (package.lisp:140)
      if (dtype == NUK_F32) return launch_map<DivOp<float>>(p);
      if (dtype == NUK_F64) return launch_map<DivOp<double>>(p);
      set_error(NUK_ERR_UNSUPPORTED, "divide: no CFUNCTION for integer dtype %d", dtype);
      return NUK_ERR_UNSUPPORTED;
  }
  set_error(NUK_ERR_UNSUPPORTED, "unknown arithmetic op %d", op);
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

int launch_bitwise(int op, int dtype, const MapProblem &p) {
  if (dtype == NUK_F32 || dtype == NUK_F64) {
    set_error(NUK_ERR_UNSUPPORTED, "bitwise op %d: no CFUNCTION for float dtype %d", op, dtype);
    return NUK_ERR_UNSUPPORTED;
  }
  switch (op) {
    case NUK_AND: return by_width<AndOp>(dtype, p);
    case NUK_OR: return by_width<OrOp>(dtype, p);
    case NUK_XOR: return by_width<XorOp>(dtype, p);
    case NUK_ANDNOT: return by_width<AndNotOp>(dtype, p);
    case NUK_SRA:
      if (dtype == NUK_I64) return launch_map<SraOp>(p);
      break;
  }
  set_error(NUK_ERR_UNSUPPORTED, "unknown bitwise op %d for dtype %d", op, dtype);
  return NUK_ERR_UNSUPPORTED;
}

int launch_op2(int op, int dtype, const MapProblem &p) {
  switch (op) {
    case NUK_ADD: case NUK_SUB: case NUK_MUL: case NUK_DIV: case NUK_MAX: case NUK_MIN:
      return launch_arith(op, dtype, p);
    case NUK_POW: case NUK_ATAN2:
      if (dtype == NUK_F32) return launch_trans_f32(op, 2, p);
      if (dtype == NUK_F64) return launch_trans_f64(op, 2, p);
      set_error(NUK_ERR_UNSUPPORTED, "op %d needs a float dtype, got %d", op, dtype);
      return NUK_ERR_UNSUPPORTED;
    default:
      return launch_bitwise(op, dtype, p);
  }
}

}  // namespace nuk
