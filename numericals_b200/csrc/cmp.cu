// cmp.cu -- nu:two-arg-< <= = /= >= > into (unsigned-byte 8) 0/1.
// Table: src/basic-math/package.lisp:182-199.
#include "dispatch.cuh"
#include "ops.cuh"

namespace nuk {

template <template <class> class Op> static int ordered(int dtype, const MapProblem &p) {
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
  set_error(NUK_ERR_UNSUPPORTED, "no compare kernel for dtype %d", dtype);
  return NUK_ERR_UNSUPPORTED;
}
// = and /= do not depend on signedness
template <template <class> class Op> static int equality(int dtype, const MapProblem &p) {
  switch (dtype) {
    case NUK_F32: return launch_map<Op<float>>(p);
    case NUK_F64: return launch_map<Op<double>>(p);
    case NUK_I64: case NUK_U64: return launch_map<Op<uint64_t>>(p);
    case NUK_I32: case NUK_U32: return launch_map<Op<uint32_t>>(p);
    case NUK_I16: case NUK_U16: return launch_map<Op<uint16_t>>(p);
    case NUK_I8: case NUK_U8: return launch_map<Op<uint8_t>>(p);
  }
  set_error(NUK_ERR_UNSUPPORTED, "no compare kernel for dtype %d", dtype);
  return NUK_ERR_UNSUPPORTED;
}

int launch_cmp(int cmp, int dtype, const MapProblem &p) {
  switch (cmp) {
    case NUK_LT: return ordered<LtOp>(dtype, p);
    case NUK_LE: return ordered<LeOp>(dtype, p);
    case NUK_GE: return ordered<GeOp>(dtype, p);
    case NUK_GT: return ordered<GtOp>(dtype, p);
    case NUK_EQ: return equality<EqOp>(dtype, p);
    case NUK_NEQ: return equality<NeqOp>(dtype, p);
  }
  set_error(NUK_ERR_UNSUPPORTED, "unknown comparison %d", cmp);
  return NUK_ERR_UNSUPPORTED;
}

}  // namespace nuk
