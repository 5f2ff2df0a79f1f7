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
    case NUK_POW: case NUK_ATAN2: case NUK_LOGB:
      if (dtype == NUK_F32) return launch_trans_f32(op, 2, p);
      if (dtype == NUK_F64) return launch_trans_f64(op, 2, p);
      set_error(NUK_ERR_UNSUPPORTED, "op %d needs a float dtype, got %d", op, dtype);
      return NUK_ERR_UNSUPPORTED;
    default:
      return launch_bitwise(op, dtype, p);
  }
}

}  // namespace nuk

// ------------------------------------------------------------------ fused n-ary fold
// nu:+ nu:- nu:* nu:/ nu:max nu:min over k same-shape contiguous arrays: the reference folds the
// binary op into OUT with k-1 full passes (src/basic-math/n-arg-fn.lisp:59-88: 3(k-1) streams);
// this kernel reads the k inputs once and writes OUT once (k+1 streams).  The fold order and every
// intermediate rounding are the same -- ((a0 op a1) op a2) ... in registers -- so the result is
// bit-identical to the pass-by-pass fold.
namespace nuk {
constexpr int kFoldMax = 8;
struct FoldArgs {
  const void *x[kFoldMax];
  int nin;
};
template <class F>
__global__ void __launch_bounds__(kThreads)
k_fold_flat(size_t n, FoldArgs a, typename F::Out *out, F f) {
  using T = typename F::In;
  constexpr int E = 16 / sizeof(T);
  using P = Pack<T, E>;
  constexpr int U = 2;
  const size_t npack = n / E, tile = (size_t)kThreads * U, nfull = npack / tile;
  if (blockIdx.x < nfull) {
    const size_t base = (size_t)blockIdx.x * tile + threadIdx.x;
    P acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = ld_stream(reinterpret_cast<const P *>(a.x[0]) + base + (size_t)u * kThreads);
    for (int j = 1; j < a.nin; ++j) {
      P v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ld_stream(reinterpret_cast<const P *>(a.x[j]) + base + (size_t)u * kThreads);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int e = 0; e < E; ++e) acc[u].v[e] = f(acc[u].v[e], v[u].v[e]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) st_stream(reinterpret_cast<P *>(out) + base + (size_t)u * kThreads, acc[u]);
  } else {
    for (size_t i = nfull * tile * E + threadIdx.x; i < n; i += kThreads) {
      T acc = static_cast<const T *>(a.x[0])[i];
      for (int j = 1; j < a.nin; ++j) acc = f(acc, static_cast<const T *>(a.x[j])[i]);
      out[i] = acc;
    }
  }
}
template <class F> static int fold_launch(size_t n, const FoldArgs &a, void *out, cudaStream_t st) {
  using T = typename F::In;
  constexpr int E = 16 / sizeof(T);
  const size_t nfull = n / E / ((size_t)kThreads * 2);
  k_fold_flat<F><<<(unsigned)(nfull + 1), kThreads, 0, st>>>(n, a, static_cast<T *>(out), F());
  return post_launch("k_fold_flat");
}
template <template <class> class Op> static int fold_by_type(int dtype, size_t n, const FoldArgs &a, void *out, cudaStream_t st) {
  switch (dtype) {
    case NUK_F32: return fold_launch<Op<float>>(n, a, out, st);
    case NUK_F64: return fold_launch<Op<double>>(n, a, out, st);
    case NUK_I64: return fold_launch<Op<int64_t>>(n, a, out, st);
    case NUK_I32: return fold_launch<Op<int32_t>>(n, a, out, st);
    case NUK_I16: return fold_launch<Op<int16_t>>(n, a, out, st);
    case NUK_I8: return fold_launch<Op<int8_t>>(n, a, out, st);
    case NUK_U64: return fold_launch<Op<uint64_t>>(n, a, out, st);
    case NUK_U32: return fold_launch<Op<uint32_t>>(n, a, out, st);
    case NUK_U16: return fold_launch<Op<uint16_t>>(n, a, out, st);
    case NUK_U8: return fold_launch<Op<uint8_t>>(n, a, out, st);
  }
  set_error(NUK_ERR_UNSUPPORTED, "fold: no kernel for dtype %d", dtype);
  return NUK_ERR_UNSUPPORTED;
}
}  // namespace nuk

NUK_API int nuk_fold(int op, int dtype, int64_t n, int nin, const void *const *xs, void *out, nuk_stream_t s) {
  using namespace nuk;
  if (nin < 2 || nin > kFoldMax || n < 0 || !xs || !out) {
    set_error(NUK_ERR_INVALID, "nuk_fold: 2 <= nin <= %d contiguous inputs required", kFoldMax);
    return NUK_ERR_INVALID;
  }
  if (!device_info()) return nuk_last_status();
  if (n == 0) return NUK_OK;
  FoldArgs a;
  a.nin = nin;
  bool ok = aligned_to(out, 16);
  for (int j = 0; j < nin; ++j) {
    a.x[j] = xs[j];
    ok = ok && xs[j] && aligned_to(xs[j], 16);
  }
  if (!ok) {
    set_error(NUK_ERR_INVALID, "nuk_fold: operands must be non-null and 16-byte aligned");
    return NUK_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)s;
  switch (op) {
    case NUK_ADD: return fold_by_type<AddOp>(dtype, (size_t)n, a, out, st);
    case NUK_SUB: return fold_by_type<SubOp>(dtype, (size_t)n, a, out, st);
    case NUK_MUL: return fold_by_type<MulOp>(dtype, (size_t)n, a, out, st);
    case NUK_MAX: return fold_by_type<MaxOp>(dtype, (size_t)n, a, out, st);
    case NUK_MIN: return fold_by_type<MinOp>(dtype, (size_t)n, a, out, st);
    case NUK_DIV:
      if (dtype == NUK_F32) return fold_launch<DivOp<float>>((size_t)n, a, out, st);
      if (dtype == NUK_F64) return fold_launch<DivOp<double>>((size_t)n, a, out, st);
      break;
  }
  set_error(NUK_ERR_UNSUPPORTED, "fold: op %d / dtype %d has no translation", op, dtype);
  return NUK_ERR_UNSUPPORTED;
}
