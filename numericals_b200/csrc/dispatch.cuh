// dispatch.cuh -- (op, dtype) -> kernel-family launchers, one translation unit per family
// so that nvcc can build them in parallel.  This is the device-side counterpart of the
// reference's translation table (src/utils/translations.lisp:66-121): an (op, dtype) pair
// with no entry returns NUK_ERR_UNSUPPORTED, like "No CFUNCTION known for ...".
#pragma once
#include "map.cuh"

namespace nuk {

int launch_arith(int op, int dtype, const MapProblem &p);      // add sub mul div max min
int launch_bitwise(int op, int dtype, const MapProblem &p);    // and or xor andnot sra (+ not via op1)
int launch_cmp(int cmp, int dtype, const MapProblem &p);       // lt le eq neq ge gt -> u8
int launch_unary(int op, int dtype, const MapProblem &p);      // copy abs round trunc floor ceil sqrt not
int launch_cast(int src, int dst, const MapProblem &p);        // any -> f32 / f64 (+ same-type copy)
int launch_trans_f32(int op, int arity, const MapProblem &p);  // transcendental unary / pow atan2
int launch_trans_f64(int op, int arity, const MapProblem &p);

// dims/strides (elements) -> normalised MapProblem with shifted base pointers (nd.cu)
int make_problem(MapProblem *p, int arity, size_t in_size, size_t out_size, int rank,
                 const int64_t *dims, const int64_t *sx, const int64_t *sy, const int64_t *so,
                 const void *x, const void *y, void *out, cudaStream_t stream);

// uniform entry points used by the C ABI layers (nd.cu, bmas.cu, staging.cu)
int launch_op2(int op, int dtype, const MapProblem &p);
int launch_op1(int op, int dtype, const MapProblem &p);

// reductions (reduce.cu)
struct ReduceProblem {
  int red, dtype;
  int rank;                        // coalesced
  int64_t dims[NUK_MAX_RANK];
  int64_t sx[NUK_MAX_RANK];
  int axis;                        // -1 = all
  int64_t so[NUK_MAX_RANK];        // strides of out per x-dim (0 at the reduced axis)
  const void *x;
  void *out;                       // device
  void *out2;                      // second output for fused moments (sum of squares), or null
  int flags;
  cudaStream_t stream;
  // sharded FULL reduction (comm.cu): pass 2 is fused with the rank-ordered exchange over peer memory,
  // pass-1 partials live in the communicator's own buffer (no allocation inside a collective)
  const struct CommLink *link;
  void *link_partials;
  size_t link_partial_bytes;
};
int launch_reduce(const ReduceProblem &p);
// 1-D strided reduction of n elements into ONE device element (BMAS-shaped call)
// flags: NUK_REDUCE_NO_SEED for a chunk that does not start at element 0 of the caller's array
int reduce_1d_device(int red, int dtype, size_t n, const void *x, int64_t incx, void *out_dev,
                     cudaStream_t stream, int flags = 0);
constexpr int NUK_REDUCE_NO_SEED = 0x100;   // internal: float hmax/hmin without the x[0] rule (see k_combine_one)
// combine `count` partial results (device, contiguous) into one device element, fixed order
int reduce_partials_device(int red, int dtype, size_t count, const void *partials, void *out_dev,
                           cudaStream_t stream);

}  // namespace nuk
