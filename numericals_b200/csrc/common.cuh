// common.cuh -- shared host/device plumbing for libnuk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <type_traits>
#include "../../include/nuk.h"

#define NUK_API extern "C" __attribute__((visibility("default")))
#define NUK_HD __host__ __device__ __forceinline__
#define NUK_D __device__ __forceinline__

namespace nuk {

// ------------------------------------------------------------------ errors
void set_error(int status, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define NUK_CUDA(expr)                                                    \
  do {                                                                    \
    cudaError_t _e = (expr);                                              \
    if (_e != cudaSuccess) return nuk::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)
#define NUK_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != NUK_OK) return _s; \
  } while (0)

// ------------------------------------------------------------------ runtime state
struct DeviceInfo {
  int device;
  int sm_count;
  size_t l2_bytes;
};
const DeviceInfo *device_info();      // current device of this thread (lazy init); nullptr + error if none
int use_device(int dev);              // make `dev` the calling thread's device (CUDA's and libnuk's notion)
cudaStream_t current_stream();        // thread-local stream used by BMAS_* symbols
void count_launch(unsigned n = 1);    // gpu_launches bookkeeping
int post_launch(const char *kernel);  // cudaGetLastError -> status (+ the NUK_LOG per-launch line)
// NVTX range around an entry point of the C ABI (NUK_NVTX=1 in the environment; otherwise one predictable branch):
// lets `ncu --nvtx --nvtx-include "BMAS/"` or a timeline tool cut a host program's calls out of everything else
struct ApiRange {
  bool on;
  explicit ApiRange(const char *name);
  ~ApiRange();
};

// NUK_* knobs: lock-free reads on the launch path (nuk_set_option / the environment write them)
enum Opt {
  OPT_GRID_MULT = 0, OPT_REDUCE_GRID_MULT, OPT_STAGE_MB, OPT_STAGE_SLOTS, OPT_COL_SPLIT_ROWS, OPT_SEQ_COLS,
  OPT_STAGE_THREADS, OPT_ROWS_CTAS_PER_SM, OPT_STAGE_BOUNCE, OPT_ROWS_RPC, OPT_PDL, OPT_STAGE_NT, OPT_COUNT
};
long opt(Opt o);

// pointer classification of a BMAS_* operand.  NUK_PTR_AUTO asks the driver (cudaPointerGetAttributes);
// in NUK_PTR_DEVICE mode (nuk_set_pointer_mode: the dense-arrays device-storage backend knows where its
// arrays live) strided operands are taken as device pointers without a driver call.
enum PtrKind { PK_NULL = 0, PK_DEVICE = 1, PK_HOST = 2 /* page-locked */, PK_PAGEABLE = 3 };
PtrKind classify_ptr(const void *p, bool strided);
inline bool on_host(PtrKind k) { return k == PK_HOST || k == PK_PAGEABLE; }

// scratch memory private to (thread, device, stream): reduction partials etc.
int scratch(cudaStream_t s, size_t bytes, void **dptr);
// same, plus the arena's zero-initialised, self-re-arming arrival ticket (single-launch full reductions)
int scratch_with_ticket(cudaStream_t s, size_t bytes, void **dptr, unsigned **ticket);
// pinned result slot private to the thread (8 * 64 bytes, portable across devices)
int pinned_slot(void **hptr);
// small device buffers private to (thread, current device, slot), grown on demand
enum BufSlot { BUF_RESULT = 0, BUF_PARTS, BUF_HS_X, BUF_HS_Y, BUF_COUNT };
int thread_buffer(BufSlot slot, size_t bytes, void **dptr);

// grid size for a persistent / grid-stride kernel: min(work_blocks, sm_count * mult)
int grid_for(size_t work_blocks, int blocks_per_sm);

// ------------------------------------------------------------------ dtype traits
template <int DT> struct CType;
template <> struct CType<NUK_F32> { using type = float; };
template <> struct CType<NUK_F64> { using type = double; };
template <> struct CType<NUK_I64> { using type = int64_t; };
template <> struct CType<NUK_I32> { using type = int32_t; };
template <> struct CType<NUK_I16> { using type = int16_t; };
template <> struct CType<NUK_I8> { using type = int8_t; };
template <> struct CType<NUK_U64> { using type = uint64_t; };
template <> struct CType<NUK_U32> { using type = uint32_t; };
template <> struct CType<NUK_U16> { using type = uint16_t; };
template <> struct CType<NUK_U8> { using type = uint8_t; };

inline size_t dtype_size(int dt) {
  switch (dt) {
    case NUK_F32: case NUK_I32: case NUK_U32: return 4;
    case NUK_F64: case NUK_I64: case NUK_U64: return 8;
    case NUK_I16: case NUK_U16: return 2;
    case NUK_I8: case NUK_U8: return 1;
    default: return 0;
  }
}

// a scalar operand travelling by value in the kernel parameters
union Scalar64 {
  double d; float f; int64_t i64; uint64_t u64; int32_t i32; uint32_t u32;
  int16_t i16; uint16_t u16; int8_t i8; uint8_t u8; unsigned char bytes[8];
};
template <typename T> NUK_HD T scalar_as(const Scalar64 &s) {
  T v;
  memcpy(&v, s.bytes, sizeof(T));
  return v;
}

// ------------------------------------------------------------------ vector packs
constexpr int kThreads = 256;

template <typename T, int N> struct alignas(sizeof(T) * N) Pack { T v[N]; };

// streaming (evict-first) 2/4/8/16-byte global accesses. Plain ld.global/st.global with the
// .cs policy: legal when OUT aliases an input (the non-coherent .nc path is not).
template <int BYTES> struct RawVec;
template <> struct RawVec<1> { using type = unsigned char; };
template <> struct RawVec<2> { using type = unsigned short; };
template <> struct RawVec<4> { using type = unsigned int; };
template <> struct RawVec<8> { using type = uint2; };
template <> struct RawVec<16> { using type = uint4; };

template <typename P> NUK_D P ld_stream(const P *p) {
  using R = typename RawVec<sizeof(P)>::type;
  R r = __ldcs(reinterpret_cast<const R *>(p));
  P out;
  memcpy(&out, &r, sizeof(P));
  return out;
}
template <typename P> NUK_D void st_stream(P *p, const P &v) {
  using R = typename RawVec<sizeof(P)>::type;
  R r;
  memcpy(&r, &v, sizeof(P));
  __stcs(reinterpret_cast<R *>(p), r);
}
template <typename P> NUK_D P ld_keep(const P *p) {  // default policy (operand is re-read: keep in L1/L2)
  using R = typename RawVec<sizeof(P)>::type;
  R r = *reinterpret_cast<const R *>(p);
  P out;
  memcpy(&out, &r, sizeof(P));
  return out;
}

inline bool aligned_to(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// functors over 8-bit elements may provide a SIMD-in-a-word form (specialised in ops.cuh):
//   static uint32_t op(uint32_t a, uint32_t b)   -- four elements per call
template <class F> struct Word4 { static constexpr bool ok = false; };

// ------------------------------------------------------------------ nd descriptor (device side)
struct NdDesc {
  int rank;                       // <= NUK_MAX_RANK, after coalescing
  int64_t dims[NUK_MAX_RANK];
  int64_t stride[3][NUK_MAX_RANK];  // operand 0 = out, 1 = x, 2 = y (elements)
};

// host side: normalise (dims, strides of nops operands) -> compact descriptor.
// strides layout: strides[op * NUK_MAX_RANK + axis]
int coalesce(int nops, int rank, int64_t *dims, int64_t *strides);

}  // namespace nuk
