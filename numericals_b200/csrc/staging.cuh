// staging.cuh -- host-pointer operands of the BMAS_* symbols: the chunked host -> HBM -> host pipeline.
//
// The reference hands BMAS plain Lisp vectors pinned against the GC for the duration of the call
// (with-pointers-to-vectors-data, src/basic-utils/common.lisp:207-212): ordinary PAGEABLE memory, any
// element stride.  Three ways a chunk of such an operand travels, chosen per operand:
//   direct   page-locked memory (nuk_host_alloc / nuk_host_register), |inc| == 1: DMA straight from / to
//            the caller's buffer;
//   bounce   pageable memory, |inc| == 1: host threads copy the chunk into a page-locked ring slot (the
//            driver's own pageable path is a single-threaded staging copy that serialises the pipeline),
//            DMA from there; results come back the same way;
//   strided  |inc| > 1 (or 0 for a reduction operand): host threads gather the view's elements into the
//            ring slot in logical order / scatter results to exactly the view's elements, so nothing
//            between them is read or written (src/basic-math/test-definition-macros.lisp:86-99) and the
//            DMA moves only useful bytes.
// Chunk k uses slot k mod S with its own stream: H2D(k+1), kernel(k), D2H(k-1) and the host copies overlap.
#pragma once
#include <functional>
#include "common.cuh"

namespace nuk {

// one operand of a BMAS-shaped (rank-1) call
struct Opnd {
  const void *ptr;
  long inc;
  size_t size;     // element size
  PtrKind kind;
};

// run fn(lo, hi) over [0, n) split across the host mover threads (the caller takes part); returns when
// every piece is done.  Re-entrant: concurrent callers share the pool.
void host_parallel_for(size_t n, size_t min_grain, const std::function<void(size_t, size_t)> &fn);
int host_mover_threads();

struct Stager {
  static constexpr int kMaxSlots = 4;
  int nslots = 0;
  size_t cap = 0;                       // bytes per operand buffer
  cudaStream_t streams[kMaxSlots] = {};
  void *buf[kMaxSlots][3] = {};         // device, [slot][x, y, out]
  void *bounce[kMaxSlots][3] = {};      // page-locked host, allocated on first use
  cudaEvent_t entry = nullptr;
  int device = -1;
  struct Pending {                      // output of a slot still to be delivered to the caller's buffer
    bool active = false;
    Opnd out{};
    long c0 = 0, m = 0;
  } pend[kMaxSlots];
  bool dirty[kMaxSlots] = {};           // the slot's ring buffers are referenced by work in flight

  ~Stager() { release(); }
  void release();
  int ensure(size_t want_cap, int want_slots, int dev);
  // make `slot` reusable: wait for its stream if its ring buffers are in flight, deliver a pending output
  int begin_slot(int slot);
  // stage elements [c0, c0 + m) of a HOST operand; *dptr / *dinc walk the staged copy in logical order
  int in(int slot, int which, const Opnd &o, long c0, long m, const void **dptr, long *dinc);
  // where the kernel writes a HOST output chunk
  void out_target(int slot, const Opnd &o, long m, void **dptr, long *dinc);
  // enqueue the trip back of the chunk the kernel wrote at out_target
  int out(int slot, const Opnd &o, long c0, long m);
  // wait for everything and deliver all pending outputs
  int finish_all();

 private:
  int bounce_buf(int slot, int which, void **p);
};
Stager &thread_stager();

inline size_t stage_chunk_bytes(size_t total_bytes) {
  size_t cap = (size_t)(opt(OPT_STAGE_MB) > 0 ? opt(OPT_STAGE_MB) : 32) << 20;
  if (total_bytes < cap) cap = ((total_bytes + 255) / 256) * 256;
  return cap;
}

}  // namespace nuk
