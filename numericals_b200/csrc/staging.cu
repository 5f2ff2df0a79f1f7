// staging.cu -- the host mover pool and the per-thread staging ring (see staging.cuh).
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include "staging.cuh"

namespace nuk {

// ------------------------------------------------------------------ host mover pool
// Plain worker threads that do nothing but memcpy / strided gather-scatter between the caller's
// buffers and the page-locked ring.  One pool per process, started on first use and never torn down
// (a library unloaded at exit must not join threads from a static destructor).
namespace {
struct Job {
  std::atomic<size_t> remaining{0};
  const std::function<void(size_t, size_t)> *fn = nullptr;
};
struct Task {
  Job *job;
  size_t lo, hi;
};
struct Pool {
  std::mutex mu;
  std::condition_variable cv;        // work available / a job finished
  std::deque<Task> q;
  int nthreads = 0;
  void worker() {
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      while (q.empty()) cv.wait(lk);
      Task t = q.front();
      q.pop_front();
      lk.unlock();
      (*t.job->fn)(t.lo, t.hi);
      lk.lock();
      if (t.job->remaining.fetch_sub(1, std::memory_order_acq_rel) == 1) cv.notify_all();
    }
  }
};
Pool *pool() {
  static Pool *p = [] {
    Pool *pp = new Pool;
    long want = opt(OPT_STAGE_THREADS);
    if (want <= 0) {
      // all cores but the caller's, at most 15: the copies are bandwidth-bound and 15 movers were 1.26x faster
      // than 8 on a 16-vCPU box (profiles/time_pageable_r02_a.txt); the reference's own lparallel kernel
      // also takes every core (src/transcendental/lparallel.lisp:60-64)
      const unsigned hc = std::thread::hardware_concurrency();
      want = hc >= 2 ? (long)(hc - 1) : 1;
      if (want > 15) want = 15;
    }
    if (want > 64) want = 64;
    pp->nthreads = (int)want;
    for (int i = 0; i < pp->nthreads; ++i) std::thread([pp] { pp->worker(); }).detach();
    return pp;
  }();
  return p;
}
}  // namespace

int host_mover_threads() { return pool()->nthreads; }

void host_parallel_for(size_t n, size_t min_grain, const std::function<void(size_t, size_t)> &fn) {
  if (n == 0) return;
  Pool *p = pool();
  size_t pieces = (size_t)(p->nthreads + 1) * 2;
  if (min_grain == 0) min_grain = 1;
  if (pieces > (n + min_grain - 1) / min_grain) pieces = (n + min_grain - 1) / min_grain;
  if (pieces <= 1) {
    fn(0, n);
    return;
  }
  Job job;
  job.fn = &fn;
  job.remaining.store(pieces, std::memory_order_relaxed);
  const size_t per = (n + pieces - 1) / pieces;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    for (size_t i = 0; i < pieces; ++i) {
      const size_t lo = i * per, hi = (lo + per < n) ? lo + per : n;
      if (lo >= hi) {
        job.remaining.fetch_sub(1, std::memory_order_relaxed);
        continue;
      }
      p->q.push_back(Task{&job, lo, hi});
    }
  }
  p->cv.notify_all();
  // the caller works too (any queued piece: keeps concurrent callers deadlock-free), then waits
  std::unique_lock<std::mutex> lk(p->mu);
  while (job.remaining.load(std::memory_order_acquire) != 0) {
    if (!p->q.empty()) {
      Task t = p->q.front();
      p->q.pop_front();
      lk.unlock();
      (*t.job->fn)(t.lo, t.hi);
      lk.lock();
      if (t.job->remaining.fetch_sub(1, std::memory_order_acq_rel) == 1) p->cv.notify_all();
    } else {
      p->cv.wait(lk);
    }
  }
}

// ------------------------------------------------------------------ host copies
// the address of logical element i of a strided host operand
static inline const char *elem_addr(const Opnd &o, int64_t i) {
  return static_cast<const char *>(o.ptr) + i * (int64_t)o.inc * (int64_t)o.size;
}
// lowest address touched by logical elements [c0, c0 + m)
static inline const char *chunk_lo(const Opnd &o, long c0, long m) {
  return o.inc >= 0 ? elem_addr(o, c0) : elem_addr(o, c0 + m - 1);
}

template <typename U> static void gather_t(void *dst, const Opnd &o, long c0, size_t lo, size_t hi) {
  U *d = static_cast<U *>(dst);
  const U *s = reinterpret_cast<const U *>(elem_addr(o, c0));
  const int64_t inc = o.inc;
  for (size_t i = lo; i < hi; ++i) d[i] = s[(int64_t)i * inc];
}
template <typename U> static void scatter_t(const void *src, const Opnd &o, long c0, size_t lo, size_t hi) {
  const U *s = static_cast<const U *>(src);
  U *d = reinterpret_cast<U *>(const_cast<char *>(elem_addr(o, c0)));
  const int64_t inc = o.inc;
  for (size_t i = lo; i < hi; ++i) d[(int64_t)i * inc] = s[i];
}
static void gather(void *dst, const Opnd &o, long c0, long m) {
  host_parallel_for((size_t)m, (size_t)1 << 16, [&](size_t lo, size_t hi) {
    switch (o.size) {
      case 1: gather_t<uint8_t>(dst, o, c0, lo, hi); break;
      case 2: gather_t<uint16_t>(dst, o, c0, lo, hi); break;
      case 4: gather_t<uint32_t>(dst, o, c0, lo, hi); break;
      default: gather_t<uint64_t>(dst, o, c0, lo, hi); break;
    }
  });
}
static void scatter(const void *src, const Opnd &o, long c0, long m) {
  host_parallel_for((size_t)m, (size_t)1 << 16, [&](size_t lo, size_t hi) {
    switch (o.size) {
      case 1: scatter_t<uint8_t>(src, o, c0, lo, hi); break;
      case 2: scatter_t<uint16_t>(src, o, c0, lo, hi); break;
      case 4: scatter_t<uint32_t>(src, o, c0, lo, hi); break;
      default: scatter_t<uint64_t>(src, o, c0, lo, hi); break;
    }
  });
}
void copy_stream(void *dst, const void *src, size_t n);   // hostcopy.cpp: non-temporal stores
static void par_memcpy(void *dst, const void *src, size_t bytes, bool streaming = true) {
  host_parallel_for(bytes, (size_t)1 << 20, [&](size_t lo, size_t hi) {
    if (streaming) copy_stream(static_cast<char *>(dst) + lo, static_cast<const char *>(src) + lo, hi - lo);
    else memcpy(static_cast<char *>(dst) + lo, static_cast<const char *>(src) + lo, hi - lo);
  });
}

// ------------------------------------------------------------------ the ring
void Stager::release() {
  for (int s = 0; s < kMaxSlots; ++s) {
    for (int o = 0; o < 3; ++o) {
      if (buf[s][o]) cudaFree(buf[s][o]);
      if (bounce[s][o]) cudaFreeHost(bounce[s][o]);
    }
    if (streams[s]) cudaStreamDestroy(streams[s]);
    pend[s].active = false;
    dirty[s] = false;
  }
  if (entry) cudaEventDestroy(entry);
  nslots = 0; cap = 0; entry = nullptr;
  memset(buf, 0, sizeof(buf));
  memset(bounce, 0, sizeof(bounce));
  memset(streams, 0, sizeof(streams));
}

int Stager::ensure(size_t want_cap, int want_slots, int dev) {
  if (want_slots > kMaxSlots) want_slots = kMaxSlots;
  if (want_slots < 1) want_slots = 1;
  if (nslots == want_slots && cap >= want_cap && device == dev) return NUK_OK;
  release();
  device = dev;
  for (int s = 0; s < want_slots; ++s) {
    NUK_CUDA(cudaStreamCreateWithFlags(&streams[s], cudaStreamNonBlocking));
    for (int o = 0; o < 3; ++o) NUK_CUDA(cudaMalloc(&buf[s][o], want_cap));
    nslots = s + 1;
  }
  NUK_CUDA(cudaEventCreateWithFlags(&entry, cudaEventDisableTiming));
  cap = want_cap;
  return NUK_OK;
}

int Stager::bounce_buf(int slot, int which, void **p) {
  if (!bounce[slot][which]) NUK_CUDA(cudaHostAlloc(&bounce[slot][which], cap, cudaHostAllocPortable));
  *p = bounce[slot][which];
  return NUK_OK;
}

int Stager::begin_slot(int slot) {
  if (!dirty[slot] && !pend[slot].active) return NUK_OK;
  NUK_CUDA(cudaStreamSynchronize(streams[slot]));
  dirty[slot] = false;
  if (pend[slot].active) {
    const Pending &p = pend[slot];
    const void *src = bounce[slot][2];
    if (p.out.inc == 1 || p.out.inc == -1)
      par_memcpy(const_cast<char *>(chunk_lo(p.out, p.c0, p.m)), src, (size_t)p.m * p.out.size);
    else
      scatter(src, p.out, p.c0, p.m);
    pend[slot].active = false;
  }
  return NUK_OK;
}

int Stager::in(int slot, int which, const Opnd &o, long c0, long m, const void **dptr, long *dinc) {
  cudaStream_t st = streams[slot];
  void *dbuf = buf[slot][which];
  const size_t bytes = (size_t)m * o.size;
  const bool unit = (o.inc == 1 || o.inc == -1);
  if (unit && (o.kind == PK_HOST || opt(OPT_STAGE_BOUNCE) == 0)) {             // direct
    NUK_CUDA(cudaMemcpyAsync(dbuf, chunk_lo(o, c0, m), bytes, cudaMemcpyHostToDevice, st));
  } else {
    void *bb = nullptr;
    NUK_TRY(bounce_buf(slot, which, &bb));
    if (unit) par_memcpy(bb, chunk_lo(o, c0, m), bytes, opt(OPT_STAGE_NT) != 0);   // address order
    else gather(bb, o, c0, m);                             // logical order
    NUK_CUDA(cudaMemcpyAsync(dbuf, bb, bytes, cudaMemcpyHostToDevice, st));
    dirty[slot] = true;
  }
  if (o.inc == -1) {   // staged in address order: walk it backwards
    *dptr = static_cast<char *>(dbuf) + (size_t)(m - 1) * o.size;
    *dinc = -1;
  } else {
    *dptr = dbuf;
    *dinc = 1;
  }
  return NUK_OK;
}

void Stager::out_target(int slot, const Opnd &o, long m, void **dptr, long *dinc) {
  char *dbuf = static_cast<char *>(buf[slot][2]);
  if (o.inc == -1) {
    *dptr = dbuf + (size_t)(m - 1) * o.size;
    *dinc = -1;
  } else {
    *dptr = dbuf;
    *dinc = 1;
  }
}

int Stager::out(int slot, const Opnd &o, long c0, long m) {
  cudaStream_t st = streams[slot];
  const size_t bytes = (size_t)m * o.size;
  const bool unit = (o.inc == 1 || o.inc == -1);
  if (unit && (o.kind == PK_HOST || opt(OPT_STAGE_BOUNCE) == 0)) {
    NUK_CUDA(cudaMemcpyAsync(const_cast<char *>(chunk_lo(o, c0, m)), buf[slot][2], bytes, cudaMemcpyDeviceToHost, st));
    return NUK_OK;
  }
  void *bb = nullptr;
  NUK_TRY(bounce_buf(slot, 2, &bb));
  NUK_CUDA(cudaMemcpyAsync(bb, buf[slot][2], bytes, cudaMemcpyDeviceToHost, st));
  pend[slot].active = true;
  pend[slot].out = o;
  pend[slot].c0 = c0;
  pend[slot].m = m;
  dirty[slot] = true;
  return NUK_OK;
}

int Stager::finish_all() {
  int status = NUK_OK;
  for (int s = 0; s < nslots; ++s) {
    cudaError_t e = cudaStreamSynchronize(streams[s]);
    if (e != cudaSuccess && status == NUK_OK) status = cuda_fail(e, "staging stream sync", __FILE__, __LINE__);
    dirty[s] = false;
  }
  // deliver in chunk order: the slot holding the oldest chunk first does not matter for correctness
  // (chunks are disjoint), so plain slot order
  for (int s = 0; s < nslots; ++s) {
    if (pend[s].active) {
      if (status == NUK_OK) {
        const int st = begin_slot(s);
        if (st != NUK_OK) status = st;
      } else {
        pend[s].active = false;
      }
    }
  }
  return status;
}

Stager &thread_stager() {
  static thread_local Stager s;
  return s;
}

}  // namespace nuk

NUK_API int nuk_host_mover_threads(void) { return nuk::host_mover_threads(); }
