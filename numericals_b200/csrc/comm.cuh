// comm.cuh -- device side of the multi-GPU combine: rank-ordered all-reduce over NVLink peer memory.
//
// The hot path shards with no data-path collective (maps are independent per element); the only exchange is
// the combine step of a reduction: one scalar per rank for a full reduction, an `ncols` vector for an axis
// reduction over the sharded axis.  Those messages are 4 B ... 256 KB -- pure latency -- so
// instead of calling a collective library after the reduction, the LAST kernel of the reduction does the
// exchange itself: every rank owns a "mailbox" in its HBM that all peers can write (cudaIpc mapping between
// processes, cudaDeviceEnablePeerAccess inside one process); a CTA pushes its piece into slot [rank] of every
// peer's mailbox with plain NVLink stores, releases a flag, acquires the same flag of every peer in its own
// mailbox and then folds the W pieces IN RANK ORDER -- so every rank computes bit-identical results and the
// result for a given world size is reproducible (a ring / tree all-reduce may re-associate).
//
// Mailbox (device memory, one per rank):
//   flags [2 parities][kCommMaxWorld sources][kCommMaxCtas]  u32 sequence numbers
//   data  [2 parities][kCommMaxWorld sources][kCommSlotBytes]
// Collective number `seq` uses parity seq & 1; a rank can be at most one collective ahead of a peer (it needs
// the peer's flag of the previous one), so two parities make slot reuse safe without a second handshake.
#pragma once
#include "common.cuh"

namespace nuk {

constexpr int kCommMaxWorld = 8;
constexpr int kCommMaxCtas = 64;                    // co-resident by construction (<= SM count)
constexpr size_t kCommSlotBytes = (size_t)1 << 20;  // payload per (parity, source)
constexpr size_t kCommFlagBytes = 2 * kCommMaxWorld * kCommMaxCtas * sizeof(unsigned);
constexpr size_t kCommMailboxBytes = kCommFlagBytes + 2 * kCommMaxWorld * kCommSlotBytes;

struct CommLink {              // passed by value to the kernels
  char *mail[kCommMaxWorld];   // mailbox of every rank, as addresses valid on THIS device
  int world, rank;
  unsigned seq;                // sequence number of this collective (same on every rank)
  int *status;                 // host-mapped error word: set to 1 when a wait times out
  unsigned long long timeout_ns;
};

#ifdef __CUDACC__
NUK_D unsigned *mail_flag(char *base, int parity, int src, int cta) {
  return reinterpret_cast<unsigned *>(base) + ((size_t)parity * kCommMaxWorld + src) * kCommMaxCtas + cta;
}
NUK_D char *mail_data(char *base, int parity, int src) {
  return base + kCommFlagBytes + ((size_t)parity * kCommMaxWorld + src) * kCommSlotBytes;
}
NUK_D void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
NUK_D unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// L2 load of any trivially copyable value (peer stores land in L2/HBM; this SM's L1 may hold a stale line
// of the same slot from two collectives ago)
template <typename A> NUK_D A ld_cg_any(const void *p) {
  A r;
  if constexpr (sizeof(A) % 4 == 0) {
    unsigned w[sizeof(A) / 4];
#pragma unroll
    for (int i = 0; i < (int)(sizeof(A) / 4); ++i) w[i] = __ldcg(static_cast<const unsigned *>(p) + i);
    memcpy(&r, w, sizeof(A));
  } else {
    unsigned char b[sizeof(A)];
#pragma unroll
    for (int i = 0; i < (int)sizeof(A); ++i) b[i] = __ldcg(static_cast<const unsigned char *>(p) + i);
    memcpy(&r, b, sizeof(A));
  }
  return r;
}
NUK_D unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Called by every thread of the CTA after its pushes into the peers' data slots: publish "source `rank`,
// piece `cta` of collective `seq` has landed" to every peer, then wait until every source's piece `cta` has
// landed HERE.  Returns false on timeout (a peer never arrived): the caller must not read the slots.
NUK_D bool comm_exchange(const CommLink &L, int cta) {
  const int parity = (int)(L.seq & 1u);
  __shared__ int ok;
  __syncthreads();                       // all pushes of this CTA are issued
  if (threadIdx.x == 0) {
    ok = 1;
    __threadfence_system();              // ... and ordered before the flags below
  }
  __syncthreads();
  if (threadIdx.x < L.world) {
    st_release_sys(mail_flag(L.mail[threadIdx.x], parity, L.rank, cta), L.seq);
    const unsigned *mine = mail_flag(L.mail[L.rank], parity, threadIdx.x, cta);
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(mine) != L.seq) {
      if (global_ns() - t0 > L.timeout_ns) {
        ok = 0;
        if (L.status) *reinterpret_cast<volatile int *>(L.status) = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  return ok != 0;
}

// pass 2 of a sharded FULL reduction fused with the exchange: fold this rank's pass-1 partials, push the
// one value to every peer, fold the W values in rank order.  One CTA.  `first` as in k_combine_one, but it
// applies to the GLOBAL first element: the caller passes it on rank 0 only, and after the exchange every
// rank re-applies the rule to rank 0's value.
template <class R, typename BlockReduce>
NUK_D void combine_allgather_body(const CommLink &L, size_t count, const typename R::Acc *partials,
                                  typename R::In *o1, typename R::In *o2, typename R::Acc init,
                                  const typename R::In *first, BlockReduce block_reduce_fn) {
  using A = typename R::Acc;
  using T = typename R::In;
  A acc = init;
  for (size_t i = threadIdx.x; i < count; i += blockDim.x) acc = R::combine(acc, partials[i]);
  acc = block_reduce_fn(acc, init);
  const int parity = (int)(L.seq & 1u);
  if (threadIdx.x == 0) {
    if constexpr (std::is_floating_point<T>::value && std::is_same<A, T>::value) {
      if (first) {
        const T f0 = *first;
        if (f0 != f0) acc = f0;
      }
    }
    for (int k = 0; k < L.world; ++k) {
      const int p = (L.rank + k) % L.world;
      *reinterpret_cast<A *>(mail_data(L.mail[p], parity, L.rank)) = acc;
    }
  }
  if (!comm_exchange(L, 0)) return;
  if (threadIdx.x == 0) {
    char *mine = L.mail[L.rank];
    A r = ld_cg_any<A>(mail_data(mine, parity, 0));
    [[maybe_unused]] const A r0 = r;
    for (int s = 1; s < L.world; ++s) r = R::combine(r, ld_cg_any<A>(mail_data(mine, parity, s)));
    if constexpr (std::is_floating_point<T>::value && std::is_same<A, T>::value) {
      if (r0 != r0) r = r0;              // hmax / hmin: a NaN in the global first position is the result
    }
    R::store(o1, o2, 0, r);
  }
}

// all-reduce of `count` elements (count * sizeof(T) <= kCommSlotBytes): CTA b owns the elements
// [b * per, (b+1) * per), per a multiple of 16 / sizeof(T) that depends on `count` and the grid only -- every
// rank cuts the payload the same way, because CTA b waits for exactly the pieces CTA b of its peers pushed.
// out[i] = local_0[i] (+) local_1[i] (+) ... in rank order.  `vec`: this rank's local / out are 16-byte
// aligned, so whole 16-byte groups move as one transfer (a per-rank choice: the slot layout is the same).
template <class R>
NUK_D void allreduce_body(const CommLink &L, int64_t count, const typename R::In *local, typename R::In *out, int vec) {
  using T = typename R::In;
  using A = typename R::Acc;
  constexpr int E = 16 / (int)sizeof(T);
  const int parity = (int)(L.seq & 1u);
  const int64_t groups = (count + E - 1) / E;
  const int64_t per = ((groups + gridDim.x - 1) / gridDim.x) * E;
  const int64_t lo = (int64_t)blockIdx.x * per;
  const int64_t hi = (lo + per < count) ? lo + per : count;
  const int64_t nv = (vec && hi > lo) ? (hi - lo) / E : 0;         // whole 16-byte groups of my piece
  const int64_t tail = lo + nv * E;                                // elements [tail, hi) go one by one
  // phase 1: my piece into slot [rank] of every peer (own mailbox included), peers visited round-robin
  for (int k = 0; k < L.world; ++k) {
    const int p = (L.rank + k) % L.world;
    T *dst = reinterpret_cast<T *>(mail_data(L.mail[p], parity, L.rank));
    for (int64_t i = threadIdx.x; i < nv; i += blockDim.x)
      reinterpret_cast<uint4 *>(dst + lo)[i] = reinterpret_cast<const uint4 *>(local + lo)[i];
    for (int64_t i = tail + threadIdx.x; i < hi; i += blockDim.x) dst[i] = local[i];
  }
  if (!comm_exchange(L, (int)blockIdx.x)) return;
  // phase 2: fold the W pieces in rank order (L2 loads: the peers' stores bypass this SM's L1)
  char *mine = L.mail[L.rank];
  for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) {
    Pack<T, E> v = ld_cg_any<Pack<T, E>>(reinterpret_cast<const uint4 *>(reinterpret_cast<const T *>(mail_data(mine, parity, 0)) + lo) + i);
    A acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = R::load(v.v[e]);
    for (int s = 1; s < L.world; ++s) {
      v = ld_cg_any<Pack<T, E>>(reinterpret_cast<const uint4 *>(reinterpret_cast<const T *>(mail_data(mine, parity, s)) + lo) + i);
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = R::combine(acc[e], R::load(v.v[e]));
    }
    Pack<T, E> r;
#pragma unroll
    for (int e = 0; e < E; ++e) r.v[e] = (T)acc[e];
    reinterpret_cast<Pack<T, E> *>(out + lo)[i] = r;
  }
  for (int64_t i = tail + threadIdx.x; i < hi; i += blockDim.x) {
    A acc = R::load(ld_cg_any<T>(reinterpret_cast<const T *>(mail_data(mine, parity, 0)) + i));
    for (int s = 1; s < L.world; ++s)
      acc = R::combine(acc, R::load(ld_cg_any<T>(reinterpret_cast<const T *>(mail_data(mine, parity, s)) + i)));
    out[i] = (T)acc;
  }
}
#endif  // __CUDACC__

// host side (comm.cu <-> reduce.cu)
int launch_allreduce(const CommLink &link, int red, int dtype, int64_t count, const void *local, void *out,
                     cudaStream_t stream);

}  // namespace nuk
