// comm.cu -- multi-GPU communicators of libnuk (host side; the kernels are in comm.cuh / reduce.cu).
//
// Sharding rule = the reference's thread-slicing rule lifted to devices: contiguous slabs of
// ceiling(n / workers) units, the last worker takes what is left (src/transcendental/lparallel.lisp:60-96);
// for an nd array the FIRST axis that is long enough is the one that is cut
// (dense-numericals-src/utils/lparallel.lisp:41-49).  Maps and reductions that keep the cut axis need no
// communication; what does is the combine step of a reduction over the cut axis, served here by mailboxes
// in peer-mapped HBM (comm.cuh).  Two ways to build a communicator:
//   nuk_comm_init_all   one process drives all GPUs (what a Common Lisp host does: one image, N devices)
//   nuk_comm_create / nuk_comm_export / nuk_comm_connect   one process per GPU (torchrun-style); the
//                       128-byte handles travel through whatever the host already has (torch.distributed,
//                       a file, a socket)
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unistd.h>
#include "comm.cuh"
#include "dispatch.cuh"

struct nuk_comm {
  int world = 0, rank = 0, device = -1;
  unsigned seq = 0;                       // collectives issued so far (same on every rank by contract)
  char *mailbox = nullptr;                // mine
  char *mail[nuk::kCommMaxWorld] = {};    // every rank's mailbox as seen from this device
  bool ipc_opened[nuk::kCommMaxWorld] = {};
  void *partials = nullptr;               // pass-1 partials of sharded full reductions
  size_t partial_bytes = 0;
  int *status_host = nullptr;             // pinned + mapped: kernels raise it on a wait timeout
  int *status_dev = nullptr;
  unsigned long long timeout_ns = 0;
  bool connected = false;
};

namespace nuk {
namespace {

struct WireHandle {              // what nuk_comm_export writes (<= NUK_COMM_HANDLE_BYTES)
  unsigned magic;
  int rank, world, device;
  long pid;
  cudaIpcMemHandle_t mem;
};
static_assert(sizeof(WireHandle) <= NUK_COMM_HANDLE_BYTES, "wire handle too large");
constexpr unsigned kMagic = 0x6e756b63u;  // "nukc"

struct DeviceGuard {             // communicator calls run on the communicator's device, then restore
  int prev = -1;
  bool changed = false;
  int enter(int dev) {
    NUK_CUDA(cudaGetDevice(&prev));
    if (prev != dev) {
      NUK_TRY(use_device(dev));
      changed = true;
    }
    return NUK_OK;
  }
  ~DeviceGuard() {
    if (changed) use_device(prev);
  }
};

int alloc_comm(int world, int rank, int device, nuk_comm **out) {
  if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world) {
    set_error(NUK_ERR_INVALID, "communicator of %d ranks (rank %d): at most %d GPUs of one box", world, rank,
              kCommMaxWorld);
    return NUK_ERR_INVALID;
  }
  nuk_comm *c = new nuk_comm;
  c->world = world;
  c->rank = rank;
  c->device = device;
  const char *t = getenv("NUK_COMM_TIMEOUT_MS");
  c->timeout_ns = (unsigned long long)(t ? atoll(t) : 20000) * 1000000ull;
  cudaError_t e;
  if ((e = cudaMalloc(&c->mailbox, kCommMailboxBytes)) != cudaSuccess ||
      (e = cudaMemset(c->mailbox, 0, kCommFlagBytes)) != cudaSuccess ||
      (e = cudaMalloc(&c->partials, (size_t)1 << 20)) != cudaSuccess ||
      (e = cudaHostAlloc(&c->status_host, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable)) != cudaSuccess ||
      (e = cudaHostGetDevicePointer(&c->status_dev, c->status_host, 0)) != cudaSuccess) {
    if (c->mailbox) cudaFree(c->mailbox);
    if (c->partials) cudaFree(c->partials);
    if (c->status_host) cudaFreeHost(c->status_host);
    delete c;
    return cuda_fail(e, "communicator allocation", __FILE__, __LINE__);
  }
  *c->status_host = 0;
  c->partial_bytes = (size_t)1 << 20;
  c->mail[rank] = c->mailbox;
  NUK_CUDA(cudaDeviceSynchronize());
  *out = c;
  return NUK_OK;
}

int check_ready(nuk_comm *c) {
  if (!c || !c->connected) {
    set_error(NUK_ERR_INVALID, "communicator is not connected");
    return NUK_ERR_INVALID;
  }
  if (*reinterpret_cast<volatile int *>(c->status_host) != 0) {
    set_error(NUK_ERR_COMM, "a peer did not arrive within %llu ms (rank %d of %d)",
              c->timeout_ns / 1000000ull, c->rank, c->world);
    return NUK_ERR_COMM;
  }
  return NUK_OK;
}

CommLink next_link(nuk_comm *c) {
  CommLink l;
  for (int r = 0; r < kCommMaxWorld; ++r) l.mail[r] = c->mail[r];
  l.world = c->world;
  l.rank = c->rank;
  l.seq = ++c->seq;
  l.status = c->status_dev;
  l.timeout_ns = c->timeout_ns;
  return l;
}

}  // namespace
}  // namespace nuk

using namespace nuk;

// ------------------------------------------------------------------ sharding rule
NUK_API int nuk_shard_bounds(int64_t n, int world, int rank, int64_t *lo, int64_t *hi) {
  if (n < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) {
    set_error(NUK_ERR_INVALID, "nuk_shard_bounds: bad arguments");
    return NUK_ERR_INVALID;
  }
  const int64_t per = (n + world - 1) / world;
  int64_t l = (int64_t)rank * per;
  if (l > n) l = n;
  int64_t h = (rank == world - 1) ? n : l + per;
  if (h > n) h = n;
  *lo = l;
  *hi = h;
  return NUK_OK;
}

NUK_API int nuk_shard_axis(int rank, const int64_t *dims, int world) {
  for (int a = 0; a < rank; ++a)
    if (dims[a] >= world) return a;
  return -1;
}

// ------------------------------------------------------------------ construction
NUK_API int nuk_comm_init_all(int ndev, const int *devices, nuk_comm_t *comms) {
  if (!comms || ndev < 1 || ndev > kCommMaxWorld) {
    set_error(NUK_ERR_INVALID, "nuk_comm_init_all: 1..%d devices", kCommMaxWorld);
    return NUK_ERR_INVALID;
  }
  const int have = nuk_device_count();
  int devs[kCommMaxWorld];
  for (int i = 0; i < ndev; ++i) {
    devs[i] = devices ? devices[i] : i;
    if (devs[i] < 0 || devs[i] >= have) {
      set_error(NUK_ERR_NO_DEVICE, "nuk_comm_init_all: device %d of %d", devs[i], have);
      return NUK_ERR_NO_DEVICE;
    }
  }
  int prev = 0;
  cudaGetDevice(&prev);
  int status = NUK_OK;
  for (int i = 0; i < ndev; ++i) comms[i] = nullptr;
  for (int i = 0; i < ndev && status == NUK_OK; ++i) {
    status = nuk_init(devs[i]);
    if (status == NUK_OK) status = alloc_comm(ndev, i, devs[i], &comms[i]);
  }
  // every device maps every other device's memory (NVLink peer access)
  for (int i = 0; i < ndev && status == NUK_OK; ++i) {
    cudaSetDevice(devs[i]);
    for (int j = 0; j < ndev && status == NUK_OK; ++j) {
      if (i == j || devs[i] == devs[j]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devs[i], devs[j]);
      if (!can) {
        set_error(NUK_ERR_COMM, "device %d cannot access device %d's memory", devs[i], devs[j]);
        status = NUK_ERR_COMM;
        break;
      }
      cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) status = cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
    }
  }
  if (status == NUK_OK) {
    for (int i = 0; i < ndev; ++i) {
      for (int j = 0; j < ndev; ++j) comms[i]->mail[j] = comms[j]->mailbox;
      comms[i]->connected = true;
    }
  } else {
    for (int i = 0; i < ndev; ++i)
      if (comms[i]) { nuk_comm_destroy(comms[i]); comms[i] = nullptr; }
  }
  nuk_init(prev);
  return status;
}

NUK_API int nuk_comm_create(int world, int rank, nuk_comm_t *comm) {
  const DeviceInfo *d = device_info();
  if (!d) return nuk_last_status();
  if (!comm) {
    set_error(NUK_ERR_INVALID, "nuk_comm_create: null handle");
    return NUK_ERR_INVALID;
  }
  nuk_comm *c = nullptr;
  NUK_TRY(alloc_comm(world, rank, d->device, &c));
  if (world == 1) c->connected = true;
  *comm = c;
  return NUK_OK;
}

NUK_API int nuk_comm_export(nuk_comm_t c, void *handle_out) {
  if (!c || !handle_out) {
    set_error(NUK_ERR_INVALID, "nuk_comm_export: null argument");
    return NUK_ERR_INVALID;
  }
  WireHandle w;
  memset(&w, 0, sizeof(w));
  w.magic = kMagic;
  w.rank = c->rank;
  w.world = c->world;
  w.device = c->device;
  w.pid = (long)getpid();
  NUK_CUDA(cudaIpcGetMemHandle(&w.mem, c->mailbox));
  memset(handle_out, 0, NUK_COMM_HANDLE_BYTES);
  memcpy(handle_out, &w, sizeof(w));
  return NUK_OK;
}

NUK_API int nuk_comm_connect(nuk_comm_t c, const void *all_handles) {
  if (!c || !all_handles) {
    set_error(NUK_ERR_INVALID, "nuk_comm_connect: null argument");
    return NUK_ERR_INVALID;
  }
  DeviceGuard g;
  NUK_TRY(g.enter(c->device));
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    WireHandle w;
    memcpy(&w, static_cast<const char *>(all_handles) + (size_t)r * NUK_COMM_HANDLE_BYTES, sizeof(w));
    if (w.magic != kMagic || w.rank != r || w.world != c->world) {
      set_error(NUK_ERR_INVALID, "nuk_comm_connect: handle %d is not rank %d of a %d-rank communicator", r, r,
                c->world);
      return NUK_ERR_INVALID;
    }
    if (w.pid == (long)getpid()) {
      set_error(NUK_ERR_INVALID, "nuk_comm_connect: rank %d lives in this process; use nuk_comm_init_all", r);
      return NUK_ERR_INVALID;
    }
    void *p = nullptr;
    NUK_CUDA(cudaIpcOpenMemHandle(&p, w.mem, cudaIpcMemLazyEnablePeerAccess));
    c->mail[r] = static_cast<char *>(p);
    c->ipc_opened[r] = true;
  }
  c->connected = true;
  return NUK_OK;
}

NUK_API int nuk_comm_destroy(nuk_comm_t c) {
  if (!c) return NUK_OK;
  DeviceGuard g;
  g.enter(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->ipc_opened[r] && c->mail[r]) cudaIpcCloseMemHandle(c->mail[r]);
  if (c->mailbox) cudaFree(c->mailbox);
  if (c->partials) cudaFree(c->partials);
  if (c->status_host) cudaFreeHost(c->status_host);
  cudaGetLastError();
  delete c;
  return NUK_OK;
}

NUK_API int nuk_comm_world(nuk_comm_t c) { return c ? c->world : 0; }
NUK_API int nuk_comm_rank(nuk_comm_t c) { return c ? c->rank : -1; }
NUK_API int nuk_comm_device(nuk_comm_t c) { return c ? c->device : -1; }
NUK_API int nuk_comm_status(nuk_comm_t c) {
  if (!c) return NUK_ERR_INVALID;
  return check_ready(c);
}

// ------------------------------------------------------------------ collectives
NUK_API int nuk_comm_allreduce(nuk_comm_t c, int red, int dtype, int64_t count, const void *local, void *out,
                               nuk_stream_t s) {
  ApiRange range("nuk/comm_allreduce");
  NUK_TRY(check_ready(c));
  const size_t sz = dtype_size(dtype);
  if (!sz || count < 0 || (count > 0 && (!local || !out))) {
    set_error(NUK_ERR_INVALID, "nuk_comm_allreduce: bad arguments (dtype %d, count %lld)", dtype, (long long)count);
    return NUK_ERR_INVALID;
  }
  DeviceGuard g;
  NUK_TRY(g.enter(c->device));
  if (!device_info()) return nuk_last_status();
  const int64_t per_round = (int64_t)(kCommSlotBytes / sz);
  for (int64_t off = 0; off < count; off += per_round) {
    const int64_t m = (count - off < per_round) ? (count - off) : per_round;
    const CommLink link = next_link(c);
    NUK_TRY(launch_allreduce(link, red, dtype, m, static_cast<const char *>(local) + (size_t)off * sz,
                             static_cast<char *>(out) + (size_t)off * sz, (cudaStream_t)s));
  }
  return NUK_OK;
}

// full reduction of this rank's slab, combined across ranks in rank order: pass 1 over the slab + ONE kernel
// that folds the partials, exchanges the value over peer memory and folds the W values (comm.cuh)
static int comm_reduce_entry(nuk_comm_t c, int red, int dtype, int rank, const int64_t *dims, const int64_t *sx,
                             const void *x, void *out, void *out2, nuk_stream_t s) {
  ApiRange range("nuk/comm_reduce");
  NUK_TRY(check_ready(c));
  if (rank < 0 || rank > NUK_MAX_RANK || !x || !out) {
    set_error(NUK_ERR_INVALID, "nuk_comm_reduce: bad descriptor (rank %d)", rank);
    return NUK_ERR_INVALID;
  }
  DeviceGuard g;
  NUK_TRY(g.enter(c->device));
  if (!device_info()) return nuk_last_status();
  ReduceProblem p;
  memset(&p, 0, sizeof(p));
  p.red = red;
  p.dtype = dtype;
  p.axis = -1;
  p.x = x;
  p.out = out;
  p.out2 = out2;
  p.stream = (cudaStream_t)s;
  p.flags = (c->rank == 0) ? 0 : NUK_REDUCE_NO_SEED;   // the global first element lives on rank 0
  int64_t d[NUK_MAX_RANK], st[2 * NUK_MAX_RANK];
  for (int a = 0; a < rank; ++a) {
    d[a] = dims[a];
    st[a] = sx[a];
  }
  const int r = coalesce(1, rank, d, st);
  p.rank = r;
  for (int a = 0; a < r; ++a) {
    p.dims[a] = d[a];
    p.sx[a] = st[a];
  }
  for (int a = 0; a < rank; ++a)
    if (dims[a] == 0) { p.rank = 1; p.dims[0] = 0; p.sx[0] = 1; }   // an empty slab still takes part
  const CommLink link = next_link(c);
  p.link = &link;
  p.link_partials = c->partials;
  p.link_partial_bytes = c->partial_bytes;
  return launch_reduce(p);
}

NUK_API int nuk_comm_reduce(nuk_comm_t c, int red, int dtype, int rank, const int64_t *dims, const int64_t *sx,
                            const void *x, void *out, nuk_stream_t s) {
  return comm_reduce_entry(c, red, dtype, rank, dims, sx, x, out, nullptr, s);
}

NUK_API int nuk_comm_moments(nuk_comm_t c, int dtype, int rank, const int64_t *dims, const int64_t *sx,
                             const void *x, void *sum_out, void *sumsq_out, nuk_stream_t s) {
  if (!sumsq_out) {
    set_error(NUK_ERR_INVALID, "nuk_comm_moments: null sumsq_out");
    return NUK_ERR_INVALID;
  }
  return comm_reduce_entry(c, NUK_SUM, dtype, rank, dims, sx, x, sum_out, sumsq_out, s);
}
