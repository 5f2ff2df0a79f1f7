// runtime.cu -- device selection, streams, events, storage, knobs, scratch arenas,
// and the broadcast/stride resolver of libnuk.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#include "common.cuh"

namespace nuk {

int reduce_1d_device(int red, int dtype, size_t n, const void *x, int64_t incx, void *out_dev, cudaStream_t stream,
                     int flags);   // reduce.cu

// ------------------------------------------------------------------ errors
static thread_local int tl_status = NUK_OK;
static thread_local char tl_msg[512] = "";

void set_error(int status, const char *fmt, ...) {
  tl_status = status;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_msg, sizeof(tl_msg), fmt, ap);
  va_end(ap);
  if (getenv("NUK_VERBOSE_ERRORS")) fprintf(stderr, "[libnuk] error %d: %s\n", status, tl_msg);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  const int st = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? NUK_ERR_NO_DEVICE
                                                                              : NUK_ERR_CUDA;
  set_error(st, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  return st;
}

// ------------------------------------------------------------------ device info
static std::mutex g_mu;
static std::vector<DeviceInfo> g_devs;  // indexed by device ordinal, device == -1 until probed
static thread_local int tl_device = -1;
static thread_local cudaStream_t tl_stream = nullptr;
static std::atomic<unsigned long long> g_launches{0};

static int probe_device(int dev, DeviceInfo *out) {
  cudaDeviceProp p;
  NUK_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major < 10) {
    set_error(NUK_ERR_NO_DEVICE, "device %d is sm_%d%d; libnuk is built for sm_100a only", dev,
              p.major, p.minor);
    return NUK_ERR_NO_DEVICE;
  }
  out->device = dev;
  out->sm_count = p.multiProcessorCount;
  out->l2_bytes = (size_t)p.l2CacheSize;
  return NUK_OK;
}

static int select_device(int dev) {
  int n = 0;
  NUK_CUDA(cudaGetDeviceCount(&n));
  if (n <= 0 || dev < 0 || dev >= n) {
    set_error(NUK_ERR_NO_DEVICE, "no CUDA device %d (count %d); libnuk has no CPU fallback", dev, n);
    return NUK_ERR_NO_DEVICE;
  }
  NUK_CUDA(cudaSetDevice(dev));
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if ((int)g_devs.size() < n) g_devs.resize(n, DeviceInfo{-1, 0, 0});
    if (g_devs[dev].device < 0) NUK_TRY(probe_device(dev, &g_devs[dev]));
  }
  tl_device = dev;
  return NUK_OK;
}

int use_device(int dev) {
  if (dev == tl_device) {
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur == dev) return NUK_OK;
  }
  return select_device(dev);
}

const DeviceInfo *device_info() {
  if (tl_device < 0) {
    int dev = 0;
    // adopt whatever device the thread already uses (e.g. torch set it), else 0
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    if (select_device(dev) != NUK_OK) return nullptr;
  }
  return &g_devs[tl_device];
}

cudaStream_t current_stream() { return tl_stream; }
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// NUK_LOG=1: one stderr line per kernel launch (kernel family, device, thread) -- the env-gated
// per-call log of SURVEY section 5; read once.
static int log_level() {
  static const int lv = [] {
    const char *v = getenv("NUK_LOG");
    return v ? atoi(v) : 0;
  }();
  return lv;
}

static bool nvtx_on() {
  static const bool on = [] {
    const char *v = getenv("NUK_NVTX");
    return v && atoi(v) != 0;
  }();
  return on;
}
ApiRange::ApiRange(const char *name) : on(nvtx_on()) {
  if (on) nvtxRangePushA(name);
}
ApiRange::~ApiRange() {
  if (on) nvtxRangePop();
}

int post_launch(const char *kernel) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(NUK_ERR_CUDA, "launch of %s failed: %s", kernel, cudaGetErrorString(e));
    return NUK_ERR_CUDA;
  }
  count_launch();
  if (log_level() > 0)
    fprintf(stderr, "[libnuk] launch %-20s dev %d launch# %llu\n", kernel, tl_device,
            (unsigned long long)g_launches.load(std::memory_order_relaxed));
  return NUK_OK;
}

// ------------------------------------------------------------------ knobs
// Plain atomics indexed by Opt: the launch path reads them without a lock or a string hash (the
// reference calls the C layer from concurrent lparallel workers, src/transcendental/lparallel.lisp:70-96).
static const char *const kOptNames[OPT_COUNT] = {"grid_mult", "reduce_grid_mult", "stage_mb", "stage_slots",
                                                 "col_split_rows", "seq_cols", "stage_threads", "rows_ctas_per_sm",
                                                 "stage_bounce", "rows_rpc", "pdl", "stage_nt"};
static std::atomic<long> g_opts[OPT_COUNT];
static std::once_flag g_opts_once;
static void init_opts() {
  g_opts[OPT_GRID_MULT] = 8;          // resident CTAs per SM targeted by the persistent nd map kernel
  g_opts[OPT_REDUCE_GRID_MULT] = 16;  // CTAs per SM in pass 1 of full reductions (profiles/reduce_grid_r01.txt)
  g_opts[OPT_STAGE_MB] = 32;          // chunk size of the host-pointer staging pipeline
  g_opts[OPT_STAGE_SLOTS] = 3;
  g_opts[OPT_COL_SPLIT_ROWS] = 256;   // rows per CTA in the split column reduction
  g_opts[OPT_SEQ_COLS] = 1;           // 1: TMA-fed sequential (reference-order) column reduction where it fits
  g_opts[OPT_STAGE_THREADS] = 0;      // host mover threads for pageable / strided host operands (0 = auto)
  g_opts[OPT_STAGE_BOUNCE] = 1;       // 0: hand pageable |inc| == 1 operands to the driver's own pageable copy (A/B only)
  g_opts[OPT_PDL] = 1;                // flat maps launch with programmatic stream serialization (map.cuh:launch_flat)
  g_opts[OPT_STAGE_NT] = 1;           // 1: non-temporal stores INTO the page-locked ring too (0: cached stores, for rings small enough to stay in L3)
  g_opts[OPT_ROWS_RPC] = 0;           // rows kernel: rows per CTA (0 = built-in rule)
  g_opts[OPT_ROWS_CTAS_PER_SM] = 0;   // rows kernel: CTAs per SM the row grouping aims for (0 = built-in rule)
  for (int i = 0; i < OPT_COUNT; ++i) {
    std::string env = "NUK_";
    for (const char *c = kOptNames[i]; *c; ++c) env.push_back((char)toupper(*c));
    if (const char *v = getenv(env.c_str())) g_opts[i] = atol(v);
  }
}
long opt(Opt o) {
  std::call_once(g_opts_once, init_opts);
  return g_opts[o].load(std::memory_order_relaxed);
}
static int opt_index(const char *name) {
  for (int i = 0; i < OPT_COUNT; ++i)
    if (name && !strcmp(name, kOptNames[i])) return i;
  return -1;
}

// ------------------------------------------------------------------ pointer classification
static thread_local int tl_ptr_mode = NUK_PTR_AUTO;
PtrKind classify_ptr(const void *p, bool strided) {
  if (!p) return PK_NULL;
  if (strided && tl_ptr_mode == NUK_PTR_DEVICE) return PK_DEVICE;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return PK_PAGEABLE;
  }
  switch (a.type) {
    case cudaMemoryTypeDevice:
    case cudaMemoryTypeManaged: return PK_DEVICE;
    case cudaMemoryTypeHost: return PK_HOST;
    default: return PK_PAGEABLE;
  }
}

int grid_for(size_t work_blocks, int blocks_per_sm) {
  const DeviceInfo *d = device_info();
  const size_t cap = (size_t)(d ? d->sm_count : 148) * (size_t)(blocks_per_sm > 0 ? blocks_per_sm : 1);
  size_t g = work_blocks < cap ? work_blocks : cap;
  return (int)(g == 0 ? 1 : g);
}

// ------------------------------------------------------------------ scratch arenas
// Everything a thread caches on a device is keyed by the DEVICE it was allocated on: a thread that
// moves between devices with nuk_init (the single-process multi-GPU mode does, once per call) must
// never hand a kernel a pointer into another device's memory.
struct Arena {
  void *ptr = nullptr;
  size_t bytes = 0;
};
struct ArenaKey {
  int device;
  cudaStream_t stream;
  bool operator==(const ArenaKey &o) const { return device == o.device && stream == o.stream; }
};
struct ArenaKeyHash {
  size_t operator()(const ArenaKey &k) const {
    return std::hash<const void *>()(k.stream) * 31u + (size_t)(k.device + 1);
  }
};
struct ThreadState {
  std::unordered_map<ArenaKey, Arena, ArenaKeyHash> arenas;
  std::vector<Arena> bufs;   // [device * BUF_COUNT + slot]
  void *pinned = nullptr;
  ~ThreadState() {
    // contexts may already be gone at thread exit; ignore errors
    for (auto &kv : arenas)
      if (kv.second.ptr) cudaFree(kv.second.ptr);
    for (auto &b : bufs)
      if (b.ptr) cudaFree(b.ptr);
    if (pinned) cudaFreeHost(pinned);
  }
};
static thread_local ThreadState tl_state;

// Every arena starts with a 256-byte header that is zero when a kernel starts: the arrival ticket of the
// single-launch full reductions (reduce.cu: finish_in_last_cta, which re-arms it).  scratch() hands out the
// memory after the header.
constexpr size_t kArenaHeader = 256;
static int arena_for(cudaStream_t s, size_t bytes, Arena **out) {
  const DeviceInfo *d = device_info();
  if (!d) return nuk_last_status();
  Arena &a = tl_state.arenas[ArenaKey{d->device, s}];
  if (a.bytes < bytes + kArenaHeader) {
    if (a.ptr) {
      NUK_CUDA(cudaStreamSynchronize(s));  // earlier kernels may still read the old arena
      NUK_CUDA(cudaFree(a.ptr));
      a.ptr = nullptr;
      a.bytes = 0;
    }
    size_t want = bytes + kArenaHeader < (1u << 20) ? (1u << 20) : bytes + kArenaHeader;
    NUK_CUDA(cudaMalloc(&a.ptr, want));
    NUK_CUDA(cudaMemsetAsync(a.ptr, 0, kArenaHeader, s));
    a.bytes = want;
  }
  *out = &a;
  return NUK_OK;
}
int scratch(cudaStream_t s, size_t bytes, void **dptr) {
  Arena *a = nullptr;
  NUK_TRY(arena_for(s, bytes, &a));
  *dptr = static_cast<char *>(a->ptr) + kArenaHeader;
  return NUK_OK;
}
int scratch_with_ticket(cudaStream_t s, size_t bytes, void **dptr, unsigned **ticket) {
  Arena *a = nullptr;
  NUK_TRY(arena_for(s, bytes, &a));
  *dptr = static_cast<char *>(a->ptr) + kArenaHeader;
  *ticket = static_cast<unsigned *>(a->ptr);
  return NUK_OK;
}

static void drop_arena(cudaStream_t s) {   // nuk_stream_destroy: a recycled handle must not inherit the arena
  for (auto it = tl_state.arenas.begin(); it != tl_state.arenas.end();) {
    if (it->first.stream == s) {
      if (it->second.ptr) cudaFree(it->second.ptr);
      it = tl_state.arenas.erase(it);
    } else {
      ++it;
    }
  }
}

int thread_buffer(BufSlot slot, size_t bytes, void **dptr) {
  const DeviceInfo *d = device_info();
  if (!d) return nuk_last_status();
  const size_t idx = (size_t)d->device * BUF_COUNT + (size_t)slot;
  if (tl_state.bufs.size() <= idx) tl_state.bufs.resize(idx + 1);
  Arena &a = tl_state.bufs[idx];
  if (a.bytes < bytes) {
    if (a.ptr) {
      NUK_CUDA(cudaDeviceSynchronize());
      NUK_CUDA(cudaFree(a.ptr));
      a.ptr = nullptr;
      a.bytes = 0;
    }
    NUK_CUDA(cudaMalloc(&a.ptr, bytes < 256 ? 256 : bytes));
    a.bytes = bytes < 256 ? 256 : bytes;
  }
  *dptr = a.ptr;
  return NUK_OK;
}

int pinned_slot(void **hptr) {
  if (!tl_state.pinned) NUK_CUDA(cudaHostAlloc(&tl_state.pinned, 512, cudaHostAllocPortable));
  *hptr = tl_state.pinned;
  return NUK_OK;
}

// ------------------------------------------------------------------ descriptor compaction
int coalesce(int nops, int rank, int64_t *dims, int64_t *strides) {
  // 1. drop size-1 axes
  int r = 0;
  for (int a = 0; a < rank; ++a) {
    if (dims[a] == 1) continue;
    dims[r] = dims[a];
    for (int o = 0; o < nops; ++o) strides[o * NUK_MAX_RANK + r] = strides[o * NUK_MAX_RANK + a];
    ++r;
  }
  rank = r;
  if (rank == 0) return 0;
  // 2. merge axis a+1 into axis a when stride[a] == stride[a+1] * dims[a+1] for every operand
  int w = 0;
  for (int a = 1; a < rank; ++a) {
    bool ok = true;
    for (int o = 0; o < nops && ok; ++o)
      ok = strides[o * NUK_MAX_RANK + w] == strides[o * NUK_MAX_RANK + a] * dims[a];
    if (ok) {
      dims[w] *= dims[a];
      for (int o = 0; o < nops; ++o) strides[o * NUK_MAX_RANK + w] = strides[o * NUK_MAX_RANK + a];
    } else {
      ++w;
      dims[w] = dims[a];
      for (int o = 0; o < nops; ++o) strides[o * NUK_MAX_RANK + w] = strides[o * NUK_MAX_RANK + a];
    }
  }
  return w + 1;
}

}  // namespace nuk

using namespace nuk;

// =================================================================== C ABI
NUK_API const char *nuk_last_error(void) { return tl_msg; }
NUK_API int nuk_last_status(void) { return tl_status; }
NUK_API void nuk_clear_error(void) {
  tl_status = NUK_OK;
  tl_msg[0] = 0;
}
NUK_API const char *nuk_version(void) { return "numericals-b200 libnuk 0.1 (sm_100a)"; }
NUK_API size_t nuk_dtype_size(int dtype) { return dtype_size(dtype); }

NUK_API int nuk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
NUK_API int nuk_init(int device) {
  NUK_TRY(select_device(device));
  NUK_CUDA(cudaFree(0));
  return NUK_OK;
}
NUK_API int nuk_current_device(void) { return device_info() ? tl_device : -1; }
NUK_API int nuk_sm_count(void) {
  const DeviceInfo *d = device_info();
  return d ? d->sm_count : 0;
}

NUK_API int nuk_stream_create(nuk_stream_t *s) {
  if (!device_info()) return nuk_last_status();
  cudaStream_t st;
  NUK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *s = st;
  return NUK_OK;
}
NUK_API int nuk_stream_destroy(nuk_stream_t s) {
  drop_arena((cudaStream_t)s);
  NUK_CUDA(cudaStreamDestroy((cudaStream_t)s));
  return NUK_OK;
}
NUK_API int nuk_stream_sync(nuk_stream_t s) {
  NUK_CUDA(cudaStreamSynchronize((cudaStream_t)s));
  return NUK_OK;
}
NUK_API int nuk_set_stream(nuk_stream_t s) {
  tl_stream = (cudaStream_t)s;
  return NUK_OK;
}
NUK_API nuk_stream_t nuk_get_stream(void) { return tl_stream; }
NUK_API int nuk_device_sync(void) {
  if (!device_info()) return nuk_last_status();
  NUK_CUDA(cudaDeviceSynchronize());
  return NUK_OK;
}

NUK_API int nuk_event_create(nuk_event_t *e) {
  if (!device_info()) return nuk_last_status();
  cudaEvent_t ev;
  NUK_CUDA(cudaEventCreate(&ev));
  *e = ev;
  return NUK_OK;
}
NUK_API int nuk_event_destroy(nuk_event_t e) {
  NUK_CUDA(cudaEventDestroy((cudaEvent_t)e));
  return NUK_OK;
}
NUK_API int nuk_event_record(nuk_event_t e, nuk_stream_t s) {
  NUK_CUDA(cudaEventRecord((cudaEvent_t)e, (cudaStream_t)s));
  return NUK_OK;
}
NUK_API int nuk_event_sync(nuk_event_t e) {
  NUK_CUDA(cudaEventSynchronize((cudaEvent_t)e));
  return NUK_OK;
}
NUK_API int nuk_event_elapsed_ms(nuk_event_t a, nuk_event_t b, float *ms) {
  NUK_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return NUK_OK;
}

NUK_API int nuk_malloc(void **dptr, size_t bytes) {
  if (!device_info()) return nuk_last_status();
  *dptr = nullptr;
  if (bytes == 0) bytes = 1;
  NUK_CUDA(cudaMalloc(dptr, bytes));
  return NUK_OK;
}
NUK_API int nuk_free(void *dptr) {
  if (dptr) NUK_CUDA(cudaFree(dptr));
  return NUK_OK;
}
NUK_API int nuk_memcpy_h2d(void *dst, const void *src, size_t bytes, nuk_stream_t s) {
  NUK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)s));
  return NUK_OK;
}
NUK_API int nuk_memcpy_d2h(void *dst, const void *src, size_t bytes, nuk_stream_t s) {
  NUK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)s));
  return NUK_OK;
}
NUK_API int nuk_memcpy_d2d(void *dst, const void *src, size_t bytes, nuk_stream_t s) {
  NUK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)s));
  return NUK_OK;
}
NUK_API int nuk_memset(void *dst, int byte, size_t bytes, nuk_stream_t s) {
  NUK_CUDA(cudaMemsetAsync(dst, byte, bytes, (cudaStream_t)s));
  return NUK_OK;
}
NUK_API int nuk_mem_info(size_t *free_bytes, size_t *total_bytes) {
  if (!device_info()) return nuk_last_status();
  NUK_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
  return NUK_OK;
}
NUK_API int nuk_host_alloc(void **hptr, size_t bytes) {
  if (!device_info()) return nuk_last_status();
  NUK_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return NUK_OK;
}
NUK_API int nuk_host_free(void *hptr) {
  if (hptr) NUK_CUDA(cudaFreeHost(hptr));
  return NUK_OK;
}
NUK_API int nuk_host_register(void *hptr, size_t bytes) {
  if (!device_info()) return nuk_last_status();
  NUK_CUDA(cudaHostRegister(hptr, bytes, cudaHostRegisterDefault));
  return NUK_OK;
}
NUK_API int nuk_host_unregister(void *hptr) {
  NUK_CUDA(cudaHostUnregister(hptr));
  return NUK_OK;
}
NUK_API int nuk_pointer_is_device(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? 1 : 0;
}
NUK_API int nuk_set_pointer_mode(int mode) {
  if (mode != NUK_PTR_AUTO && mode != NUK_PTR_DEVICE) {
    set_error(NUK_ERR_INVALID, "unknown pointer mode %d", mode);
    return NUK_ERR_INVALID;
  }
  tl_ptr_mode = mode;
  return NUK_OK;
}
NUK_API int nuk_get_pointer_mode(void) { return tl_ptr_mode; }

// ------------------------------------------------------------------ CUDA graphs
// A launch-bound inner loop (C1: 10^6-element maps, 3-4 us of GPU work per call) is captured once and
// replayed: every map / axis-reduction entry point is capture-safe (no allocation, no sync) once its
// scratch arena exists, i.e. after one ordinary warm-up call on the same stream.
NUK_API int nuk_graph_begin(nuk_stream_t s) {
  if (!device_info()) return nuk_last_status();
  if (!s) {
    set_error(NUK_ERR_INVALID, "graph capture needs a stream created with nuk_stream_create");
    return NUK_ERR_INVALID;
  }
  NUK_CUDA(cudaStreamBeginCapture((cudaStream_t)s, cudaStreamCaptureModeRelaxed));
  return NUK_OK;
}
NUK_API int nuk_graph_end(nuk_stream_t s, nuk_graph_t *g) {
  cudaGraph_t graph = nullptr;
  NUK_CUDA(cudaStreamEndCapture((cudaStream_t)s, &graph));
  cudaGraphExec_t exec = nullptr;
  cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
  *g = exec;
  return NUK_OK;
}
NUK_API int nuk_graph_launch(nuk_graph_t g, nuk_stream_t s) {
  NUK_CUDA(cudaGraphLaunch((cudaGraphExec_t)g, (cudaStream_t)s));
  return NUK_OK;
}
NUK_API int nuk_graph_destroy(nuk_graph_t g) {
  if (g) NUK_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)g));
  return NUK_OK;
}

NUK_API int nuk_set_option(const char *name, long value) {
  const int i = opt_index(name);
  if (i < 0) {
    set_error(NUK_ERR_INVALID, "unknown option %s", name ? name : "(null)");
    return NUK_ERR_INVALID;
  }
  opt((Opt)i);   // make sure the defaults / environment were applied first
  g_opts[i].store(value, std::memory_order_relaxed);
  return NUK_OK;
}
NUK_API long nuk_get_option(const char *name) {
  const int i = opt_index(name);
  return i < 0 ? -1 : opt((Opt)i);
}
NUK_API unsigned long long nuk_launch_count(void) { return g_launches.load(); }

// L2 flush (benchmark hygiene only): overwrite a buffer of 2x the L2 size, then READ it back.  The write alone
// leaves ~126 MB of DIRTY lines whose write-back is charged to whatever kernel runs next (a 128 MiB map then
// competes with as much eviction traffic as it moves itself: C3b at 4096^2 measured 0.59 of the HBM peak for
// that reason); after the read pass the cache holds clean lines only, which are dropped for free.
NUK_API int nuk_flush_l2(nuk_stream_t s) {
  const DeviceInfo *d = device_info();
  if (!d) return nuk_last_status();
  static std::mutex mu;
  static std::unordered_map<int, void *> bufs;
  const size_t bytes = 2 * (d->l2_bytes ? d->l2_bytes : (size_t)128 << 20);
  void *buf = nullptr;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = bufs.find(d->device);
    if (it == bufs.end()) {
      NUK_CUDA(cudaMalloc(&buf, bytes + 256));
      bufs[d->device] = buf;
    } else {
      buf = it->second;
    }
  }
  static thread_local int flip = 0;
  NUK_CUDA(cudaMemsetAsync(buf, (flip++) & 0xff, bytes, (cudaStream_t)s));
  return reduce_1d_device(NUK_SUM, NUK_I32, bytes / 4, buf, 1, static_cast<char *>(buf) + bytes, (cudaStream_t)s, 0);
}

// ------------------------------------------------------------------ resolver
// %broadcast-compatible-p + strides-for-broadcast (src/utils/broadcast.lisp:3-63):
// right-aligned; each axis pair equal or one of them 1; stride 0 on broadcast axes.
NUK_API int nuk_broadcast_resolve(int nargs, const int *ranks, const int64_t *const *dims,
                                  const int64_t *const *strides, int *out_rank,
                                  int64_t *out_dims, int64_t *out_strides) {
  int R = 0;
  for (int i = 0; i < nargs; ++i) {
    if (ranks[i] < 0 || ranks[i] > NUK_MAX_RANK) {
      set_error(NUK_ERR_INVALID, "rank %d of argument %d exceeds NUK_MAX_RANK", ranks[i], i);
      return NUK_ERR_INVALID;
    }
    if (ranks[i] > R) R = ranks[i];
  }
  for (int a = 0; a < R; ++a) out_dims[a] = 1;
  for (int i = 0; i < nargs; ++i) {
    const int off = R - ranks[i];
    for (int a = 0; a < ranks[i]; ++a) {
      const int64_t d = dims[i][a];
      int64_t &o = out_dims[off + a];
      if (o == 1) o = d;
      else if (d != 1 && d != o) {
        set_error(NUK_ERR_INVALID,
                  "incompatible-broadcast-dimensions: axis %d of argument %d has %lld, expected 1 or %lld",
                  a, i, (long long)d, (long long)o);
        return NUK_ERR_INVALID;
      }
    }
  }
  for (int i = 0; i < nargs; ++i) {
    const int off = R - ranks[i];
    int64_t *os = out_strides + (size_t)i * NUK_MAX_RANK;
    int64_t rm = 1;  // row-major stride when strides[i] == NULL
    for (int a = R - 1; a >= 0; --a) {
      if (a < off) { os[a] = 0; continue; }
      const int64_t d = dims[i][a - off];
      const int64_t st = (strides && strides[i]) ? strides[i][a - off] : rm;
      os[a] = (d == 1 && out_dims[a] != 1) ? 0 : (d == 1 ? 0 : st);
      rm *= d;
    }
  }
  *out_rank = R;
  return NUK_OK;
}

NUK_API int nuk_coalesce(int nops, int rank, int64_t *dims, int64_t *strides) {
  if (rank < 0 || rank > NUK_MAX_RANK || nops < 1 || nops > 3) return -1;
  return coalesce(nops, rank, dims, strides);
}
