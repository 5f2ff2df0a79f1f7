// runtime.cu -- device selection, streams, events, storage, knobs, scratch arenas,
// and the broadcast/stride resolver of libnuk.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include "common.cuh"

namespace nuk {

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

int post_launch(const char *kernel) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(NUK_ERR_CUDA, "launch of %s failed: %s", kernel, cudaGetErrorString(e));
    return NUK_ERR_CUDA;
  }
  count_launch();
  return NUK_OK;
}

// ------------------------------------------------------------------ knobs
static std::unordered_map<std::string, long> &options() {
  static std::unordered_map<std::string, long> o = [] {
    std::unordered_map<std::string, long> m;
    m["grid_mult"] = 8;        // resident CTAs per SM targeted by persistent map kernels
    m["reduce_grid_mult"] = 16; // CTAs per SM in pass 1 of full reductions (16: sum/max f32/f64 +7 % over 8, profiles/reduce_grid_r01.txt)
    m["stage_mb"] = 32;        // chunk size of the host-pointer staging pipeline
    m["stage_slots"] = 3;
    m["col_split_rows"] = 256; // rows per CTA in the split column reduction
    m["seq_cols"] = 1;         // 1: TMA-fed sequential (reference-order) column reduction where it fits
    const char *names[] = {"grid_mult", "reduce_grid_mult", "stage_mb", "stage_slots", "col_split_rows",
                           "seq_cols"};
    for (const char *n : names) {
      std::string env = "NUK_";
      for (const char *c = n; *c; ++c) env.push_back((char)toupper(*c));
      if (const char *v = getenv(env.c_str())) m[n] = atol(v);
    }
    return m;
  }();
  return o;
}
long option(const char *name) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto &o = options();
  auto it = o.find(name);
  return it == o.end() ? -1 : it->second;
}

int grid_for(size_t work_blocks, int blocks_per_sm) {
  const DeviceInfo *d = device_info();
  const size_t cap = (size_t)(d ? d->sm_count : 148) * (size_t)(blocks_per_sm > 0 ? blocks_per_sm : 1);
  size_t g = work_blocks < cap ? work_blocks : cap;
  return (int)(g == 0 ? 1 : g);
}

// ------------------------------------------------------------------ scratch arenas
struct Arena {
  void *ptr = nullptr;
  size_t bytes = 0;
};
struct ThreadState {
  std::unordered_map<cudaStream_t, Arena> arenas;
  void *pinned = nullptr;
  ~ThreadState() {
    // contexts may already be gone at thread exit; ignore errors
    for (auto &kv : arenas)
      if (kv.second.ptr) cudaFree(kv.second.ptr);
    if (pinned) cudaFreeHost(pinned);
  }
};
static thread_local ThreadState tl_state;

int scratch(cudaStream_t s, size_t bytes, void **dptr) {
  Arena &a = tl_state.arenas[s];
  if (a.bytes < bytes) {
    if (a.ptr) {
      NUK_CUDA(cudaStreamSynchronize(s));  // earlier kernels may still read the old arena
      NUK_CUDA(cudaFree(a.ptr));
      a.ptr = nullptr;
      a.bytes = 0;
    }
    size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
    NUK_CUDA(cudaMalloc(&a.ptr, want));
    a.bytes = want;
  }
  *dptr = a.ptr;
  return NUK_OK;
}

int pinned_slot(void **hptr) {
  if (!tl_state.pinned) NUK_CUDA(cudaHostAlloc(&tl_state.pinned, 512, cudaHostAllocDefault));
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

NUK_API int nuk_set_option(const char *name, long value) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto &o = options();
  auto it = o.find(name);
  if (it == o.end()) {
    set_error(NUK_ERR_INVALID, "unknown option %s", name);
    return NUK_ERR_INVALID;
  }
  it->second = value;
  return NUK_OK;
}
NUK_API long nuk_get_option(const char *name) { return option(name); }
NUK_API unsigned long long nuk_launch_count(void) { return g_launches.load(); }

// L2 flush: overwrite a buffer of 2x the L2 size (benchmark hygiene only)
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
      NUK_CUDA(cudaMalloc(&buf, bytes));
      bufs[d->device] = buf;
    } else {
      buf = it->second;
    }
  }
  static thread_local int flip = 0;
  NUK_CUDA(cudaMemsetAsync(buf, (flip++) & 0xff, bytes, (cudaStream_t)s));
  return NUK_OK;
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
