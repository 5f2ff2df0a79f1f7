// tools/c1_latency.cpp -- BASELINE config C1 (nu:add on two contiguous double-float arrays of 10^6 elements,
// benchmarks/two-arg-fn.lisp:50-81) as a C caller sees it: microseconds per BMAS_dadd call through the C ABI of
// libnuk.so with no Python in the way -- host issue time per call, device time per call (events around a
// back-to-back batch), with the driver pointer query (NUK_PTR_AUTO), without it (NUK_PTR_DEVICE, the
// device-storage contract) and as a CUDA-graph replay.
//   g++ -O2 -std=c++17 -Iinclude tools/c1_latency.cpp -o tools/_build/c1_latency -ldl
//   tools/_build/c1_latency numericals_b200/lib/libnuk.so
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <vector>
#include "nuk.h"

#define SYM(name) auto p_##name = reinterpret_cast<decltype(&name)>(dlsym(h, #name)); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }

int main(int argc, char **argv) {
  void *h = dlopen(argc > 1 ? argv[1] : "numericals_b200/lib/libnuk.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  SYM(nuk_init) SYM(nuk_malloc) SYM(nuk_memcpy_h2d) SYM(nuk_memcpy_d2h) SYM(nuk_stream_create) SYM(nuk_set_stream)
  SYM(nuk_stream_sync) SYM(nuk_event_create) SYM(nuk_event_record) SYM(nuk_event_sync) SYM(nuk_event_elapsed_ms)
  SYM(nuk_set_pointer_mode) SYM(nuk_graph_begin) SYM(nuk_graph_end) SYM(nuk_graph_launch) SYM(nuk_last_status)
  SYM(nuk_last_error) SYM(BMAS_dadd) SYM(nuk_flush_l2) SYM(nuk_set_option)
  if (p_nuk_init(0) != 0) { fprintf(stderr, "%s\n", p_nuk_last_error()); return 1; }
  const long n = 1000000;
  std::vector<double> ones(n, 1.0), out(n, 0.0);
  void *a, *b, *o;
  p_nuk_malloc(&a, n * 8); p_nuk_malloc(&b, n * 8); p_nuk_malloc(&o, n * 8);
  nuk_stream_t s;
  p_nuk_stream_create(&s);
  p_nuk_set_stream(s);
  p_nuk_memcpy_h2d(a, ones.data(), n * 8, s);
  p_nuk_memcpy_h2d(b, ones.data(), n * 8, s);
  nuk_event_t e0, e1;
  p_nuk_event_create(&e0); p_nuk_event_create(&e1);
  auto call = [&] { p_BMAS_dadd(n, (double *)a, 1, (double *)b, 1, (double *)o, 1); };
  auto batch = [&](const char *label, int reps) {
    for (int i = 0; i < 50; ++i) call();
    p_nuk_stream_sync(s);
    p_nuk_event_record(e0, s);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) call();
    const auto t1 = std::chrono::steady_clock::now();
    p_nuk_event_record(e1, s);
    p_nuk_event_sync(e1);
    float ms = 0;
    p_nuk_event_elapsed_ms(e0, e1, &ms);
    printf("%-40s host issue %6.2f us/call   device %6.2f us/call   (%.0f GB/s algorithmic, L2-resident)\n", label,
           std::chrono::duration<double, std::micro>(t1 - t0).count() / reps, ms * 1e3 / reps, 24e6 / (ms * 1e-3 / reps) / 1e9);
  };
  batch("BMAS_dadd, NUK_PTR_AUTO", 2000);
  p_nuk_set_pointer_mode(NUK_PTR_DEVICE);
  batch("BMAS_dadd, NUK_PTR_DEVICE", 2000);
  p_nuk_set_option("pdl", 0);
  batch("  same, ordinary launches (pdl = 0)", 2000);
  p_nuk_set_option("pdl", 1);
  // graph: 20 calls captured once, replayed
  call();
  p_nuk_graph_begin(s);
  for (int i = 0; i < 20; ++i) call();
  nuk_graph_t g;
  if (p_nuk_graph_end(s, &g) != 0) { fprintf(stderr, "%s\n", p_nuk_last_error()); return 1; }
  p_nuk_graph_launch(g, s);
  p_nuk_stream_sync(s);
  p_nuk_event_record(e0, s);
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < 100; ++i) p_nuk_graph_launch(g, s);
  const auto t1 = std::chrono::steady_clock::now();
  p_nuk_event_record(e1, s);
  p_nuk_event_sync(e1);
  float ms = 0;
  p_nuk_event_elapsed_ms(e0, e1, &ms);
  printf("%-40s host issue %6.2f us/call   device %6.2f us/call   (%.0f GB/s algorithmic, L2-resident)\n",
         "graph replay (20 calls per graph)", std::chrono::duration<double, std::micro>(t1 - t0).count() / 2000, ms * 1e3 / 2000,
         24e6 / (ms * 1e-3 / 2000) / 1e9);
  // one isolated call, cold L2
  float tot = 0;
  for (int i = 0; i < 20; ++i) {
    p_nuk_flush_l2(s);
    p_nuk_event_record(e0, s);
    call();
    p_nuk_event_record(e1, s);
    p_nuk_event_sync(e1);
    p_nuk_event_elapsed_ms(e0, e1, &ms);
    tot += ms;
  }
  printf("%-40s device %6.2f us/call   (%.0f GB/s algorithmic, from HBM)\n", "single call, L2 flushed", tot * 1e3 / 20,
         24e6 / (tot * 1e-3 / 20) / 1e9);
  p_nuk_memcpy_d2h(out.data(), o, n * 8, s);
  p_nuk_stream_sync(s);
  for (long i = 0; i < n; i += 9973) if (out[i] != 2.0) { fprintf(stderr, "wrong result at %ld: %g\n", i, out[i]); return 1; }
  printf("status %d\n", p_nuk_last_status());
  return 0;
}
