// tools/membench.cu -- streaming-kernel design-space probe for B200 (sm_100a): which access shape
// gets an elementwise map closest to the HBM ceiling?  Variants: vector width (128 / 256 bit),
// cache policy (default / .cs streaming / .nc+no_allocate), unroll, threads per CTA, persistent
// grid multiple vs one tile per CTA.  Prints GB/s (algorithmic bytes / CUDA-event time) for a
// unary map o = 2x (1 read + 1 write) and a binary map o = x + y (2 reads + 1 write), next to
// cudaMemcpyAsync D2D.  Results recorded in profiles/membench_r01.txt drive the choices in map.cuh.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

enum Policy { P_DEFAULT = 0, P_CS = 1, P_NC = 2 };

template <int POL> __device__ __forceinline__ float4 ld128(const float4 *p) {
  float4 v;
  if (POL == P_CS) asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else if (POL == P_NC) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
template <int POL> __device__ __forceinline__ void st128(float4 *p, float4 v) {
  if (POL == P_CS) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
  else asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
struct f8 { float v[8]; };
template <int POL> __device__ __forceinline__ f8 ld256(const f8 *p) {
  f8 r;
  if (POL == P_CS)
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  else
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}
template <int POL> __device__ __forceinline__ void st256(f8 *p, const f8 &r) {
  if (POL == P_CS)
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]));
  else
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]));
}

// persistent grid-stride over tiles of THREADS*UNROLL vectors; ARITY 1: o = 2x, 2: o = x + y
template <int ARITY, int LPOL, int SPOL, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) k128(size_t nvec, const float4 *x, const float4 *y, float4 *o) {
  const size_t tile = (size_t)THREADS * UNROLL, ntiles = nvec / tile;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const size_t base = t * tile + threadIdx.x;
    float4 a[UNROLL], b[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) a[u] = ld128<LPOL>(x + base + (size_t)u * THREADS);
    if (ARITY == 2) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) b[u] = ld128<LPOL>(y + base + (size_t)u * THREADS);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float4 r;
      if (ARITY == 2) { r.x = a[u].x + b[u].x; r.y = a[u].y + b[u].y; r.z = a[u].z + b[u].z; r.w = a[u].w + b[u].w; }
      else { r.x = a[u].x * 2; r.y = a[u].y * 2; r.z = a[u].z * 2; r.w = a[u].w * 2; }
      st128<SPOL>(o + base + (size_t)u * THREADS, r);
    }
  }
}
template <int ARITY, int LPOL, int SPOL, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) k256(size_t nvec, const f8 *x, const f8 *y, f8 *o) {
  const size_t tile = (size_t)THREADS * UNROLL, ntiles = nvec / tile;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const size_t base = t * tile + threadIdx.x;
    f8 a[UNROLL], b[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) a[u] = ld256<LPOL>(x + base + (size_t)u * THREADS);
    if (ARITY == 2) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) b[u] = ld256<LPOL>(y + base + (size_t)u * THREADS);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      f8 r;
#pragma unroll
      for (int k = 0; k < 8; ++k) r.v[k] = (ARITY == 2) ? a[u].v[k] + b[u].v[k] : a[u].v[k] * 2;
      st256<SPOL>(o + base + (size_t)u * THREADS, r);
    }
  }
}
// read-only stream (reduction-like): sum into one value per thread
template <int LPOL, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) kread(size_t nvec, const float4 *x, float *o) {
  const size_t tile = (size_t)THREADS * UNROLL, ntiles = nvec / tile;
  float s = 0;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const size_t base = t * tile + threadIdx.x;
    float4 a[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) a[u] = ld128<LPOL>(x + base + (size_t)u * THREADS);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) s += a[u].x + a[u].y + a[u].z + a[u].w;
  }
  if (s == 12345.678f) o[0] = s;
}

// transcendental maps from the product header, one tile per CTA: (threads, unroll) probe
#include "../numericals_b200/csrc/nukmath_f32.cuh"
struct FExp { __device__ static float f(float x) { return nukmath::exp_f32(x); } static const char *name() { return "exp"; } };
struct FLog { __device__ static float f(float x) { return nukmath::log_f32(x); } static const char *name() { return "log"; } };
struct FSin { __device__ static float f(float x) { return nukmath::sin_f32(x); } static const char *name() { return "sin"; } };
struct FTanh { __device__ static float f(float x) { return nukmath::tanh_f32(x); } static const char *name() { return "tanh"; } };
template <class F, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) kfun(size_t nvec, const float4 *x, float4 *o) {
  const size_t base = (size_t)blockIdx.x * THREADS * UNROLL + threadIdx.x;
  float4 a[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) a[u] = ld128<P_CS>(x + base + (size_t)u * THREADS);
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    float4 r;
    r.x = F::f(a[u].x); r.y = F::f(a[u].y); r.z = F::f(a[u].z); r.w = F::f(a[u].w);
    st128<P_CS>(o + base + (size_t)u * THREADS, r);
  }
}

static float *X, *Y, *O;
static size_t N;
static cudaEvent_t e0, e1;
static int SMS;

template <class L> static double run(L launch, int reps = 10) {
  launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}
static void report(const char *name, int arity, double ms) {
  const double bytes = (arity + 1) * 4.0 * N;
  printf("%-64s %8.4f ms %8.1f GB/s\n", name, ms, bytes / (ms * 1e-3) / 1e9);
  fflush(stdout);
}

template <int ARITY, int LPOL, int SPOL, int UNROLL, int THREADS> static void bench128(int mult) {
  const size_t nvec = N / 4, tile = (size_t)THREADS * UNROLL;
  size_t ntiles = nvec / tile;
  int grid = mult > 0 ? SMS * mult : (int)ntiles;
  if ((size_t)grid > ntiles) grid = (int)ntiles;
  char name[128];
  snprintf(name, sizeof(name), "v128 arity%d ld%d st%d unroll%d thr%d grid=%s%d", ARITY, LPOL, SPOL, UNROLL, THREADS,
           mult > 0 ? "SMx" : "tiles:", mult > 0 ? mult : grid);
  report(name, ARITY, run([&] { k128<ARITY, LPOL, SPOL, UNROLL, THREADS><<<grid, THREADS>>>(nvec, (const float4 *)X, (const float4 *)Y, (float4 *)O); }));
}
template <int ARITY, int LPOL, int SPOL, int UNROLL, int THREADS> static void bench256(int mult) {
  const size_t nvec = N / 8, tile = (size_t)THREADS * UNROLL;
  size_t ntiles = nvec / tile;
  int grid = mult > 0 ? SMS * mult : (int)ntiles;
  if ((size_t)grid > ntiles) grid = (int)ntiles;
  char name[128];
  snprintf(name, sizeof(name), "v256 arity%d ld%d st%d unroll%d thr%d grid=%s%d", ARITY, LPOL, SPOL, UNROLL, THREADS,
           mult > 0 ? "SMx" : "tiles:", mult > 0 ? mult : grid);
  report(name, ARITY, run([&] { k256<ARITY, LPOL, SPOL, UNROLL, THREADS><<<grid, THREADS>>>(nvec, (const f8 *)X, (const f8 *)Y, (f8 *)O); }));
}

template <class F, int UNROLL, int THREADS> static void benchfun() {
  const size_t nvec = N / 4, tile = (size_t)THREADS * UNROLL;
  char name[128];
  snprintf(name, sizeof(name), "%s f32 one tile per CTA unroll%d thr%d", F::name(), UNROLL, THREADS);
  report(name, 1, run([&] { kfun<F, UNROLL, THREADS><<<(unsigned)(nvec / tile), THREADS>>>(nvec, (const float4 *)X, (float4 *)O); }));
}
template <class F> static void benchfun_all() {
  benchfun<F, 1, 128>(); benchfun<F, 2, 128>(); benchfun<F, 4, 128>();
  benchfun<F, 1, 256>(); benchfun<F, 2, 256>(); benchfun<F, 4, 256>();
  benchfun<F, 1, 512>(); benchfun<F, 2, 512>(); benchfun<F, 1, 1024>();
}
__global__ void kinit(size_t n, float *x) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = 1e-3f + (float)((i * 2654435761ull) & 0xffffff) * (0.998f / 16777216.0f);   // U[1e-3, 1)
}

int main(int argc, char **argv) {
  N = (size_t)1 << (argc > 1 ? atoi(argv[1]) : 28);
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  SMS = p.multiProcessorCount;
  printf("device %s, %d SMs, N = %zu f32 (%.0f MiB per array)\n", p.name, SMS, N, N * 4.0 / (1 << 20));
  CK(cudaMalloc(&X, N * 4)); CK(cudaMalloc(&Y, N * 4)); CK(cudaMalloc(&O, N * 4));
  kinit<<<148 * 8, 256>>>(N, X);
  kinit<<<148 * 8, 256>>>(N, Y);
  CK(cudaDeviceSynchronize());
  if (argc > 2) {  // "fun": only the transcendental (threads, unroll) probe
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    benchfun_all<FExp>(); benchfun_all<FLog>(); benchfun_all<FSin>(); benchfun_all<FTanh>();
    return 0;
  }
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  report("cudaMemcpyAsync D2D", 1, run([&] { CK(cudaMemcpyAsync(O, X, N * 4, cudaMemcpyDeviceToDevice)); }));
  report("cudaMemsetAsync (write only; GB/s counts 2x)", 1, run([&] { CK(cudaMemsetAsync(O, 1, N * 4)); }));
  // --- unary
  for (int mult : {2, 4, 8, 16, 32, 0}) bench128<1, P_CS, P_CS, 4, 256>(mult);
  for (int mult : {4, 8, 16, 0}) bench128<1, P_DEFAULT, P_DEFAULT, 4, 256>(mult);
  for (int mult : {8, 0}) bench128<1, P_NC, P_DEFAULT, 4, 256>(mult);
  for (int mult : {8, 0}) bench128<1, P_NC, P_CS, 4, 256>(mult);
  for (int mult : {8, 0}) bench128<1, P_CS, P_DEFAULT, 4, 256>(mult);
  for (int mult : {8, 0}) bench128<1, P_DEFAULT, P_CS, 4, 256>(mult);
  for (int mult : {8, 16, 0}) bench128<1, P_CS, P_CS, 2, 256>(mult);
  for (int mult : {4, 8, 0}) bench128<1, P_CS, P_CS, 8, 256>(mult);
  for (int mult : {8, 16, 0}) bench128<1, P_CS, P_CS, 4, 128>(mult);
  for (int mult : {4, 8, 0}) bench128<1, P_CS, P_CS, 4, 512>(mult);
  for (int mult : {2, 4, 0}) bench128<1, P_CS, P_CS, 2, 1024>(mult);
  for (int mult : {4, 8, 16, 0}) bench256<1, P_CS, P_CS, 2, 256>(mult);
  for (int mult : {4, 8, 16, 0}) bench256<1, P_DEFAULT, P_DEFAULT, 2, 256>(mult);
  for (int mult : {4, 8, 0}) bench256<1, P_CS, P_CS, 4, 256>(mult);
  for (int mult : {4, 8, 0}) bench256<1, P_CS, P_CS, 1, 256>(mult);
  for (int mult : {4, 8, 0}) bench256<1, P_CS, P_CS, 2, 512>(mult);
  for (int mult : {8, 16, 0}) bench256<1, P_CS, P_CS, 2, 128>(mult);
  // --- binary
  for (int mult : {4, 8, 16, 0}) bench128<2, P_CS, P_CS, 4, 256>(mult);
  for (int mult : {8, 0}) bench128<2, P_DEFAULT, P_DEFAULT, 4, 256>(mult);
  for (int mult : {4, 8, 0}) bench256<2, P_CS, P_CS, 2, 256>(mult);
  // --- read only
  {
    const size_t nvec = N / 4;
    for (int mult : {4, 8, 16}) {
      char name[96];
      snprintf(name, sizeof(name), "read-only v128 ld.cs unroll4 thr256 grid=SMx%d (GB/s counts 2x)", mult);
      report(name, 1, run([&] { kread<P_CS, 4, 256><<<SMS * mult, 256>>>(nvec, (const float4 *)X, O); }));
      snprintf(name, sizeof(name), "read-only v128 ld.cs unroll8 thr256 grid=SMx%d (GB/s counts 2x)", mult);
      report(name, 1, run([&] { kread<P_CS, 8, 256><<<SMS * mult, 256>>>(nvec, (const float4 *)X, O); }));
    }
  }
  return 0;
}
