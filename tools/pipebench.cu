// tools/pipebench.cu -- issue throughput (lane-ops per clock per SM) of the instructions the
// transcendental kernels are built from, measured on the device with clock64: the numbers the
// instruction budgets in DESIGN.md are computed from.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/pipebench tools/pipebench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 8;

template <class Op>
__global__ void k_pipe(typename Op::T *out, typename Op::T seed, long long *cycles) {
  using T = typename Op::T;
  T v[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) v[c] = seed + (T)(threadIdx.x + c);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < kIters; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) v[c] = Op::f(v[c], seed);
  }
  const long long t1 = clock64();
  T s = 0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct OpFFMA { using T = float; static __device__ float f(float a, float b) { return fmaf(a, b, b); } };
struct OpDFMA { using T = double; static __device__ double f(double a, double b) { return fma(a, b, b); } };
struct OpDADD { using T = double; static __device__ double f(double a, double b) { return a + b; } };
struct OpDMUL { using T = double; static __device__ double f(double a, double b) { return a * b; } };
struct OpF2F_WIDEN { using T = float; static __device__ float f(float a, float b) {
  double d; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(a));
  // bring it back with integer ops only (not measured pipe): take the high word
  return __int_as_float(__double2hiint(d)) ; } };
struct OpF2F_NARROW { using T = double; static __device__ double f(double a, double b) {
  float r; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(r) : "d"(a));
  return __hiloint2double(__float_as_int(r), 0x12345); } };
struct OpMUFU_RCP { using T = float; static __device__ float f(float a, float b) {
  float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; } };
struct OpMUFU_RCP64H { using T = double; static __device__ double f(double a, double b) {
  double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); return r; } };
struct OpFRND_D { using T = double; static __device__ double f(double a, double b) { return rint(a) + 0.0; } };
struct OpFRND_F { using T = float; static __device__ float f(float a, float b) { return rintf(a); } };
struct OpI2F_D { using T = double; static __device__ double f(double a, double b) {
  return (double)__double2hiint(a); } };
struct OpF2I_D { using T = double; static __device__ double f(double a, double b) {
  int i; asm volatile("cvt.rni.s32.f64 %0, %1;" : "=r"(i) : "d"(a)); return __hiloint2double(i, i); } };
struct OpDDIV { using T = double; static __device__ double f(double a, double b) { return b / a; } };
struct OpDSQRT { using T = double; static __device__ double f(double a, double b) { return sqrt(a); } };
struct OpFDIV { using T = float; static __device__ float f(float a, float b) { return b / a; } };
struct OpIMAD { using T = float; static __device__ float f(float a, float b) {
  return __int_as_float(__float_as_int(a) * 3 + __float_as_int(b)); } };
struct OpSHF { using T = float; static __device__ float f(float a, float b) {
  return __int_as_float((__float_as_int(a) >> 3) ^ __float_as_int(b)); } };

template <class Op>
void run(const char *name, int ops_per_call, int sms) {
  using T = typename Op::T;
  const int threads = 256, blocks = sms * 8;
  T *out; long long *cyc;
  cudaMalloc(&out, sizeof(T) * threads * blocks);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_pipe<Op><<<blocks, threads>>>(out, (T)1.000001, cyc);
  cudaEventRecord(e0);
  k_pipe<Op><<<blocks, threads>>>(out, (T)1.000001, cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int mhz = 0; cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
  const double total = (double)threads * blocks * kIters * kChains * ops_per_call;
  // lane-ops per clock per SM at the nominal max clock (the achieved clock may be a little lower)
  const double per_clk_sm = total / (ms * 1e-3) / (mhz * 1e3) / sms;
  printf("%-14s %8.3f ms  %7.1f lane-ops/clk/SM (at %d MHz)  err=%s\n", name, ms, per_clk_sm, mhz / 1000,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  run<OpFFMA>("FFMA", 1, sms);
  run<OpDFMA>("DFMA", 1, sms);
  run<OpDADD>("DADD", 1, sms);
  run<OpDMUL>("DMUL", 1, sms);
  run<OpF2F_WIDEN>("F2F.f64.f32", 1, sms);
  run<OpF2F_NARROW>("F2F.f32.f64", 1, sms);
  run<OpMUFU_RCP>("MUFU.RCP", 1, sms);
  run<OpMUFU_RCP64H>("MUFU.RCP64H", 1, sms);
  run<OpFRND_D>("FRND.F64(+DADD)", 1, sms);
  run<OpFRND_F>("FRND.F32", 1, sms);
  run<OpI2F_D>("I2F.F64", 1, sms);
  run<OpF2I_D>("F2I.F64", 1, sms);
  run<OpDDIV>("div.rn.f64", 1, sms);
  run<OpDSQRT>("sqrt.rn.f64", 1, sms);
  run<OpFDIV>("div.rn.f32", 1, sms);
  run<OpIMAD>("IMAD", 1, sms);
  run<OpSHF>("SHF+LOP", 1, sms);
  return 0;
}
