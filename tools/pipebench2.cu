// tools/pipebench2.cu -- do Blackwell's packed single-float instructions (FFMA2 / FADD2 / FMUL2: two IEEE
// round-to-nearest operations per instruction, PTX fma.rn.f32x2) relieve the ISSUE bound of the single-float
// transcendental maps?  Measures lane-operations per clock per SM for scalar and packed forms, alone and
// interleaved with an independent ALU-pipe chain (LOP3 / SHF), 8 independent chains per thread.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/pipebench2 tools/pipebench2.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 8;
typedef unsigned long long u64;

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
  u64 r;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 r;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 pack(float lo, float hi) {
  return (u64)__float_as_uint(lo) | ((u64)__float_as_uint(hi) << 32);
}

// MODE 0 FFMA, 1 FFMA2, 2 FADD2, 3 FMUL2, 4 FFMA + LOP/SHF chain, 5 FFMA2 + LOP/SHF chain, 6 LOP/SHF alone,
// 7 FFMA2 + 2 LOP/SHF chains, 8 FSEL+FFMA (select then fma), 9 FFMA2 with per-lane different constants via pack
template <int MODE>
__global__ void k(float *out, float seed, unsigned iseed) {
  float v[kChains], w[kChains];
  u64 p[kChains];
  unsigned z[kChains], y[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) {
    v[c] = seed + threadIdx.x + c;
    w[c] = seed * 0.5f + c;
    p[c] = pack(v[c], w[c]);
    z[c] = iseed + threadIdx.x * 7u + c;
    y[c] = iseed * 3u + c;
  }
  const u64 ps = pack(seed, seed * 0.999f);
  for (int i = 0; i < kIters; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
      if (MODE == 0) v[c] = fmaf(v[c], seed, seed);
      if (MODE == 1 || MODE == 5 || MODE == 7) p[c] = ffma2(p[c], ps, ps);
      if (MODE == 2) p[c] = fadd2(p[c], ps);
      if (MODE == 3) p[c] = fmul2(p[c], ps);
      if (MODE == 4) v[c] = fmaf(v[c], seed, seed);
      if (MODE == 4 || MODE == 5 || MODE == 6 || MODE == 7) z[c] = (z[c] >> 3) ^ (z[c] + iseed);
      if (MODE == 7) y[c] = (y[c] << 5) ^ (y[c] | iseed);
      if (MODE == 8) v[c] = fmaf((z[c] & 1u) ? v[c] : w[c], seed, seed);
    }
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += v[c] + w[c] + __uint_as_float((unsigned)p[c]) + __uint_as_float((unsigned)(p[c] >> 32)) + (float)z[c] + (float)y[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, double fp_lane_ops, double alu_lane_ops, int sms) {
  const int threads = 256, blocks = sms * 8;
  float *out;
  cudaMalloc(&out, sizeof(float) * threads * blocks);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, 1.000001f, 12345u);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, 1.000001f, 12345u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int mhz = 0;
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
  const double calls = (double)threads * blocks * kIters * kChains;
  const double clk = ms * 1e-3 * mhz * 1e3 * sms;
  printf("%-34s %7.3f ms  fp %6.1f  alu %6.1f lane-ops/clk/SM   (%.2f SMSP cycles per body)  err=%s\n", name, ms,
         calls * fp_lane_ops / clk, calls * alu_lane_ops / clk, (ms * 1e-3 * mhz * 1e3) / (16.0 * kIters * kChains),
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d; fp lane-ops count each scalar FMA/ADD/MUL (a packed instruction = 2); the last column is SMSP issue cycles per loop body\n", sms);
  run<0>("FFMA", 1, 0, sms);
  run<1>("FFMA2", 2, 0, sms);
  run<2>("FADD2", 2, 0, sms);
  run<3>("FMUL2", 2, 0, sms);
  run<6>("SHF+LOP3+IADD (alu chain)", 0, 3, sms);
  run<4>("FFMA  + alu chain", 1, 3, sms);
  run<5>("FFMA2 + alu chain", 2, 3, sms);
  run<7>("FFMA2 + 2 alu chains", 2, 6, sms);
  run<8>("FSEL + FFMA", 1, 1, sms);
  return 0;
}
