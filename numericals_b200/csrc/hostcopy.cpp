// hostcopy.cpp -- streaming host copies for the staging ring (compiled by g++, not nvcc: AVX2 intrinsics).
//
// Both destinations of a staging copy are write-only for the CPU -- the page-locked ring slot is read next by
// the DMA engine, the caller's OUT buffer by nobody here -- so the copies use non-temporal stores: a regular
// memcpy of a ~1 MiB piece stays under glibc's non-temporal threshold and pays a read-for-ownership of every
// destination line (3 bytes of DRAM traffic per byte copied instead of 2), and the pageable-operand path is
// bound by exactly that traffic (profiles/time_pageable_r02_*.txt, profiles/pciebench_multi_r02_*.txt).
#include <cstddef>
#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace nuk {

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void copy_stream_avx2(char *dst, const char *src, size_t n) {
  // head: bring dst to a 32-byte boundary
  const size_t mis = (size_t)(-(intptr_t)dst) & 31u;
  if (mis) {
    const size_t h = mis < n ? mis : n;
    memcpy(dst, src, h);
    dst += h; src += h; n -= h;
  }
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < n) memcpy(dst + i, src + i, n - i);
}
#endif

void copy_stream(void *dst, const void *src, size_t n) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2 && n >= 4096) {
    copy_stream_avx2(static_cast<char *>(dst), static_cast<const char *>(src), n);
    return;
  }
#endif
  memcpy(dst, src, n);
}

}  // namespace nuk
