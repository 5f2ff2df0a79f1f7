// tabstage.cuh -- base of the table-driven functors: the lookup tables of nukmath_tab.cuh are copied
// into shared memory once per CTA by the map kernels (map.cuh:stage_tables).
#pragma once
#include "common.cuh"
#include "nukmath_f64.cuh"

namespace nuk {

// words [W0, W1) of IMAGE (uint64 words; both even) land at the same offsets in shared memory
template <const uint64_t *IMAGE, int BYTES, int W0, int W1> struct TabStageOf {
  static constexpr int kSmemBytes = BYTES;
  nukmath::TabRef tab;
  NUK_D void stage(void *smem) {
    // 16-byte copies, compile-time trip count (every map kernel runs kThreads threads per CTA)
    static_assert(W0 % 2 == 0 && W1 % 2 == 0, "table ranges are 16-byte aligned");
    const uint4 *g = reinterpret_cast<const uint4 *>(IMAGE);
    uint4 *s = static_cast<uint4 *>(smem);
#pragma unroll
    for (int w0 = W0 / 2; w0 < W1 / 2; w0 += kThreads) {
      const int w = w0 + (int)threadIdx.x;
      if (w < W1 / 2) s[w] = g[w];
    }
    // opaque to the optimiser: otherwise the shared-window base (S2UR CgaCtaId + ULEA) is rebuilt per element
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("mov.u32 %0, %1;" : "=r"(tab.base) : "r"(b));
  }
};
template <int W0, int W1> using TabStage = TabStageOf<nukmath::kTabImage, nukmath::kTabBytes, W0, W1>;
using TabAll = TabStage<0, 640>;
using TabLog = TabStage<0, 384>;
using TabExp = TabStage<384, 640>;
using TabExp32 = TabStage<640, NM_TAB_WORDS>;   // the 32-entry exp table of sinh / cosh
using TabPowF = TabStageOf<nukmath::kTabImageF, nukmath::kTabFBytes, 0, NM_TABF_WORDS>;

}  // namespace nuk
