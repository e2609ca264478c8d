// Warp-level tensor-core MMA (mma.sync) for tiles too small for tcgen05 (decode-sized M): m16n8k16, bf16 x bf16 -> fp32.
// Fragment layout (PTX ISA, lane = 4 * g + tig):  A regs {a0a1: (row g, k 2tig..), a2a3: (row g+8, same k), a4a5: (row g, k 2tig+8..),
// a6a7: (row g+8, k 2tig+8..)};  B regs {b0b1: (k 2tig.., col g), b2b3: (k 2tig+8.., col g)};  D: {d0 d1: (row g, cols 2tig, 2tig+1),
// d2 d3: (row g+8, same cols)}.  Used by the megakernel's tensor-core LINEAR task (csrc/megakernel.cu) and by the DSL intrinsic
// ll.mma_m16n8k16_bf16.
#pragma once
#include <cstdint>

namespace td {
namespace mma_sync {
__device__ __forceinline__ void m16n8k16_bf16(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
}  // namespace mma_sync
}  // namespace td
