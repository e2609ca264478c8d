// Intra-kernel profiler: per-(CTA, warp) event streams of (tag, start|end, %globaltimer) packed in one u64.
//
// Reference: python/triton_dist/tools/profiler/language.py:38-128 (Profiler.create / record writing 64-bit
// (tag | globaltimer_lo) entries into a slot-strided buffer) + context.py / viewer.py (Perfetto export).
// Here: a POD passed by value to kernels; a null buffer compiles the calls down to one predictable branch.
#pragma once
#include "ptx.cuh"

namespace td {

struct ProfBuf {
  unsigned long long* buf;   // [num_slots][cap]: entry 0 of each slot = number of events recorded
  int cap;                   // entries per slot (including the counter)
  int num_slots;
};

// entry layout: [63:56] tag, [55] 1 = start / 0 = end, [54:0] globaltimer (ns)
TD_DEVICE void prof_record(const ProfBuf& pb, int slot, uint32_t tag, bool is_start) {
  if (pb.buf == nullptr || slot >= pb.num_slots) return;
  unsigned long long* s = pb.buf + static_cast<size_t>(slot) * pb.cap;
  const unsigned long long n = s[0];
  if (n + 1 >= static_cast<unsigned long long>(pb.cap)) return;
  s[n + 1] = (static_cast<unsigned long long>(tag & 0xffu) << 56) | (static_cast<unsigned long long>(is_start ? 1 : 0) << 55) |
             (ptx::globaltimer() & ((1ull << 55) - 1));
  s[0] = n + 1;
}

}  // namespace td
