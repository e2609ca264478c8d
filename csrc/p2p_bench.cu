// Micro-benchmark kernels for NVLink transfer mechanisms (used by scripts/gpu_p2p_bench.py and documented in
// profiles/): which instruction path and how many SMs does it take to fill an NVLink-5 port?
//   mode 0: generic push  -- ld.global.nc.v4 (local) -> st.global.v4 (peer), all threads, 4-deep unroll
//   mode 1: generic pull  -- ld (peer) -> st (local)
//   mode 2: TMA bulk push -- one thread per CTA: bulk g2s (local) -> bulk s2g (peer); smem reuse gated on
//                            wait_group.read only; a single completion wait at the end
//   mode 3: TMA bulk pull -- one thread per CTA: bulk g2s (peer) -> bulk s2g (local)
// Reference analogue: python/little_kernel/benchmark/memory/* (HBM / L2 / TMA microbenchmarks).
#include "td/primitives.cuh"
#include "runtime/driver.h"

using namespace td;

namespace {
constexpr int kPiece = 16 * 1024;
constexpr int kSlots = 8;

__global__ void __launch_bounds__(1024, 1) p2p_kernel(int mode, char* dst, const char* src, long long bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const long long per = ((bytes / gridDim.x) + 1023) / 1024 * 1024;
  const long long b0 = min(bytes, per * blockIdx.x), b1 = min(bytes, b0 + per);
  if (mode <= 1) {
    copy16_strided(dst + b0, src + b0, static_cast<size_t>(b1 - b0), threadIdx.x, blockDim.x);
    return;
  }
  if (threadIdx.x != 0) return;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSlots * kPiece);
  for (int i = 0; i < kSlots; ++i) ptx::mbar_init(full + i, 1);
  ptx::fence_barrier_init();
  ptx::fence_proxy_async();
  uint32_t parity = 0;
  long long loaded = b0, stored = b0;
  uint32_t nl = 0, ns = 0;
  while (stored < b1) {
    while (loaded < b1 && nl - ns < kSlots) {
      const uint32_t slot = nl % kSlots;
      if (nl >= kSlots) ptx::bulk_wait_read<kSlots - 1>();   // the store that last used this slot has read it
      const uint32_t n = static_cast<uint32_t>(min(static_cast<long long>(kPiece), b1 - loaded));
      ptx::mbar_arrive_expect_tx(full + slot, n);
      ptx::bulk_g2s(smem + slot * kPiece, src + loaded, n, full + slot);
      loaded += n; ++nl;
      if (nl - ns >= kSlots / 2) break;
    }
    const uint32_t slot = ns % kSlots;
    const uint32_t n = static_cast<uint32_t>(min(static_cast<long long>(kPiece), b1 - stored));
    ptx::mbar_wait(full + slot, (parity >> slot) & 1u);
    parity ^= 1u << slot;
    ptx::bulk_s2g(dst + stored, smem + slot * kPiece, n);
    ptx::bulk_commit();
    stored += n; ++ns;
  }
  ptx::bulk_wait<0>();
}
}  // namespace

TD_API int td_p2p_bench(int mode, int grid, int threads, void* dst, const void* src, long long bytes, void* stream) {
  static bool attr = false;
  const int smem = kSlots * kPiece + 256;
  if (!attr) { TD_CUDA_CHECK(cudaFuncSetAttribute(p2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
  p2p_kernel<<<grid, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(mode, (char*)dst, (const char*)src, bytes);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
