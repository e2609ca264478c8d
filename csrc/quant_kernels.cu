// MXFP8 quantisation: bf16/fp16 [M, K] -> e4m3 [M, K] + UE8M0 scale per 32 consecutive K elements, written in the
// TMEM-ready tiled layout consumed by the block-scaled tcgen05 GEMM (csrc/gemm_sm100.cuh, kFP8):
//   chunk (row / 128, k / 128) = 512 bytes:  byte (row % 32) * 16 + ((row % 128) / 32) * 4 + (k % 128) / 32
// so one 512-byte TMA box per (128 rows, K = 128) can be copied to TMEM with tcgen05.cp.32x128b.warpx4 unchanged.
// The reference has no block-scaled GEMM (per-tensor fp8 only: test_gemm_rs.py:130-145); its only block scaling is
// the per-128 activation quantisation of the LL dispatch (low_latency_all_to_all_v2.py:253-259).
#include <cuda_fp8.h>
#include "td/ptx.cuh"
#include "runtime/driver.h"

using namespace td;

namespace {
template <bool kBF16>
__global__ void __launch_bounds__(256) quant_mxfp8_kernel(const uint2* __restrict__ x, uint32_t* __restrict__ q, uint8_t* __restrict__ sf,
                                                          int M, int K, long long ldx /*elements*/) {
  const int kblocks = K / 128;
  const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp_global >= static_cast<long long>(M) * kblocks) return;
  const int row = static_cast<int>(warp_global / kblocks), kb = static_cast<int>(warp_global % kblocks);
  const uint2 raw = x[(static_cast<long long>(row) * ldx + kb * 128) / 4 + lane];
  float f[4];
  if constexpr (kBF16) { f[0] = ptx::bf16_lo(raw.x); f[1] = ptx::bf16_hi(raw.x); f[2] = ptx::bf16_lo(raw.y); f[3] = ptx::bf16_hi(raw.y); }
  else {
    const __half2 a = *reinterpret_cast<const __half2*>(&raw.x), b = *reinterpret_cast<const __half2*>(&raw.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
  }
  float amax = fmaxf(fmaxf(fabsf(f[0]), fabsf(f[1])), fmaxf(fabsf(f[2]), fabsf(f[3])));
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));   // 8 lanes = 32 elements
  // power-of-two scale 2^e with amax / 2^e <= 448 (no saturation):  e = ceil(log2(amax / 448))
  int e = (amax > 0.f) ? static_cast<int>(ceilf(log2f(amax * (1.f / 448.f)))) : -127;
  e = max(-127, min(127, e));
  const float inv = exp2f(static_cast<float>(-e));
  __nv_fp8x4_e4m3 v(make_float4(f[0] * inv, f[1] * inv, f[2] * inv, f[3] * inv));
  q[(static_cast<long long>(row) * K + kb * 128) / 4 + lane] = v.__x;
  if ((lane & 7) == 0) {
    const size_t chunk = static_cast<size_t>(row / 128) * kblocks + kb;
    sf[chunk * 512 + (row % 32) * 16 + ((row % 128) / 32) * 4 + (lane >> 3)] = static_cast<uint8_t>(e + 127);
  }
}
}  // namespace

// q: [M, K] bytes (contiguous); sf: [ceil(M/128) * (K/128) * 512] bytes, zero-initialised by the caller
TD_API int td_quant_mxfp8(const void* x, void* q, void* sf, int M, int K, long long ldx, int is_bf16, void* stream) {
  if (K % 128 || ldx % 4) { td::drv::set_error("quant_mxfp8: K must be a multiple of 128"); return -1; }
  if (M == 0) return 0;
  const long long warps = static_cast<long long>(M) * (K / 128);
  const int grid = static_cast<int>((warps * 32 + 255) / 256);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (is_bf16) quant_mxfp8_kernel<true><<<grid, 256, 0, s>>>((const uint2*)x, (uint32_t*)q, (uint8_t*)sf, M, K, ldx);
  else quant_mxfp8_kernel<false><<<grid, 256, 0, s>>>((const uint2*)x, (uint32_t*)q, (uint8_t*)sf, M, K, ldx);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
