// Gated delta rule, fused recurrent form (decode and short prefill of Qwen3-Next style linear attention).
//   S_t = exp(g_t) * S_{t-1} + beta_t * k_t^T (v_t - k_t (exp(g_t) S_{t-1}));   o_t = scale * q_t S_t
// Reference contract: kernels/nvidia/gdn.py:926 (chunk_gated_delta_rule_fwd); this is the token-sequential companion
// used for decode steps, where the [Dk, Dv] state stays in registers for the whole call.
//
// CTA = (batch, head, 32 value columns); 128 threads: thread (kq = tid / 32, vc = tid % 32) owns S[kq*Dk/4 .. , vc].
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "runtime/driver.h"

namespace {

template <typename T> __device__ __forceinline__ float to_f(T x);
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <> __device__ __forceinline__ float to_f<__half>(__half x) { return __half2float(x); }
template <> __device__ __forceinline__ float to_f<float>(float x) { return x; }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16(x); }
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half(x); }
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }

template <typename T, int DK>
__global__ void __launch_bounds__(128) gdn_recurrent_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                            const float* __restrict__ g, const float* __restrict__ beta,
                                                            float* __restrict__ state /* [B,H,DK,Dv] in/out */, T* __restrict__ o,
                                                            int Tlen, int H, int Dv, float scale) {
  constexpr int KQ = DK / 4;
  const int vt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kq = threadIdx.x >> 5, vc = threadIdx.x & 31;
  const int col = vt * 32 + vc;
  const bool live = col < Dv;
  __shared__ float sk[DK], sq[DK], red[2][4][32];
  float S[KQ];
  float* st = state + ((static_cast<size_t>(b) * H + h) * DK) * Dv;
#pragma unroll
  for (int i = 0; i < KQ; ++i) S[i] = live ? st[static_cast<size_t>(kq * KQ + i) * Dv + col] : 0.f;
  for (int t = 0; t < Tlen; ++t) {
    const size_t tok = (static_cast<size_t>(b) * Tlen + t) * H + h;
    for (int i = threadIdx.x; i < DK; i += 128) { sk[i] = to_f(k[tok * DK + i]); sq[i] = to_f(q[tok * DK + i]) * scale; }
    __syncthreads();
    const float decay = __expf(g[tok]), bt = beta[tok];
    float pred = 0.f;
#pragma unroll
    for (int i = 0; i < KQ; ++i) { S[i] *= decay; pred = fmaf(sk[kq * KQ + i], S[i], pred); }
    red[0][kq][vc] = pred;
    __syncthreads();
    const float delta = ((live ? to_f(v[tok * Dv + col]) : 0.f) - (red[0][0][vc] + red[0][1][vc] + red[0][2][vc] + red[0][3][vc])) * bt;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < KQ; ++i) { S[i] = fmaf(sk[kq * KQ + i], delta, S[i]); acc = fmaf(sq[kq * KQ + i], S[i], acc); }
    red[1][kq][vc] = acc;
    __syncthreads();
    if (kq == 0 && live) o[tok * Dv + col] = from_f<T>(red[1][0][vc] + red[1][1][vc] + red[1][2][vc] + red[1][3][vc]);
  }
#pragma unroll
  for (int i = 0; i < KQ; ++i)
    if (live) st[static_cast<size_t>(kq * KQ + i) * Dv + col] = S[i];
}

template <typename T>
int launch(const void* q, const void* k, const void* v, const float* g, const float* beta, float* state, void* o, int B, int Tlen, int H,
           int Dk, int Dv, float scale, cudaStream_t s) {
  dim3 grid((Dv + 31) / 32, H, B);
  auto Q = reinterpret_cast<const T*>(q); auto K = reinterpret_cast<const T*>(k); auto V = reinterpret_cast<const T*>(v);
  auto O = reinterpret_cast<T*>(o);
  if (Dk == 128) gdn_recurrent_kernel<T, 128><<<grid, 128, 0, s>>>(Q, K, V, g, beta, state, O, Tlen, H, Dv, scale);
  else if (Dk == 64) gdn_recurrent_kernel<T, 64><<<grid, 128, 0, s>>>(Q, K, V, g, beta, state, O, Tlen, H, Dv, scale);
  else if (Dk == 256) gdn_recurrent_kernel<T, 256><<<grid, 128, 0, s>>>(Q, K, V, g, beta, state, O, Tlen, H, Dv, scale);
  else { td::drv::set_error("gdn: Dk must be 64, 128 or 256"); return -1; }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { td::drv::set_error("gdn launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}
}  // namespace

// dtype: 0 fp16, 1 bf16, 2 fp32.  q,k: [B,T,H,Dk]; v,o: [B,T,H,Dv]; g,beta: fp32 [B,T,H]; state fp32 [B,H,Dk,Dv] (updated in place)
extern "C" __attribute__((visibility("default"))) int td_gdn_recurrent(const void* q, const void* k, const void* v, const float* g,
                                                                      const float* beta, float* state, void* o, int B, int T, int H,
                                                                      int Dk, int Dv, float scale, int dtype, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == 1) return launch<__nv_bfloat16>(q, k, v, g, beta, state, o, B, T, H, Dk, Dv, scale, s);
  if (dtype == 0) return launch<__half>(q, k, v, g, beta, state, o, B, T, H, Dk, Dv, scale, s);
  return launch<float>(q, k, v, g, beta, state, o, B, T, H, Dk, Dv, scale, s);
}
