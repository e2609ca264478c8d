// Memory-bound model glue kernels (bf16/fp16 I/O, fp32 math, 16-byte vector accesses):
//   rmsnorm, fused residual-add + rmsnorm, silu(gate)*up, and the fused
//   q/k-RMSNorm + RoPE + KV-cache append used by the TP attention layer.
//
// Reference: the TP demo calls flashinfer.norm.rmsnorm / apply_rope_with_cos_sin_cache_inplace and HF SiLU
// (/root/reference/python/triton_dist/layers/nvidia/tp_attn.py:61-68,165-176, tp_mlp.py:159); the megakernel has
// Triton versions (mega_triton_kernel/kernels/{norm,activation,rope}.py); swiglu.py:374 has fwd/bwd.
// Here they are small CUDA kernels so that the whole decode step is our code and graph-capturable.
#include <algorithm>

#include "td/ptx.cuh"
#include "runtime/driver.h"

using namespace td;

namespace {

template <bool kBF16>
TD_DEVICE void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (kBF16) { f[2 * i] = ptx::bf16_lo(w[i]); f[2 * i + 1] = ptx::bf16_hi(w[i]); }
    else { const __half2 h = *reinterpret_cast<const __half2*>(&w[i]); f[2 * i] = __low2float(h); f[2 * i + 1] = __high2float(h); }
  }
}
template <bool kBF16>
TD_DEVICE uint4 pack8(const float (&f)[8]) {
  uint4 o;
  if constexpr (kBF16) {
    o.x = ptx::pack_bf16x2(f[0], f[1]); o.y = ptx::pack_bf16x2(f[2], f[3]); o.z = ptx::pack_bf16x2(f[4], f[5]); o.w = ptx::pack_bf16x2(f[6], f[7]);
  } else {
    o.x = ptx::pack_f16x2(f[0], f[1]); o.y = ptx::pack_f16x2(f[2], f[3]); o.z = ptx::pack_f16x2(f[4], f[5]); o.w = ptx::pack_f16x2(f[6], f[7]);
  }
  return o;
}

TD_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
TD_DEVICE float block_sum(float v, float* smem) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) smem[w] = v;
  __syncthreads();
  float t = (l < nw) ? smem[l] : 0.f;
  t = warp_sum(t);
  __syncthreads();
  return t;
}

// out = rmsnorm(x (+ residual)) * w ; if residual != null also writes residual_out = x + residual
template <bool kBF16>
__global__ void __launch_bounds__(256) rmsnorm_kernel(uint4* __restrict__ out, const uint4* __restrict__ x,
                                                      const uint4* __restrict__ w, const uint4* __restrict__ residual,
                                                      uint4* __restrict__ residual_out, int H, float eps) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const int nvec = H / 8;
  const uint4* xr = x + static_cast<size_t>(row) * nvec;
  float ss = 0.f;
  // H <= 16384: each thread keeps up to 8 vectors in registers
  float v[8][8];
  int cnt = 0;
  for (int i = threadIdx.x; i < nvec && cnt < 8; i += blockDim.x, ++cnt) {
    unpack8<kBF16>(xr[i], v[cnt]);
    if (residual) {
      float r[8];
      unpack8<kBF16>(residual[static_cast<size_t>(row) * nvec + i], r);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[cnt][e] += r[e];
      residual_out[static_cast<size_t>(row) * nvec + i] = pack8<kBF16>(v[cnt]);
      // normalise what was actually stored (bf16-rounded), like the eager reference does
      unpack8<kBF16>(pack8<kBF16>(v[cnt]), v[cnt]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) ss += v[cnt][e] * v[cnt][e];
  }
  ss = block_sum(ss, red);
  const float inv = rsqrtf(ss / H + eps);
  cnt = 0;
  for (int i = threadIdx.x; i < nvec && cnt < 8; i += blockDim.x, ++cnt) {
    float g[8], o[8];
    unpack8<kBF16>(w[i], g);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = v[cnt][e] * inv * g[e];
    out[static_cast<size_t>(row) * nvec + i] = pack8<kBF16>(o);
  }
}

// out[m, i] = silu(x[m, i]) * x[m, I + i]
template <bool kBF16>
__global__ void silu_mul_kernel(uint4* __restrict__ out, const uint4* __restrict__ x, long long M, int I) {
  const int nvec = I / 8;
  const long long total = M * nvec;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / nvec;
    const int i = static_cast<int>(t - m * nvec);
    float g[8], u[8], o[8];
    unpack8<kBF16>(x[m * 2 * nvec + i], g);
    unpack8<kBF16>(x[m * 2 * nvec + nvec + i], u);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = g[e] / (1.f + __expf(-g[e])) * u[e];
    out[t] = pack8<kBF16>(o);
  }
}

// dx[m, :I] = dy * u * sig(g) * (1 + g * (1 - sig(g))),  dx[m, I:] = dy * g * sig(g)     (backward of silu(g) * u)
template <bool kBF16>
__global__ void silu_mul_bwd_kernel(uint4* __restrict__ dx, const uint4* __restrict__ dy, const uint4* __restrict__ x, long long M, int I) {
  const int nvec = I / 8;
  const long long total = M * nvec;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / nvec;
    const int i = static_cast<int>(t - m * nvec);
    float g[8], u[8], d[8], dg[8], du[8];
    unpack8<kBF16>(x[m * 2 * nvec + i], g);
    unpack8<kBF16>(x[m * 2 * nvec + nvec + i], u);
    unpack8<kBF16>(dy[t], d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sg = 1.f / (1.f + __expf(-g[e]));
      dg[e] = d[e] * u[e] * sg * (1.f + g[e] * (1.f - sg));
      du[e] = d[e] * g[e] * sg;
    }
    dx[m * 2 * nvec + i] = pack8<kBF16>(dg);
    dx[m * 2 * nvec + nvec + i] = pack8<kBF16>(du);
  }
}

// One warp per (token, head).  qkv: [T, (Hq + 2 Hkv) * 128]; q heads are normalised + rotated and written to
// q_out [T, Hq, 128]; k heads normalised + rotated and appended to k_cache; v heads appended to v_cache.
// cache layout [B, max_len, Hkv, 128]; token t goes to (batch_idx[t], positions[t]).  head_dim fixed at 128.
template <bool kBF16>
__global__ void __launch_bounds__(128) qk_norm_rope_kv_kernel(
    const uint2* __restrict__ qkv, uint2* __restrict__ q_out, uint2* __restrict__ k_cache, uint2* __restrict__ v_cache,
    const uint2* __restrict__ q_norm_w, const uint2* __restrict__ k_norm_w, const int* __restrict__ positions,
    const int* __restrict__ batch_idx, int T, int Hq, int Hkv, long long max_len, float eps, float rope_theta) {
  const int warp_global = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int heads = Hq + 2 * Hkv;
  if (warp_global >= T * heads) return;
  const int t = warp_global / heads, h = warp_global % heads;
  const uint2 raw = qkv[(static_cast<size_t>(t) * heads + h) * 32 + lane];   // 4 x 16-bit
  float f[4];
  if constexpr (kBF16) { f[0] = ptx::bf16_lo(raw.x); f[1] = ptx::bf16_hi(raw.x); f[2] = ptx::bf16_lo(raw.y); f[3] = ptx::bf16_hi(raw.y); }
  else {
    const __half2 a = *reinterpret_cast<const __half2*>(&raw.x), b = *reinterpret_cast<const __half2*>(&raw.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
  }
  const int pos = positions[t];
  const bool is_v = h >= Hq + Hkv;
  if (!is_v) {
    const uint2* nw = (h < Hq) ? q_norm_w : k_norm_w;
    if (nw != nullptr) {
      float ss = f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3];
      ss = warp_sum(ss);
      const float inv = rsqrtf(ss / 128.f + eps);
      const uint2 wr = nw[lane];
      float g[4];
      if constexpr (kBF16) { g[0] = ptx::bf16_lo(wr.x); g[1] = ptx::bf16_hi(wr.x); g[2] = ptx::bf16_lo(wr.y); g[3] = ptx::bf16_hi(wr.y); }
      else {
        const __half2 a = *reinterpret_cast<const __half2*>(&wr.x), b = *reinterpret_cast<const __half2*>(&wr.y);
        g[0] = __low2float(a); g[1] = __high2float(a); g[2] = __low2float(b); g[3] = __high2float(b);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        f[e] = f[e] * inv * g[e];
        // the eager reference rounds to 16 bit after the norm
        if constexpr (kBF16) f[e] = __bfloat162float(__float2bfloat16(f[e])); else f[e] = __half2float(__float2half(f[e]));
      }
    }
    // neox-style rotary: element d pairs with d +/- 64; lane l holds d = 4l..4l+3, partner lane = l ^ 16
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float other = __shfl_xor_sync(0xffffffffu, f[e], 16);
      const int d = (lane & 15) * 4 + e;                                 // frequency index 0..63
      const float inv_freq = powf(rope_theta, -static_cast<float>(2 * d) / 128.f);
      float sn, cs;
      sincosf(static_cast<float>(pos) * inv_freq, &sn, &cs);   // precise: angles reach thousands of radians
      f[e] = (lane < 16) ? (f[e] * cs - other * sn) : (f[e] * cs + other * sn);
    }
  }
  uint2 o;
  if constexpr (kBF16) { o.x = ptx::pack_bf16x2(f[0], f[1]); o.y = ptx::pack_bf16x2(f[2], f[3]); }
  else { o.x = ptx::pack_f16x2(f[0], f[1]); o.y = ptx::pack_f16x2(f[2], f[3]); }
  if (h < Hq) {
    q_out[(static_cast<size_t>(t) * Hq + h) * 32 + lane] = o;
  } else {
    const int kvh = is_v ? h - Hq - Hkv : h - Hq;
    uint2* cache = is_v ? v_cache : k_cache;
    const size_t b = batch_idx ? batch_idx[t] : 0;
    // never write past the cache row: a position beyond max_len would land in the next batch row / past the allocation
    // (Engine.serve and KV_Cache.inc_offset reject such requests on the host; this is the device-side guard)
    if (pos >= 0 && pos < max_len) cache[((b * max_len + pos) * Hkv + kvh) * 32 + lane] = o;
  }
}

}  // namespace

TD_API int td_rmsnorm(void* out, const void* x, const void* w, const void* residual, void* residual_out, long long rows,
                      int H, float eps, int is_bf16, void* stream) {
  if (H % 8 || H > 8 * 8 * 256) { td::drv::set_error("rmsnorm: H must be a multiple of 8 and <= 16384"); return -1; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (rows == 0) return 0;
  if (is_bf16)
    rmsnorm_kernel<true><<<(unsigned)rows, 256, 0, s>>>((uint4*)out, (const uint4*)x, (const uint4*)w, (const uint4*)residual, (uint4*)residual_out, H, eps);
  else
    rmsnorm_kernel<false><<<(unsigned)rows, 256, 0, s>>>((uint4*)out, (const uint4*)x, (const uint4*)w, (const uint4*)residual, (uint4*)residual_out, H, eps);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

TD_API int td_silu_mul_bwd(void* dx, const void* dy, const void* x, long long M, int I, int is_bf16, void* stream) {
  if (I % 8) { td::drv::set_error("silu_mul_bwd: I must be a multiple of 8"); return -1; }
  if (M == 0) return 0;
  const long long total = M * (I / 8);
  const int grid = (int)std::min<long long>(148 * 8, (total + 255) / 256);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (is_bf16) silu_mul_bwd_kernel<true><<<grid, 256, 0, s>>>((uint4*)dx, (const uint4*)dy, (const uint4*)x, M, I);
  else silu_mul_bwd_kernel<false><<<grid, 256, 0, s>>>((uint4*)dx, (const uint4*)dy, (const uint4*)x, M, I);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

TD_API int td_silu_mul(void* out, const void* x, long long M, int I, int is_bf16, void* stream) {
  if (I % 8) { td::drv::set_error("silu_mul: I must be a multiple of 8"); return -1; }
  if (M == 0) return 0;
  const long long total = M * (I / 8);
  const int grid = (int)min((long long)148 * 8, (total + 255) / 256);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (is_bf16) silu_mul_kernel<true><<<grid, 256, 0, s>>>((uint4*)out, (const uint4*)x, M, I);
  else silu_mul_kernel<false><<<grid, 256, 0, s>>>((uint4*)out, (const uint4*)x, M, I);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

TD_API int td_qk_norm_rope_kv(const void* qkv, void* q_out, void* k_cache, void* v_cache, const void* q_norm_w,
                              const void* k_norm_w, const void* positions, const void* batch_idx, int T, int Hq, int Hkv,
                              long long max_len, float eps, float rope_theta, int is_bf16, void* stream) {
  if (T == 0) return 0;
  const int warps = T * (Hq + 2 * Hkv);
  const int grid = (warps + 3) / 4;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (is_bf16)
    qk_norm_rope_kv_kernel<true><<<grid, 128, 0, s>>>((const uint2*)qkv, (uint2*)q_out, (uint2*)k_cache, (uint2*)v_cache,
                                                       (const uint2*)q_norm_w, (const uint2*)k_norm_w, (const int*)positions,
                                                       (const int*)batch_idx, T, Hq, Hkv, max_len, eps, rope_theta);
  else
    qk_norm_rope_kv_kernel<false><<<grid, 128, 0, s>>>((const uint2*)qkv, (uint2*)q_out, (uint2*)k_cache, (uint2*)v_cache,
                                                        (const uint2*)q_norm_w, (const uint2*)k_norm_w, (const int*)positions,
                                                        (const int*)batch_idx, T, Hq, Hkv, max_len, eps, rope_theta);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Skinny GEMM (decode): out[b, n] = sum_k x[b, k] * W[n, k]  for b < B <= 8.
// At M <= 8 a 128-row tcgen05 tile wastes 94% of the tensor core and the op is purely weight-streaming, so this is
// a CUDA-core GEMV: one warp per 4 rows of W (4 independent 16-byte loads per lane in flight), x staged in shared
// memory, fp32 accumulation, warp-shuffle reduction.  Same math as the megakernel's LINEAR task (csrc/megakernel.cu).
// ---------------------------------------------------------------------------------------------------------
namespace {
template <bool kBF16>
__global__ void __launch_bounds__(256) gemv_kernel(const uint4* __restrict__ x, const uint4* __restrict__ W, void* __restrict__ out,
                                                   int B, int N, int K, int ldx /*elements*/, int ldo) {
  extern __shared__ __align__(16) uint8_t gsm[];
  uint4* xs = reinterpret_cast<uint4*>(gsm);
  const int kvec = K / 8;
  for (int i = threadIdx.x; i < B * kvec; i += blockDim.x) xs[i] = x[(i / kvec) * (ldx / 8) + (i % kvec)];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int R = 4;
  const int rows_per_cta = (blockDim.x / 32) * R;
  for (int n0 = blockIdx.x * rows_per_cta + warp * R; n0 < N; n0 += gridDim.x * rows_per_cta) {
    float acc[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[r][b] = 0.f;
    constexpr int U = 2;      // k-steps in flight per lane: R * U independent 16-byte loads
    for (int kv0 = lane; kv0 < kvec; kv0 += 32 * U) {
      uint4 wv[U][R];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int kv = kv0 + u * 32;
          wv[u][r] = (n0 + r < N && kv < kvec) ? ptx::ld_nc_v4(W + static_cast<size_t>(n0 + r) * kvec + kv) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int kv = kv0 + u * 32;
        if (kv < kvec) {
          float wf[R][8];
#pragma unroll
          for (int r = 0; r < R; ++r) unpack8<kBF16>(wv[u][r], wf[r]);
#pragma unroll
          for (int b = 0; b < 8; ++b)
            if (b < B) {
              float xf[8];
              unpack8<kBF16>(xs[b * kvec + kv], xf);
#pragma unroll
              for (int r = 0; r < R; ++r)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[r][b] += wf[r][e] * xf[e];
            }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < 8; ++b)
        if (b < B) {
          const float v = warp_sum(acc[r][b]);
          if (lane == 0 && n0 + r < N) {
            if constexpr (kBF16) reinterpret_cast<__nv_bfloat16*>(out)[static_cast<size_t>(b) * ldo + n0 + r] = __float2bfloat16(v);
            else reinterpret_cast<__half*>(out)[static_cast<size_t>(b) * ldo + n0 + r] = __float2half(v);
          }
        }
  }
}
}  // namespace

TD_API int td_gemv(const void* x, const void* W, void* out, int B, int N, int K, int ldx, int ldo, int is_bf16, void* stream) {
  if (B < 1 || B > 8 || K % 8 || ldx % 8) { td::drv::set_error("gemv: 1 <= B <= 8, K % 8 == 0"); return -1; }
  const size_t smem = static_cast<size_t>(B) * K * 2;
  if (smem > 200 * 1024) { td::drv::set_error("gemv: B * K too large for shared memory"); return -1; }
  static size_t smem_set = 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (smem > 48 * 1024 && smem > smem_set) {
    TD_CUDA_CHECK(cudaFuncSetAttribute(gemv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    TD_CUDA_CHECK(cudaFuncSetAttribute(gemv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  const int rows_per_cta = 8 * 4;
  int grid = (N + rows_per_cta - 1) / rows_per_cta;
  if (grid > 148 * 4) grid = 148 * 4;
  if (is_bf16) gemv_kernel<true><<<grid, 256, smem, s>>>((const uint4*)x, (const uint4*)W, out, B, N, K, ldx, ldo);
  else gemv_kernel<false><<<grid, 256, smem, s>>>((const uint4*)x, (const uint4*)W, out, B, N, K, ldx, ldo);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
