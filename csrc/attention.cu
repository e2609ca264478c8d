// GQA flash-decode (split-KV) + LSE combine, single GPU and KV-sharded across ranks.
//
// Reference: kernels/nvidia/flash_decode.py (kernel_gqa_fwd_batch_decode_split_kv :130, ..._combine_kv :308,
// kernel_inter_rank_gqa_fwd_batch_decode_combine_kv :482) and layers/nvidia/sp_flash_decode_layer.py:79-185.
// Decode attention is bandwidth bound (each K/V byte is used once per q-head group), so the kernel is a
// straight streaming design: one CTA per (batch, kv head, split), 4 warps striding over the split's positions,
// lanes splitting head_dim=128 (8 B per lane per row -> fully coalesced 256 B rows), online softmax in fp32,
// then an in-CTA merge of the 4 warps and a tiny combine kernel that merges splits (and ranks).
#include <cstdlib>

#include "td/ptx.cuh"
#include "runtime/driver.h"

using namespace td;

namespace {

constexpr int kD = 128;          // head_dim
constexpr int kMaxG = 8;         // q heads per kv head handled by one CTA
constexpr int kWarps = 4;

TD_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <bool kBF16>
TD_DEVICE void unpack4(const uint2& r, float (&f)[4]) {
  if constexpr (kBF16) { f[0] = ptx::bf16_lo(r.x); f[1] = ptx::bf16_hi(r.x); f[2] = ptx::bf16_lo(r.y); f[3] = ptx::bf16_hi(r.y); }
  else {
    const __half2 a = *reinterpret_cast<const __half2*>(&r.x), b = *reinterpret_cast<const __half2*>(&r.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
  }
}

struct DecodeParams {
  const uint2* q;          // [B, Hq, 128]
  const uint2* k_cache;    // [B, max_len, Hkv, 128] or paged [num_pages, page, Hkv, 128]
  const uint2* v_cache;
  const int* kv_lens;      // [B] number of valid positions in THIS rank's cache
  const int* block_table;  // [B, max_pages] or null
  float* o_part;           // [B, Hq, S, 128] fp32 (unnormalised / normalised: normalised by l)
  float* lse_part;         // [B, Hq, S]
  int B, Hq, Hkv, S;
  long long max_len;       // row stride (contiguous) or pages per batch (paged)
  int page_size, max_pages;
  float sm_scale, soft_cap;
};

template <bool kBF16, int G>
__global__ void __launch_bounds__(kWarps * 32) decode_splitkv_kernel(const DecodeParams p) {
  const int b = blockIdx.x, kvh = blockIdx.y, sp = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = p.kv_lens[b];
  const int per = (len + p.S - 1) / p.S;
  const int j0 = sp * per, j1 = min(len, j0 + per);

  float q[G][4];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    unpack4<kBF16>(p.q[(static_cast<size_t>(b) * p.Hq + kvh * G + g) * 32 + lane], q[g]);
#pragma unroll
    for (int e = 0; e < 4; ++e) q[g][e] *= p.sm_scale;
  }
  float m[G], l[G], o[G][4];
#pragma unroll
  for (int g = 0; g < G; ++g) { m[g] = -INFINITY; l[g] = 0.f; o[g][0] = o[g][1] = o[g][2] = o[g][3] = 0.f; }

  auto row_ptr = [&](const uint2* cache, int j) -> const uint2* {
    size_t row;
    if (p.block_table) {
      const int page = p.block_table[static_cast<size_t>(b) * p.max_pages + j / p.page_size];
      row = static_cast<size_t>(page) * p.page_size + (j % p.page_size);
    } else {
      row = static_cast<size_t>(b) * p.max_len + j;
    }
    return cache + (row * p.Hkv + kvh) * 32 + lane;
  };

  for (int j = j0 + warp; j < j1; j += kWarps * 2) {
    // two positions per iteration: both K rows and both V rows are in flight together
    const int ja = j, jb = j + kWarps;
    const bool hb = jb < j1;
    const uint2 ka = *row_ptr(p.k_cache, ja);
    const uint2 va = *row_ptr(p.v_cache, ja);
    uint2 kb = make_uint2(0, 0), vb = make_uint2(0, 0);
    if (hb) { kb = *row_ptr(p.k_cache, jb); vb = *row_ptr(p.v_cache, jb); }
    float kf[4], vf[4];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      if (it == 1 && !hb) break;
      unpack4<kBF16>(it == 0 ? ka : kb, kf);
      unpack4<kBF16>(it == 0 ? va : vb, vf);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float s = q[g][0] * kf[0] + q[g][1] * kf[1] + q[g][2] * kf[2] + q[g][3] * kf[3];
        s = warp_sum(s);
        if (p.soft_cap > 0.f) s = p.soft_cap * tanhf(s / p.soft_cap);
        const float mn = fmaxf(m[g], s);
        const float corr = __expf(m[g] - mn), pj = __expf(s - mn);
        l[g] = l[g] * corr + pj;
#pragma unroll
        for (int e = 0; e < 4; ++e) o[g][e] = o[g][e] * corr + pj * vf[e];
        m[g] = mn;
      }
    }
  }
  // merge the 4 warps
  __shared__ float sm_m[kWarps][G], sm_l[kWarps][G], sm_o[kWarps][G][kD];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (lane == 0) { sm_m[warp][g] = m[g]; sm_l[warp][g] = l[g]; }
#pragma unroll
    for (int e = 0; e < 4; ++e) sm_o[warp][g][lane * 4 + e] = o[g][e];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < G * kD; idx += kWarps * 32) {
    const int g = idx / kD, d = idx % kD;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) mm = fmaxf(mm, sm_m[w][g]);
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const float c = (sm_m[w][g] == -INFINITY) ? 0.f : __expf(sm_m[w][g] - mm);
      ll += sm_l[w][g] * c; oo += sm_o[w][g][d] * c;
    }
    const size_t hq = static_cast<size_t>(b) * p.Hq + kvh * G + g;
    p.o_part[(hq * p.S + sp) * kD + d] = ll > 0.f ? oo / ll : 0.f;
    if (d == 0) p.lse_part[hq * p.S + sp] = ll > 0.f ? mm + __logf(ll) : -INFINITY;
  }
}

// ---- v2: bandwidth-oriented streaming kernel -------------------------------------------------------------------------------
// ncu of the kernel above (B=8, 8 K context, 8 kv heads): 0.86 TB/s -- two 8-byte loads per lane per array in flight are ~8 KB
// per SM, Little's law needs ~45 KB.  Here 16 lanes cover one 256-byte row with 16-byte loads, a warp covers 2 rows per load
// instruction and keeps kU = 4 K loads + 4 V loads in flight per lane (4 KB per warp, 32 KB per 8-warp CTA, two CTAs per SM).
// Every 16-lane group runs its own online softmax over the rows it sees (16 independent states per CTA), merged by one shuffle
// round inside the warp and through shared memory across warps.
constexpr int kWarps2 = 8;
constexpr int kU = 4;

template <bool kBF16>
TD_DEVICE void unpack8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if constexpr (kBF16) { f[2 * e] = ptx::bf16_lo(w[e]); f[2 * e + 1] = ptx::bf16_hi(w[e]); }
    else { const __half2 h = *reinterpret_cast<const __half2*>(&w[e]); f[2 * e] = __low2float(h); f[2 * e + 1] = __high2float(h); }
  }
}

template <bool kBF16, int G>
__global__ void __launch_bounds__(kWarps2 * 32, (G <= 4) ? 2 : 1) decode_splitkv_kernel_v2(const DecodeParams p) {
  const int b = blockIdx.x, kvh = blockIdx.y, sp = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 4, l8 = lane & 15;         // row inside the warp's pair, 16-byte piece (8 elements) of the 128-wide row
  const int len = p.kv_lens[b];
  const int per = (len + p.S - 1) / p.S;
  const int j0 = sp * per, j1 = min(len, j0 + per);
  const uint4* kc = reinterpret_cast<const uint4*>(p.k_cache);
  const uint4* vc = reinterpret_cast<const uint4*>(p.v_cache);

  float q[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    unpack8<kBF16>(reinterpret_cast<const uint4*>(p.q)[(static_cast<size_t>(b) * p.Hq + kvh * G + g) * 16 + l8], q[g]);
#pragma unroll
    for (int e = 0; e < 8; ++e) q[g][e] *= p.sm_scale;
  }
  float m[G], l[G], o[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[g][e] = 0.f;
  }
  auto row_off = [&](int j) -> size_t {
    size_t row;
    if (p.block_table) {
      const int page = p.block_table[static_cast<size_t>(b) * p.max_pages + j / p.page_size];
      row = static_cast<size_t>(page) * p.page_size + (j % p.page_size);
    } else {
      row = static_cast<size_t>(b) * p.max_len + j;
    }
    return (row * p.Hkv + kvh) * 16 + l8;
  };
  constexpr int kRowsPerIter = kWarps2 * 2 * kU;      // 64 positions per CTA iteration
  for (int jb = j0; jb < j1; jb += kRowsPerIter) {
    uint4 kr[kU], vr[kU];
    int jj[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      jj[u] = jb + (u * kWarps2 + warp) * 2 + sub;
      if (jj[u] < j1) { const size_t off = row_off(jj[u]); kr[u] = ptx::ld_nc_v4(kc + off); vr[u] = ptx::ld_nc_v4(vc + off); }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const bool ok = jj[u] < j1;
      float kf[8], vf[8];
      if (ok) { unpack8<kBF16>(kr[u], kf); unpack8<kBF16>(vr[u], vf); }
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) { kf[e] = 0.f; vf[e] = 0.f; }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) s += q[g][e] * kf[e];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        if (p.soft_cap > 0.f) s = p.soft_cap * tanhf(s / p.soft_cap);
        if (!ok) s = -INFINITY;
        const float mn = fmaxf(m[g], s);
        if (mn > -INFINITY) {
          const float corr = __expf(m[g] - mn), pj = __expf(s - mn);
          l[g] = l[g] * corr + pj;
#pragma unroll
          for (int e = 0; e < 8; ++e) o[g][e] = o[g][e] * corr + pj * vf[e];
          m[g] = mn;
        }
      }
    }
  }
  // merge: first the 2 row groups of a warp with one shuffle round (lanes l and l + 16 hold the same columns) ...
#pragma unroll
  for (int off = 16; off <= 16; off <<= 1) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float mo = __shfl_xor_sync(0xffffffffu, m[g], off), lo = __shfl_xor_sync(0xffffffffu, l[g], off);
      const float mn = fmaxf(m[g], mo);
      const float ca = (m[g] == -INFINITY) ? 0.f : __expf(m[g] - mn), cb = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
      l[g] = l[g] * ca + lo * cb;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float oo = __shfl_xor_sync(0xffffffffu, o[g][e], off);
        o[g][e] = o[g][e] * ca + oo * cb;
      }
      m[g] = mn;
    }
  }
  // ... then the 8 warps through shared memory
  constexpr int kStates = kWarps2;
  __shared__ float sm_m[kStates][G], sm_l[kStates][G];
  __shared__ float sm_o[kStates][G][kD];
  const int st = warp;
  if (sub == 0) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (l8 == 0) { sm_m[st][g] = m[g]; sm_l[st][g] = l[g]; }
#pragma unroll
      for (int e = 0; e < 8; ++e) sm_o[st][g][l8 * 8 + e] = o[g][e];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < G * kD; idx += kWarps2 * 32) {
    const int g = idx / kD, d = idx % kD;
    float mm = -INFINITY;
#pragma unroll 4
    for (int w = 0; w < kStates; ++w) mm = fmaxf(mm, sm_m[w][g]);
    float ll = 0.f, oo = 0.f;
#pragma unroll 4
    for (int w = 0; w < kStates; ++w) {
      const float c = (sm_m[w][g] == -INFINITY) ? 0.f : __expf(sm_m[w][g] - mm);
      ll += sm_l[w][g] * c; oo += sm_o[w][g][d] * c;
    }
    const size_t hq = static_cast<size_t>(b) * p.Hq + kvh * G + g;
    p.o_part[(hq * p.S + sp) * kD + d] = ll > 0.f ? oo / ll : 0.f;
    if (d == 0) p.lse_part[hq * p.S + sp] = ll > 0.f ? mm + __logf(ll) : -INFINITY;
  }
}

// out[b, h, :] = sum_s w_s * o_part[b, h, s, :],  w_s = exp(lse_s - lse_total); one warp per (b, h)
// n_parts = S (intra-rank) or W * S / W ... any flat list of partials for that head.
template <bool kBF16>
__global__ void decode_combine_kernel(uint2* out, float* lse_out, const float* o_part, const float* lse_part, int BH, int n_parts) {
  const int bh = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (bh >= BH) return;
  float mm = -INFINITY;
  for (int s = 0; s < n_parts; ++s) mm = fmaxf(mm, lse_part[static_cast<size_t>(bh) * n_parts + s]);
  float tot = 0.f, acc[4] = {0, 0, 0, 0};
  for (int s = 0; s < n_parts; ++s) {
    const float ls = lse_part[static_cast<size_t>(bh) * n_parts + s];
    if (ls == -INFINITY) continue;
    const float w = __expf(ls - mm);
    tot += w;
    const float4 v = *reinterpret_cast<const float4*>(o_part + (static_cast<size_t>(bh) * n_parts + s) * kD + lane * 4);
    acc[0] += w * v.x; acc[1] += w * v.y; acc[2] += w * v.z; acc[3] += w * v.w;
  }
  const float inv = tot > 0.f ? 1.f / tot : 0.f;
  if (out) {
    uint2 o;
    if constexpr (kBF16) { o.x = ptx::pack_bf16x2(acc[0] * inv, acc[1] * inv); o.y = ptx::pack_bf16x2(acc[2] * inv, acc[3] * inv); }
    else { o.x = ptx::pack_f16x2(acc[0] * inv, acc[1] * inv); o.y = ptx::pack_f16x2(acc[2] * inv, acc[3] * inv); }
    out[static_cast<size_t>(bh) * 32 + lane] = o;
  }
  if (lse_out && lane == 0) lse_out[bh] = tot > 0.f ? mm + __logf(tot) : -INFINITY;
}
// fp32 variant used between ranks: writes normalised fp32 O (for a further merge) instead of 16-bit
__global__ void decode_combine_f32_kernel(float* o_out, float* lse_out, const float* o_part, const float* lse_part, int BH, int n_parts) {
  const int bh = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (bh >= BH) return;
  float mm = -INFINITY;
  for (int s = 0; s < n_parts; ++s) mm = fmaxf(mm, lse_part[static_cast<size_t>(bh) * n_parts + s]);
  float tot = 0.f, acc[4] = {0, 0, 0, 0};
  for (int s = 0; s < n_parts; ++s) {
    const float ls = lse_part[static_cast<size_t>(bh) * n_parts + s];
    if (ls == -INFINITY) continue;
    const float w = __expf(ls - mm);
    tot += w;
    const float4 v = *reinterpret_cast<const float4*>(o_part + (static_cast<size_t>(bh) * n_parts + s) * kD + lane * 4);
    acc[0] += w * v.x; acc[1] += w * v.y; acc[2] += w * v.z; acc[3] += w * v.w;
  }
  const float inv = tot > 0.f ? 1.f / tot : 0.f;
  *reinterpret_cast<float4*>(o_out + static_cast<size_t>(bh) * kD + lane * 4) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
  if (lane == 0) lse_out[bh] = tot > 0.f ? mm + __logf(tot) : -INFINITY;
}

template <bool kBF16>
int launch_split(const DecodeParams& p, int G, cudaStream_t s) {
  dim3 grid(p.B, p.Hkv, p.S);
  static const bool use_v1 = [] { const char* e = getenv("TD_DECODE_V1"); return e && e[0] == '1'; }();
  if (!use_v1) {
    switch (G) {
      case 1: decode_splitkv_kernel_v2<kBF16, 1><<<grid, kWarps2 * 32, 0, s>>>(p); break;
      case 2: decode_splitkv_kernel_v2<kBF16, 2><<<grid, kWarps2 * 32, 0, s>>>(p); break;
      case 4: decode_splitkv_kernel_v2<kBF16, 4><<<grid, kWarps2 * 32, 0, s>>>(p); break;
      case 8: decode_splitkv_kernel_v2<kBF16, 8><<<grid, kWarps2 * 32, 0, s>>>(p); break;
      default: td::drv::set_error("flash_decode: q heads per kv head must be 1, 2, 4 or 8"); return -1;
    }
    TD_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  switch (G) {
    case 1: decode_splitkv_kernel<kBF16, 1><<<grid, kWarps * 32, 0, s>>>(p); break;
    case 2: decode_splitkv_kernel<kBF16, 2><<<grid, kWarps * 32, 0, s>>>(p); break;
    case 4: decode_splitkv_kernel<kBF16, 4><<<grid, kWarps * 32, 0, s>>>(p); break;
    case 8: decode_splitkv_kernel<kBF16, 8><<<grid, kWarps * 32, 0, s>>>(p); break;
    default: td::drv::set_error("flash_decode: q heads per kv head must be 1, 2, 4 or 8"); return -1;
  }
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

struct TdDecodeArgs {
  const void* q; const void* k_cache; const void* v_cache; const void* kv_lens; const void* block_table;
  void* o_part; void* lse_part;
  long long B, Hq, Hkv, S, max_len, page_size, max_pages, is_bf16;
  double sm_scale, soft_cap;
};

TD_API int td_flash_decode_split(const TdDecodeArgs* a, void* stream) {
  DecodeParams p;
  p.q = (const uint2*)a->q; p.k_cache = (const uint2*)a->k_cache; p.v_cache = (const uint2*)a->v_cache;
  p.kv_lens = (const int*)a->kv_lens; p.block_table = (const int*)a->block_table;
  p.o_part = (float*)a->o_part; p.lse_part = (float*)a->lse_part;
  p.B = (int)a->B; p.Hq = (int)a->Hq; p.Hkv = (int)a->Hkv; p.S = (int)a->S; p.max_len = a->max_len;
  p.page_size = (int)a->page_size; p.max_pages = (int)a->max_pages;
  p.sm_scale = (float)a->sm_scale; p.soft_cap = (float)a->soft_cap;
  if (p.Hq % p.Hkv) { td::drv::set_error("flash_decode: Hq must be a multiple of Hkv"); return -1; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return a->is_bf16 ? launch_split<true>(p, p.Hq / p.Hkv, s) : launch_split<false>(p, p.Hq / p.Hkv, s);
}

// out may be null (then only lse_out / o_f32 are produced); o_f32 non-null selects the fp32 (inter-rank) variant
TD_API int td_flash_decode_combine(void* out, void* o_f32, void* lse_out, const void* o_part, const void* lse_part, long long BH,
                                   int n_parts, int is_bf16, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (int)((BH + 3) / 4);
  if (o_f32) decode_combine_f32_kernel<<<grid, 128, 0, s>>>((float*)o_f32, (float*)lse_out, (const float*)o_part, (const float*)lse_part, (int)BH, n_parts);
  else if (is_bf16) decode_combine_kernel<true><<<grid, 128, 0, s>>>((uint2*)out, (float*)lse_out, (const float*)o_part, (const float*)lse_part, (int)BH, n_parts);
  else decode_combine_kernel<false><<<grid, 128, 0, s>>>((uint2*)out, (float*)lse_out, (const float*)o_part, (const float*)lse_part, (int)BH, n_parts);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
