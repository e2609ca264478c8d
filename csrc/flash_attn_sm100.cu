// Flash-attention forward for sm_100a: QK^T and PV on tcgen05 tensor cores, accumulators (S double-buffered, O) in
// TMEM, Q/K/V staged by TMA, online softmax by one warpgroup reading S straight out of TMEM (one query row per
// thread, no shuffles), P handed back to the tensor core through swizzled shared memory.
//
// Used by: prefill attention of the TP layers, context-parallel prefill over the gathered KV
// (reference: kernels/nvidia/sp_ag_attention_intra_node.py:257-427 -- a Triton flash kernel with BM=128, BN=64,
// varlen + GQA + causal + zig-zag), Ulysses attention.
//
// CTA = one 128-query tile of one (batch, q-head); 6 warps:
//   warp 0      TMA producer: Q once, then K/V tiles of 128 keys through a 2-stage ring
//   warp 1      TMEM allocation + MMA issuer (one elected thread): S[j&1] = Q K_j^T, O += P_j V_j
//   warps 2-5   softmax: S -> registers (tcgen05.ld), mask, running max with lazy rescale, exp2, P (bf16) -> smem,
//               O rescale in TMEM when the max moved, final O / l -> global (+ LSE)
// QK^T of tile j+1 is issued before PV of tile j, so the tensor core works on S_{j+1} while the softmax warps
// process S_j (the two S buffers ping-pong).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "runtime/driver.h"
#include "td/ptx.cuh"

namespace td {
namespace fa {

constexpr int BMQ = 128;          // queries per CTA
constexpr int HD = 128;           // head dim
constexpr int kThreads = 192;
constexpr int kKVStages = 2;
constexpr int kQBytes = BMQ * HD * 2;         // Q tile: two [128 rows, 64 elements] 16 KB swizzle slabs
constexpr int kSlab = 128 * 128;              // bytes of one [128 rows, 64 elements] slab (Q, P)

struct Params {
  CUtensorMap tmap_q, tmap_k, tmap_v;
  void* o;                 // [B, Sq, Hq, D]
  float* lse;              // [B, Hq, Sq] natural-log LSE (optional)
  const int* q_tile_pos;   // [B, ceil(Sq/128)] position of the first query of each tile in the KV sequence (optional)
  long long o_stride_b, o_stride_s, o_stride_h;
  int B, Sq, Sk, Hq, Hkv;
  int causal, is_bf16;
  float scale_log2;        // sm_scale * log2(e)
  const int* cu_q;         // varlen (v2 kernel): int32 [B + 1] cumulative query lengths; Sq / Sk then hold the packed totals
  const int* cu_k;         // varlen: int32 [B + 1] cumulative key lengths
  const int* seqused_k;    // varlen, optional: int32 [B] keys actually used of each sequence's slot (padded KV caches)
};

// BNK = keys per pipeline step.  128: one CTA per SM (197 KB smem, 384 TMEM columns).  64: TWO CTAs per SM (115 KB, 256
// columns each) -- the softmax of one CTA runs under the MMAs of the other, which is what the kernel is bound by.
template <int BNK>
struct Smem {
  static constexpr int kKVSlab = BNK * 128;            // [BNK keys, 64 elements]
  static constexpr int kKVTile = 2 * kKVSlab;          // [BNK keys, 128 elements]
  static constexpr int kPBytes = BMQ * BNK * 2;        // BNK / 64 slabs of [128 rows, 64 keys]
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kQBytes;
  static constexpr int kV = kK + kKVStages * kKVTile;
  static constexpr int kP = kV + kKVStages * kKVTile;
  static constexpr int kBar = kP + kPBytes;
  static constexpr int kTotal = kBar + 256;            // dynamic smem starts 1024-byte aligned (checked at run time)
  static constexpr int kTmemCols = (2 * BNK + HD) <= 256 ? 256 : 512;
};

enum Bar { Q_FULL = 0, K_FULL = 1, V_FULL = 3, KV_EMPTY = 5, S_FULL = 7, S_FREE = 9, P_READY = 11, PV_DONE = 12, Q_TMEM = 13, NBAR = 14 };

// kTS ("tensor-core operands from TMEM"): Q is copied to TMEM once and P is written over its S buffer as packed 16-bit
// pairs, so both MMAs take their A operand from TMEM and only K / V are read from shared memory.  With M = N = 128 a
// tcgen05.mma whose two operands come from smem needs 8 KB per 64 clk = the full 128 B/clk of the SM's shared memory, and
// the TMA writes and P stores compete for the same port: the smem-operand version is shared-memory bound (224 KB per
// 128x128 tile), the TMEM-operand version moves 128 KB per tile.
template <int BNK, bool kTS = false>
__global__ void __launch_bounds__(kThreads, (BNK == 64 && !kTS) ? 2 : 1) flash_fwd_kernel(const __grid_constant__ Params p) {
  static_assert(!kTS || BNK == 128, "TMEM-operand variant: S0 | S1 | O | Q = 448 columns");
  using Smem = fa::Smem<BNK>;
  constexpr int kKVTile = Smem::kKVTile, kKVSlab = Smem::kKVSlab;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();      // SWIZZLE_128B atoms need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::kBar);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = p.causal ? static_cast<int>(gridDim.x - 1 - blockIdx.x) : static_cast<int>(blockIdx.x);   // causal: longest tiles first
  const int head = blockIdx.y, batch = blockIdx.z;
  const int kv_head = head / (p.Hq / p.Hkv);
  const int n_q_tiles = gridDim.x;
  const int q_row0 = q_tile * BMQ;
  const int q_pos0 = p.q_tile_pos ? p.q_tile_pos[batch * n_q_tiles + q_tile] : q_row0 + (p.Sk - p.Sq);
  const int rows_here = min(BMQ, p.Sq - q_row0);

  int n_tiles = (p.Sk + BNK - 1) / BNK;
  if (p.causal) n_tiles = min(n_tiles, (q_pos0 + rows_here - 1) / BNK + 1);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tmap_q);
    ptx::prefetch_tensormap(&p.tmap_k);
    ptx::prefetch_tensormap(&p.tmap_v);
    ptx::mbar_init(&bars[Q_FULL], 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&bars[K_FULL + s], 1);
      ptx::mbar_init(&bars[V_FULL + s], 1);
      ptx::mbar_init(&bars[KV_EMPTY + s], 1);
      ptx::mbar_init(&bars[S_FULL + s], 1);
      ptx::mbar_init(&bars[S_FREE + s], 128);
    }
    ptx::mbar_init(&bars[P_READY], 128);
    ptx::mbar_init(&bars[PV_DONE], 1);
    ptx::mbar_init(&bars[Q_TMEM], 128);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<1>(tmem_ptr_smem, kTS ? 512 : Smem::kTmemCols);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_o = tmem_base + 2 * BNK;
  const uint32_t tmem_q = tmem_base + 2 * BNK + HD;          // kTS only: Q as 64 columns of packed 16-bit pairs

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars[Q_FULL], kQBytes);
      ptx::tma_load_4d(&p.tmap_q, &bars[Q_FULL], smem + Smem::kQ, 0, q_row0, head, batch);
      ptx::tma_load_4d(&p.tmap_q, &bars[Q_FULL], smem + Smem::kQ + kSlab, 64, q_row0, head, batch);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        if (j >= kKVStages) ptx::mbar_wait(&bars[KV_EMPTY + st], ((j >> 1) - 1) & 1);
        uint8_t* ks = smem + Smem::kK + st * kKVTile;
        uint8_t* vs = smem + Smem::kV + st * kKVTile;
        ptx::mbar_arrive_expect_tx(&bars[K_FULL + st], kKVTile);
        ptx::tma_load_4d(&p.tmap_k, &bars[K_FULL + st], ks, 0, j * BNK, kv_head, batch, ptx::kEvictLast);
        ptx::tma_load_4d(&p.tmap_k, &bars[K_FULL + st], ks + kKVSlab, 64, j * BNK, kv_head, batch, ptx::kEvictLast);
        ptx::mbar_arrive_expect_tx(&bars[V_FULL + st], kKVTile);
        ptx::tma_load_4d(&p.tmap_v, &bars[V_FULL + st], vs, 0, j * BNK, kv_head, batch, ptx::kEvictLast);
        ptx::tma_load_4d(&p.tmap_v, &bars[V_FULL + st], vs + kKVSlab, 64, j * BNK, kv_head, batch, ptx::kEvictLast);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t fmt = p.is_bf16 ? 1u : 0u;
      const uint32_t idesc_qk = ptx::make_idesc(fmt, fmt, BMQ, BNK);
      const uint32_t idesc_pv = ptx::make_idesc(fmt, fmt, BMQ, HD, 0, 1);       // V is the MN-major B operand
      const uint32_t q_addr = ptx::smem_u32(smem + Smem::kQ);
      const uint32_t p_addr = ptx::smem_u32(smem + Smem::kP);
      auto issue_qk = [&](int j) {
        const int st = j & 1, b = j & 1;
        ptx::mbar_wait(&bars[K_FULL + st], (j >> 1) & 1);
        if (j >= 2) ptx::mbar_wait(&bars[S_FREE + b], ((j >> 1) - 1) & 1);
        ptx::tc_fence_after();
        const uint32_t k_addr = ptx::smem_u32(smem + Smem::kK + st * kKVTile);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t qoff = (kk >> 2) * kSlab + (kk & 3) * 32, koff = (kk >> 2) * kKVSlab + (kk & 3) * 32;
          if constexpr (kTS)
            ptx::mma_f16_ts(tmem_base + b * BNK, tmem_q + kk * 8, ptx::make_smem_desc_k128(k_addr + koff), idesc_qk, kk > 0 ? 1u : 0u);
          else
            ptx::mma_f16<1>(tmem_base + b * BNK, ptx::make_smem_desc_k128(q_addr + qoff), ptx::make_smem_desc_k128(k_addr + koff),
                            idesc_qk, kk > 0 ? 1u : 0u);
        }
        ptx::mma_commit(&bars[S_FULL + b]);
      };
      if constexpr (kTS) { ptx::mbar_wait(&bars[Q_TMEM], 0); ptx::tc_fence_after(); }
      else ptx::mbar_wait(&bars[Q_FULL], 0);
      if (n_tiles > 0) issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        if (j + 1 < n_tiles) issue_qk(j + 1);
        ptx::mbar_wait(&bars[V_FULL + st], (j >> 1) & 1);
        ptx::mbar_wait(&bars[P_READY], j & 1);
        ptx::tc_fence_after();
        const uint32_t v_addr = ptx::smem_u32(smem + Smem::kV + st * kKVTile);
#pragma unroll
        for (int kk = 0; kk < BNK / 16; ++kk) {
          const uint64_t a = ptx::make_smem_desc_k128(p_addr + (kk >> 2) * kSlab + (kk & 3) * 32);
          const uint64_t bdesc = ptx::make_smem_desc_mn128(v_addr + kk * 16 * 128, kKVSlab);
          if constexpr (kTS) ptx::mma_f16_ts(tmem_o, tmem_base + (j & 1) * BNK + kk * 8, bdesc, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
          else ptx::mma_f16<1>(tmem_o, a, bdesc, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::mma_commit(&bars[KV_EMPTY + st]);
        ptx::mma_commit(&bars[PV_DONE]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may touch
    const int row = quad * 32 + lane;                // query row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int q_pos = q_pos0 + row;
    const uint32_t p_row = ptx::smem_u32(smem + Smem::kP) + row * 128;
    float m_ref = -INFINITY, l = 0.f;

    if constexpr (kTS) {
      // Q: swizzled smem (TMA) -> registers -> TMEM, one row per thread; column c holds elements (2c, 2c+1)
      ptx::mbar_wait(&bars[Q_FULL], 0);
      const uint32_t q_row = ptx::smem_u32(smem + Smem::kQ) + row * 128;
#pragma unroll
      for (int slab = 0; slab < 2; ++slab) {
        uint32_t qv[32];
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16) {
          const uint4 v = ptx::ld_shared_v4(q_row + slab * kSlab + ((c16 ^ (row & 7)) << 4));
          qv[c16 * 4 + 0] = v.x; qv[c16 * 4 + 1] = v.y; qv[c16 * 4 + 2] = v.z; qv[c16 * 4 + 3] = v.w;
        }
        ptx::tmem_st_32x32b_x32(tmem_q + lane_off + slab * 32, qv);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[Q_TMEM]);
    }

    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      ptx::mbar_wait(&bars[S_FULL + b], (j >> 1) & 1);
      ptx::tc_fence_after();
      constexpr int NC = BNK / 32;
      uint32_t s[NC][32];
#pragma unroll
      for (int c = 0; c < NC; ++c) ptx::tmem_ld_32x32b_x32(tmem_base + lane_off + b * BNK + c * 32, s[c]);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[S_FREE + b]);

      const int key0 = j * BNK;
      const bool need_mask = (p.causal && key0 + BNK - 1 > q_pos0) || (key0 + BNK > p.Sk);
      float mx = -INFINITY;
      if (need_mask) {
        const int limit = p.causal ? min(p.Sk - 1, q_pos) : p.Sk - 1;       // last visible key
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = __uint_as_float(s[c][i]);
            v = (key0 + c * 32 + i <= limit) ? v : -INFINITY;
            s[c][i] = __float_as_uint(v);
            mx = fmaxf(mx, v);
          }
      } else {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s[c][i]));
      }
      mx *= p.scale_log2;
      // lazy rescale: keep the old reference max unless the new one is more than 2^8 larger (p stays <= 256)
      const bool bump = mx > m_ref + 8.f;
      float alpha = 1.f;
      if (bump) {
        alpha = (m_ref == -INFINITY) ? 0.f : ptx::ex2_approx(m_ref - mx);
        m_ref = mx;
        l *= alpha;
      }
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = ptx::ex2_approx(fmaf(__uint_as_float(s[c][i]), p.scale_log2, neg_m));
          sum += e;
          s[c][i] = __float_as_uint(e);
        }
      l += sum;

      if constexpr (kTS) {
        // P (packed 16-bit pairs) overwrites the first BNK/2 columns of this tile's S buffer: same thread, same row.  The
        // previous reader of that buffer (PV of tile j-2) retired before QK of tile j was issued (tcgen05 ops of one
        // thread execute in order), so no wait is needed here.
#pragma unroll
        for (int c = 0; c < NC; c += 2) {
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            pk[i] = p.is_bf16 ? ptx::pack_bf16x2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1]))
                              : ptx::pack_f16x2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1]));
            pk[16 + i] = p.is_bf16 ? ptx::pack_bf16x2(__uint_as_float(s[c + 1][2 * i]), __uint_as_float(s[c + 1][2 * i + 1]))
                                   : ptx::pack_f16x2(__uint_as_float(s[c + 1][2 * i]), __uint_as_float(s[c + 1][2 * i + 1]));
          }
          ptx::tmem_st_32x32b_x32(tmem_base + lane_off + b * BNK + c * 16, pk);
        }
        // O may only be touched after PV of the previous tile has completed -- needed only when some row rescales
        if (j > 0 && __any_sync(0xFFFFFFFFu, bump)) {
          ptx::mbar_wait(&bars[PV_DONE], (j - 1) & 1);
          ptx::tc_fence_after();
        }
      } else {
      // P buffer and O are free once PV of the previous tile has completed
        if (j > 0) {
          ptx::mbar_wait(&bars[PV_DONE], (j - 1) & 1);
          ptx::tc_fence_after();
        }
        // P (bf16 / fp16) -> K-major SWIZZLE_128B: 16-byte chunk c16 of row r lands at chunk (c16 ^ (r & 7))
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 v;
            if (p.is_bf16) {
              v.x = ptx::pack_bf16x2(__uint_as_float(s[c][g * 8 + 0]), __uint_as_float(s[c][g * 8 + 1]));
              v.y = ptx::pack_bf16x2(__uint_as_float(s[c][g * 8 + 2]), __uint_as_float(s[c][g * 8 + 3]));
              v.z = ptx::pack_bf16x2(__uint_as_float(s[c][g * 8 + 4]), __uint_as_float(s[c][g * 8 + 5]));
              v.w = ptx::pack_bf16x2(__uint_as_float(s[c][g * 8 + 6]), __uint_as_float(s[c][g * 8 + 7]));
            } else {
              v.x = ptx::pack_f16x2(__uint_as_float(s[c][g * 8 + 0]), __uint_as_float(s[c][g * 8 + 1]));
              v.y = ptx::pack_f16x2(__uint_as_float(s[c][g * 8 + 2]), __uint_as_float(s[c][g * 8 + 3]));
              v.z = ptx::pack_f16x2(__uint_as_float(s[c][g * 8 + 4]), __uint_as_float(s[c][g * 8 + 5]));
              v.w = ptx::pack_f16x2(__uint_as_float(s[c][g * 8 + 6]), __uint_as_float(s[c][g * 8 + 7]));
            }
            const int col = c * 32 + g * 8;                 // first key of this 16-byte chunk
            const int slab = col >> 6, c16 = (col & 63) >> 3;
            ptx::st_shared_v4(p_row + slab * kSlab + ((c16 ^ (row & 7)) << 4), v);
          }
      }
      // O *= alpha for the rows whose reference max moved (warp-uniform branch: tcgen05.ld/st are warp collectives)
      if (j > 0 && __any_sync(0xFFFFFFFFu, bump)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t o[32];
          ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + c * 32, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          ptx::tmem_st_32x32b_x32(tmem_o + lane_off + c * 32, o);
        }
        ptx::tmem_st_wait();
      }
      if constexpr (kTS) ptx::tmem_st_wait();
      else ptx::fence_proxy_async_smem();   // generic-proxy P stores -> visible to the tensor core (async proxy)
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[P_READY]);
    }

    // epilogue: O / l -> global, LSE
    if (n_tiles > 0) {
      ptx::mbar_wait(&bars[PV_DONE], (n_tiles - 1) & 1);
      ptx::tc_fence_after();
    }
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    const bool live = row < rows_here;
    char* o_row = reinterpret_cast<char*>(p.o) +
                  2 * (static_cast<long long>(batch) * p.o_stride_b + static_cast<long long>(q_row0 + row) * p.o_stride_s +
                       static_cast<long long>(head) * p.o_stride_h);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      if (n_tiles > 0) {
        ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + c * 32, o);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0u;
      }
      if (live) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[g * 8 + i]) * inv_l;
          if (p.is_bf16) {
            v.x = ptx::pack_bf16x2(f[0], f[1]); v.y = ptx::pack_bf16x2(f[2], f[3]);
            v.z = ptx::pack_bf16x2(f[4], f[5]); v.w = ptx::pack_bf16x2(f[6], f[7]);
          } else {
            v.x = ptx::pack_f16x2(f[0], f[1]); v.y = ptx::pack_f16x2(f[2], f[3]);
            v.z = ptx::pack_f16x2(f[4], f[5]); v.w = ptx::pack_f16x2(f[6], f[7]);
          }
          ptx::st_v4(o_row + (c * 32 + g * 8) * 2, v);
        }
      }
    }
    if (p.lse && live) {
      const float lse = (l > 0.f) ? (m_ref + log2f(l)) * 0.6931471805599453f : -INFINITY;
      p.lse[(static_cast<long long>(batch) * p.Hq + head) * p.Sq + q_row0 + row] = lse;
    }
    ptx::tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, kTS ? 512 : Smem::kTmemCols);
  }
}


// =================================================================================================================
// v2: 128-key tiles, Q and P as TMEM A operands, 3-stage K/V ring, TWO softmax warpgroups that split the columns of
// every S tile (two resident warps per scheduler hide the tcgen05.ld / MUFU / reduction latencies one warp per
// scheduler cannot), 4-way ILP in the row reductions, CTAs ordered longest-first for causal masks and adjacent CTAs =
// the q-heads of one GQA group (their K/V tiles hit in L2).
// =================================================================================================================
constexpr int kThreadsV2 = 320;
constexpr int kStagesV2 = 3;
constexpr int kKVTileV2 = 128 * 128 * 2;
struct SmemV2 {
  static constexpr int kQ = 0;                          // reused for the max / sum exchange once Q sits in TMEM
  static constexpr int kK = kQ + kQBytes;
  static constexpr int kV = kK + kStagesV2 * kKVTileV2;
  static constexpr int kBar = kV + kStagesV2 * kKVTileV2;
  static constexpr int kTotal = kBar + 256;
};
enum BarV2 { V2_Q_FULL = 0, V2_K_FULL = 1, V2_V_FULL = 4, V2_KV_EMPTY = 7, V2_S_FULL = 10, V2_P_READY = 12, V2_PV_DONE = 13, V2_Q_TMEM = 14, V2_NBAR = 15 };

// kVarlen: packed variable-length batch -- sequence `batch` owns rows [cu_q[b], cu_q[b+1]) of Q / O and [cu_k[b], cu_k[b+1]) of K / V
// (tensor maps over the packed [T, H, D] tensors, batch coordinate 0); CTAs of query tiles a sequence does not have exit at once.
// Rows of a tile past the end of its sequence belong to the next sequence: they are loaded, never stored (`live`), and keys past
// the end are masked exactly like the tail of a padded batch.
template <bool kVarlen>
__global__ void __launch_bounds__(kThreadsV2, 1) flash_fwd_kernel_v2(const __grid_constant__ Params p, int n_q_tiles) {
  constexpr int BNK = 128;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SmemV2::kBar);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + V2_NBAR);
  float* xmax = reinterpret_cast<float*>(smem + SmemV2::kQ);               // [2 parity][2 halves][128]
  float* xsum = xmax + 2 * 2 * 128;                                        // [2 halves][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_tile = p.Hq * p.B;
  const int tile_rank = static_cast<int>(blockIdx.x) / per_tile, hb = static_cast<int>(blockIdx.x) % per_tile;
  const int q_tile = p.causal ? n_q_tiles - 1 - tile_rank : tile_rank;     // causal: most keys first
  const int head = hb % p.Hq, batch = hb / p.Hq;
  const int kv_head = head / (p.Hq / p.Hkv);
  const int q_row0 = q_tile * BMQ;
  int Sq = p.Sq, Sk = p.Sk, q_base = 0, k_base = 0, tb = batch;          // tb: batch coordinate of the tensor maps / output
  if constexpr (kVarlen) {
    q_base = p.cu_q[batch]; Sq = p.cu_q[batch + 1] - q_base;
    k_base = p.cu_k[batch]; Sk = p.seqused_k ? p.seqused_k[batch] : p.cu_k[batch + 1] - k_base;
    tb = 0;
    if (q_row0 >= Sq || Sk <= 0) return;       // uniform for the CTA, before any barrier / TMEM state exists
  }
  const int q_pos0 = p.q_tile_pos ? p.q_tile_pos[batch * n_q_tiles + q_tile] : q_row0 + (Sk - Sq);
  const int rows_here = min(BMQ, Sq - q_row0);
  int n_tiles = (Sk + BNK - 1) / BNK;
  if (p.causal) n_tiles = min(n_tiles, (q_pos0 + rows_here - 1) / BNK + 1);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tmap_q);
    ptx::prefetch_tensormap(&p.tmap_k);
    ptx::prefetch_tensormap(&p.tmap_v);
    ptx::mbar_init(&bars[V2_Q_FULL], 1);
    for (int s = 0; s < kStagesV2; ++s) {
      ptx::mbar_init(&bars[V2_K_FULL + s], 1);
      ptx::mbar_init(&bars[V2_V_FULL + s], 1);
      ptx::mbar_init(&bars[V2_KV_EMPTY + s], 1);
    }
    ptx::mbar_init(&bars[V2_S_FULL], 1);
    ptx::mbar_init(&bars[V2_S_FULL + 1], 1);
    ptx::mbar_init(&bars[V2_P_READY], 256);
    ptx::mbar_init(&bars[V2_PV_DONE], 1);
    ptx::mbar_init(&bars[V2_Q_TMEM], 256);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<1>(tmem_ptr_smem, 512);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_o = tmem_base + 2 * BNK;
  const uint32_t tmem_q = tmem_base + 2 * BNK + HD;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars[V2_Q_FULL], kQBytes);
      ptx::tma_load_4d(&p.tmap_q, &bars[V2_Q_FULL], smem + SmemV2::kQ, 0, q_base + q_row0, head, tb);
      ptx::tma_load_4d(&p.tmap_q, &bars[V2_Q_FULL], smem + SmemV2::kQ + kSlab, 64, q_base + q_row0, head, tb);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % kStagesV2, use = j / kStagesV2;
        if (use > 0) ptx::mbar_wait(&bars[V2_KV_EMPTY + st], (use - 1) & 1);
        uint8_t* ks = smem + SmemV2::kK + st * kKVTileV2;
        uint8_t* vs = smem + SmemV2::kV + st * kKVTileV2;
        ptx::mbar_arrive_expect_tx(&bars[V2_K_FULL + st], kKVTileV2);
        ptx::tma_load_4d(&p.tmap_k, &bars[V2_K_FULL + st], ks, 0, k_base + j * BNK, kv_head, tb, ptx::kEvictLast);
        ptx::tma_load_4d(&p.tmap_k, &bars[V2_K_FULL + st], ks + kSlab, 64, k_base + j * BNK, kv_head, tb, ptx::kEvictLast);
        ptx::mbar_arrive_expect_tx(&bars[V2_V_FULL + st], kKVTileV2);
        ptx::tma_load_4d(&p.tmap_v, &bars[V2_V_FULL + st], vs, 0, k_base + j * BNK, kv_head, tb, ptx::kEvictLast);
        ptx::tma_load_4d(&p.tmap_v, &bars[V2_V_FULL + st], vs + kSlab, 64, k_base + j * BNK, kv_head, tb, ptx::kEvictLast);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t fmt = p.is_bf16 ? 1u : 0u;
      const uint32_t idesc_qk = ptx::make_idesc(fmt, fmt, BMQ, BNK);
      const uint32_t idesc_pv = ptx::make_idesc(fmt, fmt, BMQ, HD, 0, 1);
      // tcgen05 operations of one thread execute in issue order: QK(j+2) (which overwrites S[j&1], holding P_j) is issued
      // after PV(j), and PV(j) is issued only after every softmax thread has read S_j and written P_j -> no S_FREE barrier
      auto issue_qk = [&](int j) {
        const int st = j % kStagesV2;
        ptx::mbar_wait(&bars[V2_K_FULL + st], (j / kStagesV2) & 1);
        ptx::tc_fence_after();
        const uint32_t k_addr = ptx::smem_u32(smem + SmemV2::kK + st * kKVTileV2);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          ptx::mma_f16_ts(tmem_base + (j & 1) * BNK, tmem_q + kk * 8,
                          ptx::make_smem_desc_k128(k_addr + (kk >> 2) * kSlab + (kk & 3) * 32), idesc_qk, kk > 0 ? 1u : 0u);
        ptx::mma_commit(&bars[V2_S_FULL + (j & 1)]);
      };
      ptx::mbar_wait(&bars[V2_Q_TMEM], 0);
      ptx::tc_fence_after();
      if (n_tiles > 0) issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % kStagesV2;
        if (j + 1 < n_tiles) issue_qk(j + 1);
        ptx::mbar_wait(&bars[V2_V_FULL + st], (j / kStagesV2) & 1);
        ptx::mbar_wait(&bars[V2_P_READY], j & 1);
        ptx::tc_fence_after();
        const uint32_t v_addr = ptx::smem_u32(smem + SmemV2::kV + st * kKVTileV2);
#pragma unroll
        for (int kk = 0; kk < BNK / 16; ++kk)
          ptx::mma_f16_ts(tmem_o, tmem_base + (j & 1) * BNK + kk * 8, ptx::make_smem_desc_mn128(v_addr + kk * 16 * 128, kSlab), idesc_pv,
                          (j > 0 || kk > 0) ? 1u : 0u);
        ptx::mma_commit(&bars[V2_KV_EMPTY + st]);
        ptx::mma_commit(&bars[V2_PV_DONE]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax: 2 warpgroups x 64 columns
    const int half = (warp - 2) >> 2;                // which 64 columns of S / O this thread owns
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int q_pos = q_pos0 + row;
    float m_ref = -INFINITY, l = 0.f;

    // Q slab `half` (64 elements = 32 packed columns): swizzled smem -> registers -> TMEM
    ptx::mbar_wait(&bars[V2_Q_FULL], 0);
    {
      const uint32_t q_row = ptx::smem_u32(smem + SmemV2::kQ) + half * kSlab + row * 128;
      uint32_t qv[32];
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        const uint4 v = ptx::ld_shared_v4(q_row + ((c16 ^ (row & 7)) << 4));
        qv[c16 * 4 + 0] = v.x; qv[c16 * 4 + 1] = v.y; qv[c16 * 4 + 2] = v.z; qv[c16 * 4 + 3] = v.w;
      }
      ptx::tmem_st_32x32b_x32(tmem_q + lane_off + half * 32, qv);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[V2_Q_TMEM]);
    }
    ptx::named_bar_sync(3, 256);                     // every row of Q has been read: its smem becomes the exchange area

    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      ptx::mbar_wait(&bars[V2_S_FULL + b], (j >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t s[2][32];
      ptx::tmem_ld_32x32b_x32(tmem_base + lane_off + b * BNK + half * 64, s[0]);
      ptx::tmem_ld_32x32b_x32(tmem_base + lane_off + b * BNK + half * 64 + 32, s[1]);
      ptx::tmem_ld_wait();

      const int key0 = j * BNK + half * 64;
      const bool need_mask = (p.causal && j * BNK + BNK - 1 > q_pos0) || (j * BNK + BNK > Sk);
      if (need_mask) {
        const int limit = p.causal ? min(Sk - 1, q_pos) : Sk - 1;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (key0 + c * 32 + i > limit) s[c][i] = 0xFF800000u;      // -inf
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(s[c][i]));
      const float mx_local = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      xmax[(b * 2 + half) * 128 + row] = mx_local;
      ptx::tc_fence_before();                          // my S reads are complete before the partner half may overwrite
      ptx::named_bar_sync(3, 256);                     // columns 32..63 of this buffer with its packed P
      ptx::tc_fence_after();
      float mx = fmaxf(mx_local, xmax[(b * 2 + (half ^ 1)) * 128 + row]) * p.scale_log2;

      const bool bump = mx > m_ref + 8.f;
      float alpha = 1.f;
      if (bump) {
        alpha = (m_ref == -INFINITY) ? 0.f : ptx::ex2_approx(m_ref - mx);
        m_ref = mx;
        l *= alpha;
      }
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float e0 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][i]), p.scale_log2, neg_m));
          const float e1 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][i + 1]), p.scale_log2, neg_m));
          sum4[(i >> 1) & 3] += e0 + e1;
          pk[c * 16 + (i >> 1)] = p.is_bf16 ? ptx::pack_bf16x2(e0, e1) : ptx::pack_f16x2(e0, e1);
        }
      l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      // P (64 keys of this half = 32 packed columns) over the S buffer
      ptx::tmem_st_32x32b_x32(tmem_base + lane_off + b * BNK + half * 32, pk);
      if (j > 0 && __any_sync(0xFFFFFFFFu, bump)) {
        ptx::mbar_wait(&bars[V2_PV_DONE], (j - 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + half * 64 + c * 32, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          ptx::tmem_st_32x32b_x32(tmem_o + lane_off + half * 64 + c * 32, o);
        }
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[V2_P_READY]);
    }

    // epilogue: total row sum across the two halves, O / l -> global (each half its 64 columns), LSE
    xsum[half * 128 + row] = l;
    ptx::named_bar_sync(3, 256);
    l += xsum[(half ^ 1) * 128 + row];
    if (n_tiles > 0) {
      ptx::mbar_wait(&bars[V2_PV_DONE], (n_tiles - 1) & 1);
      ptx::tc_fence_after();
    }
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    const bool live = row < rows_here;
    char* o_row = reinterpret_cast<char*>(p.o) +
                  2 * (static_cast<long long>(tb) * p.o_stride_b + static_cast<long long>(q_base + q_row0 + row) * p.o_stride_s +
                       static_cast<long long>(head) * p.o_stride_h + half * 64);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      if (n_tiles > 0) {
        ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + half * 64 + c * 32, o);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0u;
      }
      if (live) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[g * 8 + i]) * inv_l;
          if (p.is_bf16) {
            v.x = ptx::pack_bf16x2(f[0], f[1]); v.y = ptx::pack_bf16x2(f[2], f[3]);
            v.z = ptx::pack_bf16x2(f[4], f[5]); v.w = ptx::pack_bf16x2(f[6], f[7]);
          } else {
            v.x = ptx::pack_f16x2(f[0], f[1]); v.y = ptx::pack_f16x2(f[2], f[3]);
            v.z = ptx::pack_f16x2(f[4], f[5]); v.w = ptx::pack_f16x2(f[6], f[7]);
          }
          ptx::st_v4(o_row + (c * 32 + g * 8) * 2, v);
        }
      }
    }
    if (p.lse && live && half == 0) {
      const float lse = (l > 0.f) ? (m_ref + log2f(l)) * 0.6931471805599453f : -INFINITY;
      p.lse[(static_cast<long long>(tb) * p.Hq + head) * p.Sq + q_base + q_row0 + row] = lse;     // varlen: [Hq, total_q]
    }
    ptx::tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, 512);
  }
}


// =================================================================================================================
// v3 (opt-in, written after the round's GPU budget was spent): TWO 128-query tiles per CTA, one softmax warpgroup per tile.
// The two tiles are independent instruction streams over the same K/V tiles: while warpgroup 0 is in the latency chain of
// its tile (tcgen05.ld -> max -> ex2 -> tcgen05.st -> barrier), the tensor core works on the other tile's MMAs and
// warpgroup 1 keeps the MUFU pipe busy.  TMEM: S0 | S1 | O0 | O1 = 512 columns, P_i overwrites S_i (TMEM A operand of PV_i);
// Q stays in shared memory.  MMA issue order per K/V tile j:  PV0(j) QK0(j+1) PV1(j) QK1(j+1).
// =================================================================================================================
constexpr int kThreadsV3 = 320;
struct SmemV3 {
  static constexpr int kQ = 0;                               // 2 x [128, 128]
  static constexpr int kK = kQ + 2 * kQBytes;
  static constexpr int kV = kK + 2 * kKVTileV2;              // 2 stages
  static constexpr int kBar = kV + 2 * kKVTileV2;
  static constexpr int kTotal = kBar + 256;
};
enum BarV3 { V3_Q_FULL = 0, V3_K_FULL = 1, V3_V_FULL = 3, V3_KV_EMPTY = 5, V3_S_FULL = 7, V3_P_READY = 9, V3_PV_DONE = 11, V3_NBAR = 13 };

__global__ void __launch_bounds__(kThreadsV3, 1) flash_fwd_kernel_v3(const __grid_constant__ Params p, int n_q_tiles) {
  constexpr int BNK = 128;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SmemV3::kBar);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + V3_NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pairs = (n_q_tiles + 1) / 2;                                  // CTAs per (head, batch): two q tiles each
  const int per_tile = p.Hq * p.B;
  const int pair_rank = static_cast<int>(blockIdx.x) / per_tile, hb = static_cast<int>(blockIdx.x) % per_tile;
  const int qp = p.causal ? n_pairs - 1 - pair_rank : pair_rank;            // causal: most keys first
  const int head = hb % p.Hq, batch = hb / p.Hq;
  const int kv_head = head / (p.Hq / p.Hkv);
  int q_row0[2], q_pos0[2], rows_here[2], nt[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int qt = 2 * qp + i;
    q_row0[i] = qt * BMQ;
    rows_here[i] = max(0, min(BMQ, p.Sq - q_row0[i]));
    q_pos0[i] = (p.q_tile_pos && qt < n_q_tiles) ? p.q_tile_pos[batch * n_q_tiles + qt] : q_row0[i] + (p.Sk - p.Sq);
    nt[i] = (p.Sk + BNK - 1) / BNK;
    if (p.causal) nt[i] = min(nt[i], (q_pos0[i] + max(rows_here[i], 1) - 1) / BNK + 1);
    if (rows_here[i] == 0) nt[i] = 0;
  }
  const int n_tiles = max(nt[0], nt[1]);          // both tiles walk the same K/V tiles; a tile past its causal range is fully masked

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tmap_q);
    ptx::prefetch_tensormap(&p.tmap_k);
    ptx::prefetch_tensormap(&p.tmap_v);
    ptx::mbar_init(&bars[V3_Q_FULL], 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&bars[V3_K_FULL + s], 1);
      ptx::mbar_init(&bars[V3_V_FULL + s], 1);
      ptx::mbar_init(&bars[V3_KV_EMPTY + s], 1);
      ptx::mbar_init(&bars[V3_S_FULL + s], 1);
      ptx::mbar_init(&bars[V3_P_READY + s], 128);
      ptx::mbar_init(&bars[V3_PV_DONE + s], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<1>(tmem_ptr_smem, 512);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars[V3_Q_FULL], 2 * kQBytes);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        ptx::tma_load_4d(&p.tmap_q, &bars[V3_Q_FULL], smem + SmemV3::kQ + i * kQBytes, 0, q_row0[i], head, batch);
        ptx::tma_load_4d(&p.tmap_q, &bars[V3_Q_FULL], smem + SmemV3::kQ + i * kQBytes + kSlab, 64, q_row0[i], head, batch);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        if (j >= 2) ptx::mbar_wait(&bars[V3_KV_EMPTY + st], ((j >> 1) - 1) & 1);
        uint8_t* ks = smem + SmemV3::kK + st * kKVTileV2;
        uint8_t* vs = smem + SmemV3::kV + st * kKVTileV2;
        ptx::mbar_arrive_expect_tx(&bars[V3_K_FULL + st], kKVTileV2);
        ptx::tma_load_4d(&p.tmap_k, &bars[V3_K_FULL + st], ks, 0, j * BNK, kv_head, batch, ptx::kEvictLast);
        ptx::tma_load_4d(&p.tmap_k, &bars[V3_K_FULL + st], ks + kSlab, 64, j * BNK, kv_head, batch, ptx::kEvictLast);
        ptx::mbar_arrive_expect_tx(&bars[V3_V_FULL + st], kKVTileV2);
        ptx::tma_load_4d(&p.tmap_v, &bars[V3_V_FULL + st], vs, 0, j * BNK, kv_head, batch, ptx::kEvictLast);
        ptx::tma_load_4d(&p.tmap_v, &bars[V3_V_FULL + st], vs + kSlab, 64, j * BNK, kv_head, batch, ptx::kEvictLast);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t fmt = p.is_bf16 ? 1u : 0u;
      const uint32_t idesc_qk = ptx::make_idesc(fmt, fmt, BMQ, BNK);
      const uint32_t idesc_pv = ptx::make_idesc(fmt, fmt, BMQ, HD, 0, 1);
      // S_i = Q_i K_j^T into TMEM columns [i * 128, +128); in-order issue protects S_i / P_i (see v2)
      auto issue_qk = [&](int i, int j) {
        const int st = j & 1;
        ptx::mbar_wait(&bars[V3_K_FULL + st], (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t q_addr = ptx::smem_u32(smem + SmemV3::kQ + i * kQBytes);
        const uint32_t k_addr = ptx::smem_u32(smem + SmemV3::kK + st * kKVTileV2);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kSlab + (kk & 3) * 32;
          ptx::mma_f16<1>(tmem_base + i * BNK, ptx::make_smem_desc_k128(q_addr + off), ptx::make_smem_desc_k128(k_addr + off), idesc_qk,
                          kk > 0 ? 1u : 0u);
        }
        ptx::mma_commit(&bars[V3_S_FULL + i]);
      };
      auto issue_pv = [&](int i, int j) {
        const int st = j & 1;
        ptx::mbar_wait(&bars[V3_V_FULL + st], (j >> 1) & 1);
        ptx::mbar_wait(&bars[V3_P_READY + i], j & 1);
        ptx::tc_fence_after();
        const uint32_t v_addr = ptx::smem_u32(smem + SmemV3::kV + st * kKVTileV2);
#pragma unroll
        for (int kk = 0; kk < BNK / 16; ++kk)
          ptx::mma_f16_ts(tmem_base + 2 * BNK + i * HD, tmem_base + i * BNK + kk * 8, ptx::make_smem_desc_mn128(v_addr + kk * 16 * 128, kSlab),
                          idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        ptx::mma_commit(&bars[V3_PV_DONE + i]);
      };
      ptx::mbar_wait(&bars[V3_Q_FULL], 0);
      if (n_tiles > 0) { issue_qk(0, 0); issue_qk(1, 0); }
      for (int j = 0; j < n_tiles; ++j) {
        issue_pv(0, j);
        if (j + 1 < n_tiles) issue_qk(0, j + 1);
        issue_pv(1, j);
        ptx::mma_commit(&bars[V3_KV_EMPTY + (j & 1)]);          // K_j / V_j are no longer read once PV1(j) retires
        if (j + 1 < n_tiles) issue_qk(1, j + 1);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax: warpgroup i owns query tile i
    const int i = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + i * BNK, tmem_o = tmem_base + 2 * BNK + i * HD;
    const int q_pos = q_pos0[i] + row;
    const bool tile_live = rows_here[i] > 0;
    float m_ref = -INFINITY, l = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      ptx::mbar_wait(&bars[V3_S_FULL + i], j & 1);
      ptx::tc_fence_after();
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_ld_32x32b_x32(tmem_s + lane_off + c * 32, s[c]);
      ptx::tmem_ld_wait();

      const int key0 = j * BNK;
      const bool need_mask = !tile_live || (p.causal && key0 + BNK - 1 > q_pos0[i]) || (key0 + BNK > p.Sk);
      if (need_mask) {
        const int limit = !tile_live ? -1 : (p.causal ? min(p.Sk - 1, q_pos) : p.Sk - 1);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (key0 + c * 32 + e > limit) s[c][e] = 0xFF800000u;
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int e = 0; e < 32; ++e) mx4[e & 3] = fmaxf(mx4[e & 3], __uint_as_float(s[c][e]));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;
      const bool bump = mx > m_ref + 8.f;
      float alpha = 1.f;
      if (bump) {
        alpha = (m_ref == -INFINITY) ? 0.f : ptx::ex2_approx(m_ref - mx);
        m_ref = mx;
        l *= alpha;
      }
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; c += 2) {
        uint32_t pk[32];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float a0 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][e]), p.scale_log2, neg_m));
          const float a1 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][e + 1]), p.scale_log2, neg_m));
          const float b0 = ptx::ex2_approx(fmaf(__uint_as_float(s[c + 1][e]), p.scale_log2, neg_m));
          const float b1 = ptx::ex2_approx(fmaf(__uint_as_float(s[c + 1][e + 1]), p.scale_log2, neg_m));
          sum4[(e >> 1) & 3] += (a0 + a1) + (b0 + b1);
          pk[e >> 1] = p.is_bf16 ? ptx::pack_bf16x2(a0, a1) : ptx::pack_f16x2(a0, a1);
          pk[16 + (e >> 1)] = p.is_bf16 ? ptx::pack_bf16x2(b0, b1) : ptx::pack_f16x2(b0, b1);
        }
        ptx::tmem_st_32x32b_x32(tmem_s + lane_off + c * 16, pk);        // packed pairs of keys [c*32, c*32+64)
      }
      l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      if (j > 0 && __any_sync(0xFFFFFFFFu, bump)) {
        ptx::mbar_wait(&bars[V3_PV_DONE + i], (j - 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t o[32];
          ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + c * 32, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          ptx::tmem_st_32x32b_x32(tmem_o + lane_off + c * 32, o);
        }
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars[V3_P_READY + i]);
    }

    if (n_tiles > 0) {
      ptx::mbar_wait(&bars[V3_PV_DONE + i], (n_tiles - 1) & 1);
      ptx::tc_fence_after();
    }
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    const bool live = row < rows_here[i];
    char* o_row = reinterpret_cast<char*>(p.o) +
                  2 * (static_cast<long long>(batch) * p.o_stride_b + static_cast<long long>(q_row0[i] + row) * p.o_stride_s +
                       static_cast<long long>(head) * p.o_stride_h);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      if (n_tiles > 0) {
        ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + c * 32, o);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = 0u;
      }
      if (live) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[g * 8 + e]) * inv_l;
          if (p.is_bf16) {
            v.x = ptx::pack_bf16x2(f[0], f[1]); v.y = ptx::pack_bf16x2(f[2], f[3]);
            v.z = ptx::pack_bf16x2(f[4], f[5]); v.w = ptx::pack_bf16x2(f[6], f[7]);
          } else {
            v.x = ptx::pack_f16x2(f[0], f[1]); v.y = ptx::pack_f16x2(f[2], f[3]);
            v.z = ptx::pack_f16x2(f[4], f[5]); v.w = ptx::pack_f16x2(f[6], f[7]);
          }
          ptx::st_v4(o_row + (c * 32 + g * 8) * 2, v);
        }
      }
    }
    if (p.lse && live) {
      const float lse = (l > 0.f) ? (m_ref + log2f(l)) * 0.6931471805599453f : -INFINITY;
      p.lse[(static_cast<long long>(batch) * p.Hq + head) * p.Sq + q_row0[i] + row] = lse;
    }
    ptx::tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace fa
}  // namespace td

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct TdFlashArgs {
  const void* q; const void* k; const void* v; void* o; float* lse; const int* q_tile_pos;
  long long B, Sq, Sk, Hq, Hkv, D;
  long long q_stride_b, q_stride_s, q_stride_h;       // element strides; the head dim is contiguous
  long long k_stride_b, k_stride_s, k_stride_h;
  long long v_stride_b, v_stride_s, v_stride_h;
  long long o_stride_b, o_stride_s, o_stride_h;
  double sm_scale;
  long long causal, is_bf16;
  long long block_n;          // keys per step: 64 (two CTAs per SM, default) or 128; 129 = 128 with Q and P in TMEM
  const int* cu_q; const int* cu_k;   // varlen (block_n 130 only): cumulative lengths, B = number of sequences, Sq / Sk = packed totals
  long long max_sq;                   // varlen: longest query sequence (grid bound)
  const int* seqused_k;               // varlen, optional: keys used per sequence (the slot [cu_k[b], cu_k[b+1]) may be longer)
};

template <int BNK, bool kTS = false>
static int fa_launch(const td::fa::Params& p, dim3 grid, cudaStream_t s) {
  using namespace td::fa;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(flash_fwd_kernel<BNK, kTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BNK>::kTotal);
    if (e != cudaSuccess) { td::drv::set_error("flash_attn: smem attribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set = true;
  }
  flash_fwd_kernel<BNK, kTS><<<grid, kThreads, Smem<BNK>::kTotal, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { td::drv::set_error("flash_attn launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

static int fa_tmap(CUtensorMap* out, const void* base, long long S, long long H, long long B, long long sb, long long ss,
                   long long sh, int is_bf16, int box_rows) {
  auto enc = td::drv::cuTensorMapEncodeTiled_fn();
  if (!enc) { td::drv::set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return -1; }
  const cuuint64_t dims[4] = {128, (cuuint64_t)S, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)ss * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { td::drv::set_error("flash_attn: cuTensorMapEncodeTiled failed: %s", td::drv::err_str(r)); return -1; }
  return 0;
}

extern "C" __attribute__((visibility("default"))) int td_flash_attn_fwd(const TdFlashArgs* a, void* stream_) {
  using namespace td::fa;
  if (a->D != HD) { td::drv::set_error("flash_attn: head_dim must be 128"); return -1; }
  if (a->Hq % a->Hkv != 0) { td::drv::set_error("flash_attn: Hq must be a multiple of Hkv"); return -1; }
  const long long st[] = {a->q_stride_b, a->q_stride_s, a->q_stride_h, a->k_stride_b, a->k_stride_s, a->k_stride_h,
                          a->v_stride_b, a->v_stride_s, a->v_stride_h, a->o_stride_s, a->o_stride_h, a->o_stride_b};
  for (long long s : st)
    if (s % 8 != 0) { td::drv::set_error("flash_attn: strides must be multiples of 8 elements (16 bytes)"); return -1; }
  Params p{};
  const int bnk = (a->block_n >= 128) ? 128 : 64;
  const bool varlen = a->cu_q != nullptr;
  if (varlen && (a->block_n != 130 || a->cu_k == nullptr || a->q_tile_pos != nullptr || a->max_sq <= 0)) {
    td::drv::set_error("flash_attn varlen: v2 kernel only, needs cu_seqlens_k and max_seqlen_q, no q_tile_pos"); return -1;
  }
  const long long tmB = varlen ? 1 : a->B;           // packed tensors: one "batch" of Sq / Sk total rows
  if (fa_tmap(&p.tmap_q, a->q, a->Sq, a->Hq, tmB, a->q_stride_b, a->q_stride_s, a->q_stride_h, (int)a->is_bf16, BMQ)) return -1;
  if (fa_tmap(&p.tmap_k, a->k, a->Sk, a->Hkv, tmB, a->k_stride_b, a->k_stride_s, a->k_stride_h, (int)a->is_bf16, bnk)) return -1;
  if (fa_tmap(&p.tmap_v, a->v, a->Sk, a->Hkv, tmB, a->v_stride_b, a->v_stride_s, a->v_stride_h, (int)a->is_bf16, bnk)) return -1;
  p.cu_q = a->cu_q; p.cu_k = a->cu_k; p.seqused_k = a->seqused_k;
  p.o = a->o; p.lse = a->lse; p.q_tile_pos = a->q_tile_pos;
  p.o_stride_b = a->o_stride_b; p.o_stride_s = a->o_stride_s; p.o_stride_h = a->o_stride_h;
  p.B = (int)a->B; p.Sq = (int)a->Sq; p.Sk = (int)a->Sk; p.Hq = (int)a->Hq; p.Hkv = (int)a->Hkv;
  p.causal = (int)a->causal; p.is_bf16 = (int)a->is_bf16;
  p.scale_log2 = static_cast<float>(a->sm_scale * 1.4426950408889634);
  dim3 grid((unsigned)((a->Sq + BMQ - 1) / BMQ), (unsigned)a->Hq, (unsigned)a->B);
  cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream_);
  if (a->block_n == 131) {     // v3 kernel (opt-in): two q tiles per CTA
    static bool v3_attr = false;
    if (!v3_attr) {
      cudaError_t e = cudaFuncSetAttribute(flash_fwd_kernel_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemV3::kTotal);
      if (e != cudaSuccess) { td::drv::set_error("flash_attn v3: smem attribute: %s", cudaGetErrorString(e)); return -1; }
      v3_attr = true;
    }
    const int nq = (int)((a->Sq + BMQ - 1) / BMQ);
    flash_fwd_kernel_v3<<<dim3((unsigned)(((nq + 1) / 2) * a->Hq * a->B)), kThreadsV3, SmemV3::kTotal, st_>>>(p, nq);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { td::drv::set_error("flash_attn v3 launch: %s", cudaGetErrorString(e)); return -1; }
    return 0;
  }
  if (a->block_n == 130) {     // v2 kernel
    static bool v2_attr = false;
    if (!v2_attr) {
      cudaError_t e = cudaFuncSetAttribute(flash_fwd_kernel_v2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemV2::kTotal);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_fwd_kernel_v2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemV2::kTotal);
      if (e != cudaSuccess) { td::drv::set_error("flash_attn v2: smem attribute: %s", cudaGetErrorString(e)); return -1; }
      v2_attr = true;
    }
    const int nq = (int)(((varlen ? a->max_sq : a->Sq) + BMQ - 1) / BMQ);
    if (varlen) flash_fwd_kernel_v2<true><<<dim3((unsigned)(nq * a->Hq * a->B)), kThreadsV2, SmemV2::kTotal, st_>>>(p, nq);
    else flash_fwd_kernel_v2<false><<<dim3((unsigned)(nq * a->Hq * a->B)), kThreadsV2, SmemV2::kTotal, st_>>>(p, nq);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { td::drv::set_error("flash_attn v2 launch: %s", cudaGetErrorString(e)); return -1; }
    return 0;
  }
  if (a->block_n == 129) return fa_launch<128, true>(p, grid, st_);
  return bnk == 128 ? fa_launch<128>(p, grid, st_) : fa_launch<64>(p, grid, st_);
}
