// Whole-model persistent decode kernel ("megakernel") for small batches: one launch runs every layer of a
// tensor-parallel dense LLM decode step as a stream of tile-level TASKS pulled from per-CTA static queues and
// ordered by a scoreboard of monotone counters.
//
// Reference: python/triton_dist/mega_triton_kernel/** -- ModelBuilder records ops, ops are tiled into tasks
// (core/task_base.py:162-257), a scheduler fills per-SM queues (core/scheduler.py:103-168), a generated Triton kernel
// loops {fetch task -> scoreboard.wait_deps -> run tile -> release} (core/code_generator.py:101-180).
// Here the interpreter is one hand-written CUDA kernel; the builder/scheduler is native Python+C (triton_dist/mega_kernel).
//
// Design points:
//   * decode with B <= 8 tokens is HBM-bound: linears are weight-streaming GEMVs on CUDA cores (tensor cores cannot
//     help at M <= 8), 16-byte loads, 4 rows of W in flight per warp;
//   * 9 <= B <= 64 tokens: the same LINEAR tasks run on the tensor cores (mma.sync m16n8k16, fp32 accumulate): still
//     weight-streaming -- every lane keeps 8 independent 16-byte weight loads in flight and feeds them to the MMAs straight
//     from registers (the K order inside a 32-wide chunk is permuted identically for both operands, so a lane's 16 bytes
//     ARE its B fragments), the activations are staged once per K chunk in fragment order (conflict-free 16-byte LDS).
//     tcgen05 needs 128-row tiles and a TMEM round trip per task; at <= 64 rows per tile the warp-level MMA is the fit;
//   * scoreboard counters are never reset: task k waits for  sb[dep] >= epoch * dep_count  where `epoch` is a
//     device-resident step counter (CUDA-graph replayable, no memset);
//   * the tensor-parallel all-reduce is a task too: one-shot over NVLink (multimem.ld_reduce when the heap has a
//     multicast mapping, P2P loads otherwise) with per-(op, slice) epoch flags, fused with the residual add.
#include "td/primitives.cuh"
#include "runtime/driver.h"

using namespace td;

namespace {

constexpr int kMKThreads = 256;
constexpr int kMaxB = 8;

enum TaskType : int { T_RMSNORM = 1, T_LINEAR = 2, T_QKROPE = 3, T_ATTN = 4, T_ALLREDUCE = 5, T_COPY = 6, T_ATTN_COMBINE = 7, T_SILU_MUL = 8, T_ADD = 9, T_PREFETCH = 10,
                      T_QKROPE_PAGED = 11, T_ATTN_PAGED = 12, T_FLASH_ATTN = 13, T_QKROPE_SPLIT = 14 };   // paged KV cache: a[9] = page_size | max_pages << 16, block table = ptrs[index of vcache + 1]

struct Task {            // 16 x int32
  int type, dep_idx, dep_count, sig_idx;
  int a[12];
};

struct MKParams {
  const Task* tasks;           // all tasks, grouped by CTA
  const int* queue_off;        // [grid + 1] range of each CTA's queue
  void* const* ptrs;           // pointer table referenced by the tasks
  uint32_t* sb;                // scoreboard counters (monotone)
  uint32_t* epoch;             // [0] completed steps, [1] exit counter, [2] dynamic-scheduler cursor
  int dynamic;                 // 1: one global queue, CTAs fetch the next task with atomicAdd (runtime scheduler)
  int num_tasks;
  SymmCtx symm;
  int B;
};

TD_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
TD_DEVICE void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = ptx::bf16_lo(w[i]); f[2 * i + 1] = ptx::bf16_hi(w[i]); }
}
TD_DEVICE uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = ptx::pack_bf16x2(f[0], f[1]); o.y = ptx::pack_bf16x2(f[2], f[3]); o.z = ptx::pack_bf16x2(f[4], f[5]); o.w = ptx::pack_bf16x2(f[6], f[7]);
  return o;
}

// ---- RMSNORM: out = rmsnorm(in + res) * w ; res_out = in + res.  a: in, res(-1), w, out, res_out(-1), H, eps_bits ----
TD_DEVICE void task_rmsnorm(const MKParams& p, const Task& t, float* red) {
  const uint4* in = (const uint4*)p.ptrs[t.a[0]];
  const uint4* res = t.a[1] >= 0 ? (const uint4*)p.ptrs[t.a[1]] : nullptr;
  const uint4* w = (const uint4*)p.ptrs[t.a[2]];
  uint4* out = (uint4*)p.ptrs[t.a[3]];
  uint4* res_out = t.a[4] >= 0 ? (uint4*)p.ptrs[t.a[4]] : nullptr;
  const int H = t.a[5], nvec = H / 8;
  const float eps = __int_as_float(t.a[6]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = 0; b < p.B; ++b) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < nvec; i += kMKThreads) {
      float f[8];
      unpack8(in[b * nvec + i], f);
      if (res) {
        float r[8];
        unpack8(res[b * nvec + i], r);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += r[e];
        const uint4 pk = pack8(f);
        if (res_out) res_out[b * nvec + i] = pk;
        unpack8(pk, f);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += f[e] * f[e];
    }
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < kMKThreads / 32; ++i) tot += red[i];
    const float inv = rsqrtf(tot / H + eps);
    __syncthreads();
    for (int i = threadIdx.x; i < nvec; i += kMKThreads) {
      float f[8], g[8];
      if (res) { if (res_out) unpack8(res_out[b * nvec + i], f); else { float r[8]; unpack8(in[b * nvec + i], f); unpack8(res[b * nvec + i], r); for (int e = 0; e < 8; ++e) f[e] += r[e]; } }
      else unpack8(in[b * nvec + i], f);
      unpack8(w[i], g);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = f[e] * inv * g[e];
      out[b * nvec + i] = pack8(f);
    }
  }
}

// ---- LINEAR (GEMV tile): out[b, n0 + j] = sum_k act(x)[b, k] * W[n0 + j, k].  a: x, w, out, K, ldo, n0, n_cnt, act, ldx ----
// One CTA per SM means 8 warps have to keep the SM's share of HBM bandwidth busy (~45 KB in flight): every lane keeps
// R * U independent 16-byte weight loads in flight; the batch dimension is a template parameter so the decode case
// (B = 1) can afford U = 4 without spilling.
template <int MB, int U>
TD_DEVICE void linear_rows(const uint4* __restrict__ W, const uint4* __restrict__ xs, __nv_bfloat16* __restrict__ out, int B, int kvec,
                           int ldo, int n0, int n_cnt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int R = 4;
  for (int j0 = warp * R; j0 < n_cnt; j0 += (kMKThreads / 32) * R) {
    float acc[R][MB];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < MB; ++b) acc[r][b] = 0.f;
    for (int kv0 = lane; kv0 < kvec; kv0 += 32 * U) {
      uint4 wv[U][R];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int kv = kv0 + u * 32;
          wv[u][r] = (j0 + r < n_cnt && kv < kvec) ? ptx::ld_nc_v4(W + static_cast<size_t>(n0 + j0 + r) * kvec + kv) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int kv = kv0 + u * 32;
        if (kv < kvec) {
#pragma unroll
          for (int b = 0; b < MB; ++b) {
            if (b < B) {
              float xf[8];
              unpack8(xs[b * kvec + kv], xf);
#pragma unroll
              for (int r = 0; r < R; ++r) {
                float wf[8];
                unpack8(wv[u][r], wf);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[r][b] += wf[e] * xf[e];
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < MB; ++b)
        if (b < B) {
          const float v = warp_sum(acc[r][b]);
          if (lane == 0 && j0 + r < n_cnt) out[static_cast<size_t>(b) * ldo + n0 + j0 + r] = __float2bfloat16(v);
        }
  }
}

// ---- LINEAR on tensor cores (9 <= B <= 64) ------------------------------------------------------------------------------
constexpr int kMmaMaxB = 64;
constexpr int kMmaStageBytes = 128 * 1024;       // activation fragments of one K chunk: Bpad x KC x 2 bytes

TD_DEVICE void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

TD_DEVICE void mma_bf16_16816p(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {   // d: 4 accumulators
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 8 activated elements x[b, kv*8 .. kv*8+7] (act 0: plain, 1: silu(gate) * up, 2: RMSNorm with the row's rs and the norm weight)
TD_DEVICE uint4 linear_act8(const uint4* x, int b, int kv, int act, int ldx8, int kvec, const float* rs, const uint4* nw) {
  if (act == 0) return x[b * ldx8 + kv];
  float f[8], g[8];
  unpack8(x[b * ldx8 + kv], f);
  if (act == 1) {
    unpack8(x[b * ldx8 + kvec + kv], g);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = f[e] / (1.f + __expf(-f[e])) * g[e];
  } else {
    unpack8(nw[kv], g);
    const float r = rs[b];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = f[e] * r * g[e];
  }
  return pack8(f);
}

// mt = 16-row tiles of the (padded) batch (1..4).  The task's n_cnt columns are 8-column GROUPS.  With >= 8 groups warp w owns groups
// {pass * 8 NGW + w + 8 i}; with fewer, 8 / G warps share one group and split the K chunks between them (partial sums meet in shared
// memory) so that every warp streams weights whatever the tile width.  NGW groups x U chunks = 8 independent 16-byte weight loads per
// lane.  Lane (g = lane / 4, tig = lane % 4).
// Fragment layout of the staged activations: frag[((m * C32 + c) * 32 + lane) * 2 + hi] (uint4) = row m*16 + g + 8*hi,
// elements k = c*32 + tig*8 .. +7 -- exactly the 8 k's of the lane's weight vector, so for the two MMAs of a 32-wide chunk
// (s = 0, 1):  A regs = {X[2s], Y[2s], X[2s+1], Y[2s+1]},  B regs = {Wv[2s], Wv[2s+1]}.
template <int NGW>
TD_DEVICE void linear_mma(const MKParams& p, const Task& t, uint8_t* smem) {
  constexpr int MTMAX = kMmaMaxB / 16, U = 8 / NGW, NW = kMKThreads / 32;
  const uint4* x = (const uint4*)p.ptrs[t.a[0]];
  const uint4* W = (const uint4*)p.ptrs[t.a[1]];
  __nv_bfloat16* out = (__nv_bfloat16*)p.ptrs[t.a[2]];
  const int K = t.a[3], ldo = t.a[4], n0 = t.a[5], n_cnt = t.a[6], act = t.a[7], ldx8 = t.a[8] / 8;
  const uint4* nw = act == 2 ? (const uint4*)p.ptrs[t.a[9]] : nullptr;
  const int kvec = K / 8, B = p.B, mt = (B + 15) / 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  uint4* frag = reinterpret_cast<uint4*>(smem);
  float* rs = reinterpret_cast<float*>(smem + kMmaStageBytes);          // [64] per-row rsqrt(mean square) for act 2
  if (act == 2) {
    const float eps = __int_as_float(t.a[10]);
    for (int b = warp; b < B; b += NW) {
      float ss = 0.f;
      for (int kv = lane; kv < kvec; kv += 32) {
        float f[8];
        unpack8(x[b * ldx8 + kv], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += f[e] * f[e];
      }
      ss = warp_sum(ss);
      if (lane == 0) rs[b] = rsqrtf(ss / static_cast<float>(K) + eps);
    }
  }
  const int KC = min(K, (kMmaStageBytes / (mt * 16 * 2)) / 32 * 32);  // elements of K staged at a time (multiple of 32)
  const int groups = n_cnt / 8;
  int G = 1;
  while (G < groups && G < NW) G <<= 1;
  const int kparts = NW / G;                       // warps that share one group and split the K chunks (1 when groups >= 8)
  const int gi = warp % G, kp = warp / G;
  for (int pass0 = 0; pass0 < groups; pass0 += G * NGW) {
    float acc[NGW][MTMAX][4];
#pragma unroll
    for (int i = 0; i < NGW; ++i)
#pragma unroll
      for (int m = 0; m < MTMAX; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][m][e] = 0.f;
    for (int kc0 = 0; kc0 < K; kc0 += KC) {
      const int kc = min(KC, K - kc0), C32 = kc / 32;
      __syncthreads();                                                  // previous chunk fully consumed (and rs written)
      for (int i = threadIdx.x; i < mt * 16 * (kc / 8); i += kMKThreads) {
        const int b = i / (kc / 8), kg = i % (kc / 8);
        const uint4 v = b < B ? linear_act8(x, b, kc0 / 8 + kg, act, ldx8, kvec, rs, nw) : make_uint4(0, 0, 0, 0);
        const int m = b >> 4, r = b & 15;
        frag[(((m * C32 + (kg >> 2)) * 32 + (r & 7) * 4 + (kg & 3)) << 1) + (r >> 3)] = v;
      }
      __syncthreads();
      for (int c0 = kp * U; c0 < C32; c0 += U * kparts) {
        uint4 wv[U][NGW];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int i = 0; i < NGW; ++i) {
            const int ng = pass0 + gi + i * G;
            const bool ok = (c0 + u < C32) && (ng < groups);
            wv[u][i] = ok ? ptx::ld_nc_v4(W + static_cast<size_t>(n0 + ng * 8 + g) * kvec + (kc0 / 8) + (c0 + u) * 4 + tig)
                          : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (c0 + u < C32) {
#pragma unroll
            for (int m = 0; m < MTMAX; ++m) {
              if (m < mt) {
                const uint4* fp = frag + (((m * C32 + c0 + u) * 32 + lane) << 1);
                const uint4 X = fp[0], Y = fp[1];
#pragma unroll
                for (int i = 0; i < NGW; ++i) {
                  mma_bf16_16816(acc[i][m], X.x, Y.x, X.y, Y.y, wv[u][i].x, wv[u][i].y);
                  mma_bf16_16816(acc[i][m], X.z, Y.z, X.w, Y.w, wv[u][i].z, wv[u][i].w);
                }
              }
            }
          }
        }
      }
    }
    if (kparts > 1) {
      // K was split between the warps of a group: partial sums of warps kp > 0 go through shared memory to warp kp == 0
      __syncthreads();                                                  // every warp is done with the staged fragments
      float* part = reinterpret_cast<float*>(smem);                     // [NW][NGW][MTMAX][4][32]
      if (kp > 0) {
#pragma unroll
        for (int i = 0; i < NGW; ++i)
#pragma unroll
          for (int m = 0; m < MTMAX; ++m)
#pragma unroll
            for (int e = 0; e < 4; ++e) part[(((warp * NGW + i) * MTMAX + m) * 4 + e) * 32 + lane] = acc[i][m][e];
      }
      __syncthreads();
      if (kp == 0) {
        for (int q = 1; q < kparts; ++q) {
          const int w2 = q * G + gi;
#pragma unroll
          for (int i = 0; i < NGW; ++i)
#pragma unroll
            for (int m = 0; m < MTMAX; ++m)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[i][m][e] += part[(((w2 * NGW + i) * MTMAX + m) * 4 + e) * 32 + lane];
        }
      }
    }
    // D fragment: c0, c1 = row g, columns tig*2 + {0, 1}; c2, c3 = row g + 8
    if (kp == 0) {
#pragma unroll
      for (int i = 0; i < NGW; ++i) {
        const int ng = pass0 + gi + i * G;
        if (ng < groups) {
          const int col = n0 + ng * 8 + tig * 2;
#pragma unroll
          for (int m = 0; m < MTMAX; ++m) {
            const int r0 = m * 16 + g, r1 = r0 + 8;
            if (r0 < B) *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(r0) * ldo + col) = ptx::pack_bf16x2(acc[i][m][0], acc[i][m][1]);
            if (r1 < B) *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(r1) * ldo + col) = ptx::pack_bf16x2(acc[i][m][2], acc[i][m][3]);
          }
        }
      }
    }
  }
  __syncthreads();        // the next task may restage smem
}

template <bool kTensorCore>
TD_DEVICE void task_linear(const MKParams& p, const Task& t, uint8_t* smem) {
  if constexpr (kTensorCore) {
    if (p.B > kMaxB) {
      if (t.a[6] > 8 * (kMKThreads / 32)) linear_mma<2>(p, t, smem);      // more than 8 column groups: two per warp
      else linear_mma<1>(p, t, smem);
      return;
    }
  }
  const uint4* x = (const uint4*)p.ptrs[t.a[0]];
  const uint4* W = (const uint4*)p.ptrs[t.a[1]];
  __nv_bfloat16* out = (__nv_bfloat16*)p.ptrs[t.a[2]];
  const int K = t.a[3], ldo = t.a[4], n0 = t.a[5], n_cnt = t.a[6], act = t.a[7], ldx = t.a[8];
  const int kvec = K / 8;
  const int B = p.B;
  uint4* xs = reinterpret_cast<uint4*>(smem);
  if (act == 2) {
    // RMSNorm fused into the operand staging (a[9] = norm weight, a[10] = eps): every CTA normalises the (tiny) activation
    // row itself, which removes a single-CTA task and a grid-wide dependency per norm
    const uint4* nw = (const uint4*)p.ptrs[t.a[9]];
    const float eps = __int_as_float(t.a[10]);
    float* red = reinterpret_cast<float*>(smem + static_cast<size_t>(B) * kvec * 16);       // [kMaxB][8 warps]
    float ssq[kMaxB];
#pragma unroll
    for (int b = 0; b < kMaxB; ++b) {
      ssq[b] = 0.f;
      if (b < B)
        for (int kv = threadIdx.x; kv < kvec; kv += kMKThreads) {
          float f[8];
          unpack8(x[b * (ldx / 8) + kv], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) ssq[b] += f[e] * f[e];
        }
      ssq[b] = warp_sum(ssq[b]);
      if ((threadIdx.x & 31) == 0 && b < B) red[b * 8 + (threadIdx.x >> 5)] = ssq[b];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B * kvec; i += kMKThreads) {
      const int b = i / kvec, kv = i % kvec;
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kMKThreads / 32; ++w) tot += red[b * 8 + w];
      const float rs = rsqrtf(tot / static_cast<float>(K) + eps);
      float f[8], g[8];
      unpack8(x[b * (ldx / 8) + kv], f);
      unpack8(nw[kv], g);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = f[e] * rs * g[e];
      xs[i] = pack8(f);
    }
  } else
  for (int i = threadIdx.x; i < B * kvec; i += kMKThreads) {
    const int b = i / kvec, kv = i % kvec;
    if (act == 0) xs[i] = x[b * (ldx / 8) + kv];
    else {
      float g[8], u[8];
      unpack8(x[b * (ldx / 8) + kv], g);
      unpack8(x[b * (ldx / 8) + kvec + kv], u);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = g[e] / (1.f + __expf(-g[e])) * u[e];
      xs[i] = pack8(g);
    }
  }
  __syncthreads();
  if (B == 1) linear_rows<1, 4>(W, xs, out, B, kvec, ldo, n0, n_cnt);
  else if (B <= 4) linear_rows<4, 2>(W, xs, out, B, kvec, ldo, n0, n_cnt);
  else linear_rows<kMaxB, 2>(W, xs, out, B, kvec, ldo, n0, n_cnt);
}

// ---- QKROPE: q/k RMSNorm + RoPE + KV append.  a: qkv, q_out, kcache, vcache, qn(-1), kn(-1), pos, Hq, Hkv, max_len, eps, theta ----
// kMode 0: dense cache [B, max_len, Hkv, D];  1: paged (see T_QKROPE_PAGED);  2: prefill split -- rows are B x S tokens, a[9] = S, the
// position of token (b, s) is pos[b] + s (pos = tokens already in the cache) and k / v go to dense [B * S, Hkv, D] outputs.
template <int kMode>
TD_DEVICE void task_qkrope(const MKParams& p, const Task& t) {
  const uint2* qkv = (const uint2*)p.ptrs[t.a[0]];
  uint2* q_out = (uint2*)p.ptrs[t.a[1]];
  uint2* kc = (uint2*)p.ptrs[t.a[2]];
  uint2* vc = (uint2*)p.ptrs[t.a[3]];
  const uint2* qn = t.a[4] >= 0 ? (const uint2*)p.ptrs[t.a[4]] : nullptr;
  const uint2* kn = t.a[5] >= 0 ? (const uint2*)p.ptrs[t.a[5]] : nullptr;
  const int* pos = (const int*)p.ptrs[t.a[6]];
  const int Hq = t.a[7], Hkv = t.a[8], max_len = t.a[9];
  const float eps = __int_as_float(t.a[10]), theta = __int_as_float(t.a[11]);
  const int heads = Hq + 2 * Hkv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int wi = warp; wi < p.B * heads; wi += kMKThreads / 32) {
    const int b = wi / heads, h = wi % heads;
    const uint2 raw = qkv[(static_cast<size_t>(b) * heads + h) * 32 + lane];
    float f[4] = {ptx::bf16_lo(raw.x), ptx::bf16_hi(raw.x), ptx::bf16_lo(raw.y), ptx::bf16_hi(raw.y)};
    const int ps = (kMode == 2) ? pos[b / max_len] + b % max_len : pos[b];
    const bool is_v = h >= Hq + Hkv;
    if (!is_v) {
      const uint2* nw = (h < Hq) ? qn : kn;
      if (nw) {
        float ss = warp_sum(f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3]);
        const float inv = rsqrtf(ss / 128.f + eps);
        const uint2 wr = nw[lane];
        const float g[4] = {ptx::bf16_lo(wr.x), ptx::bf16_hi(wr.x), ptx::bf16_lo(wr.y), ptx::bf16_hi(wr.y)};
#pragma unroll
        for (int e = 0; e < 4; ++e) f[e] = __bfloat162float(__float2bfloat16(f[e] * inv * g[e]));
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float other = __shfl_xor_sync(0xffffffffu, f[e], 16);
        const int d = (lane & 15) * 4 + e;
        float sn, cs;
        sincosf(static_cast<float>(ps) * powf(theta, -static_cast<float>(2 * d) / 128.f), &sn, &cs);
        f[e] = (lane < 16) ? (f[e] * cs - other * sn) : (f[e] * cs + other * sn);
      }
    }
    uint2 o;
    o.x = ptx::pack_bf16x2(f[0], f[1]); o.y = ptx::pack_bf16x2(f[2], f[3]);
    if (h < Hq) q_out[(static_cast<size_t>(b) * Hq + h) * 32 + lane] = o;
    else {
      const int kvh = is_v ? h - Hq - Hkv : h - Hq;
      size_t tok = static_cast<size_t>(b) * max_len + ps;
      if constexpr (kMode == 2) tok = b;
      if constexpr (kMode == 1) {      // max_len packs (page_size, max_pages); pages of a sequence are listed in its block-table row
        const int page_size = max_len & 0xFFFF, max_pages = max_len >> 16;
        const int* bt = (const int*)p.ptrs[t.a[3] + 1];
        tok = static_cast<size_t>(bt[b * max_pages + ps / page_size]) * page_size + ps % page_size;
      }
      (is_v ? vc : kc)[(tok * Hkv + kvh) * 32 + lane] = o;
    }
  }
}

// ---- ATTN: GQA decode for (b, kv head) over the whole context.  a: q, kcache, vcache, pos, out, b, kvh, Hq, Hkv, max_len, scale ----
template <bool kPaged>
TD_DEVICE void task_attn(const MKParams& p, const Task& t, uint8_t* smem) {
  const uint2* q = (const uint2*)p.ptrs[t.a[0]];
  const uint2* kc = (const uint2*)p.ptrs[t.a[1]];
  const uint2* vc = (const uint2*)p.ptrs[t.a[2]];
  const int* pos = (const int*)p.ptrs[t.a[3]];
  uint2* out = (uint2*)p.ptrs[t.a[4]];
  const int b = t.a[5], kvh = t.a[6], Hq = t.a[7], Hkv = t.a[8], max_len = t.a[9];
  const float scale = __int_as_float(t.a[10]);
  const int split = t.a[11] & 0xFFFF, n_splits = max(1, t.a[11] >> 16);   // split-KV: this task owns keys [j0, j1)
  const int G = Hq / Hkv;                       // <= 8
  const int len = pos[b] + 1;
  const int per = (len + n_splits - 1) / n_splits;
  const int j0 = split * per, j1 = min(len, j0 + per);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = kMKThreads / 32;
  float qf[8][4], m[8], l[8], o[8][4];
  for (int g = 0; g < 8; ++g) {
    m[g] = -INFINITY; l[g] = 0.f; o[g][0] = o[g][1] = o[g][2] = o[g][3] = 0.f;
    if (g < G) {
      const uint2 r = q[(static_cast<size_t>(b) * Hq + kvh * G + g) * 32 + lane];
      qf[g][0] = ptx::bf16_lo(r.x) * scale; qf[g][1] = ptx::bf16_hi(r.x) * scale; qf[g][2] = ptx::bf16_lo(r.y) * scale; qf[g][3] = ptx::bf16_hi(r.y) * scale;
    }
  }
  for (int j = j0 + warp; j < j1; j += NW) {
    size_t tok = static_cast<size_t>(b) * max_len + j;
    if constexpr (kPaged) {
      const int page_size = max_len & 0xFFFF, max_pages = max_len >> 16;
      const int* bt = (const int*)p.ptrs[t.a[2] + 1];
      tok = static_cast<size_t>(bt[b * max_pages + j / page_size]) * page_size + j % page_size;
    }
    const size_t row = (tok * Hkv + kvh) * 32 + lane;
    const uint2 kr = kc[row], vr = vc[row];
    const float kf[4] = {ptx::bf16_lo(kr.x), ptx::bf16_hi(kr.x), ptx::bf16_lo(kr.y), ptx::bf16_hi(kr.y)};
    const float vf[4] = {ptx::bf16_lo(vr.x), ptx::bf16_hi(vr.x), ptx::bf16_lo(vr.y), ptx::bf16_hi(vr.y)};
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (g < G) {
        const float s = warp_sum(qf[g][0] * kf[0] + qf[g][1] * kf[1] + qf[g][2] * kf[2] + qf[g][3] * kf[3]);
        const float mn = fmaxf(m[g], s);
        const float corr = __expf(m[g] - mn), pj = __expf(s - mn);
        l[g] = l[g] * corr + pj;
#pragma unroll
        for (int e = 0; e < 4; ++e) o[g][e] = o[g][e] * corr + pj * vf[e];
        m[g] = mn;
      }
    }
  }
  float* sm_m = reinterpret_cast<float*>(smem);             // [NW][8]
  float* sm_l = sm_m + NW * 8;                              // [NW][8]
  float* sm_o = sm_l + NW * 8;                              // [NW][8][128]
  for (int g = 0; g < G; ++g) {
    if (lane == 0) { sm_m[warp * 8 + g] = m[g]; sm_l[warp * 8 + g] = l[g]; }
#pragma unroll
    for (int e = 0; e < 4; ++e) sm_o[(warp * 8 + g) * 128 + lane * 4 + e] = o[g][e];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < G * 64; idx += kMKThreads) {    // two output elements per thread
    const int g = idx / 64, d2 = idx % 64;
    float mm = -INFINITY;
    for (int w = 0; w < NW; ++w) mm = fmaxf(mm, sm_m[w * 8 + g]);
    float ll = 0.f, o0 = 0.f, o1 = 0.f;
    for (int w = 0; w < NW; ++w) {
      const float c = (sm_m[w * 8 + g] == -INFINITY) ? 0.f : __expf(sm_m[w * 8 + g] - mm);
      ll += sm_l[w * 8 + g] * c; o0 += sm_o[(w * 8 + g) * 128 + 2 * d2] * c; o1 += sm_o[(w * 8 + g) * 128 + 2 * d2 + 1] * c;
    }
    if (n_splits > 1) {
      // partial (m, l, o) of this split: part[(((b * Hkv + kvh) * n_splits + split) * 8 + g) * 130 + {0: m, 1: l, 2..129: o}]
      float* part = reinterpret_cast<float*>(out) + ((static_cast<size_t>(b * Hkv + kvh) * n_splits + split) * 8 + g) * 130;
      if (d2 == 0) { part[0] = mm; part[1] = ll; }
      part[2 + 2 * d2] = o0; part[3 + 2 * d2] = o1;
    } else {
      const float inv = ll > 0.f ? 1.f / ll : 0.f;
      reinterpret_cast<uint32_t*>(out)[(static_cast<size_t>(b) * Hq + kvh * G + g) * 64 + d2] = ptx::pack_bf16x2(o0 * inv, o1 * inv);
    }
  }
}

// ---- FLASH_ATTN: prefill attention for 128 query rows of one (batch, q head) on warp-level tensor cores (FA2 dataflow on mma.sync).
// a: q, k, v, out, b, h, q block | causal << 30, S, Hq | Hkv << 16, q token stride / 128 | kv token stride / 128 << 16, scale, soft cap.
// Same statements as the DSL kernel triton_dist/lk/kernels/flash_mma.py (which the CPU interpreter runs against the fp32 reference):
// a lane's 16-byte loads of its two query rows are its A fragments (the order of d inside a 32-wide chunk is permuted identically for
// q and k); K row-major in shared memory, V transposed as packed key pairs; scores stay in registers between the two MMAs.
TD_DEVICE void task_flash_attn(const MKParams& p, const Task& t, uint8_t* smem) {
  constexpr int D = 128, BKV = 64, KS = 80, VS = 36, BQ = kMKThreads / 32 * 16;
  const __nv_bfloat16* q = (const __nv_bfloat16*)p.ptrs[t.a[0]];
  const __nv_bfloat16* k = (const __nv_bfloat16*)p.ptrs[t.a[1]];
  const __nv_bfloat16* v = (const __nv_bfloat16*)p.ptrs[t.a[2]];
  uint32_t* ow = (uint32_t*)p.ptrs[t.a[3]];
  const int b = t.a[4], h = t.a[5], qb = t.a[6] & 0xFFFFFF, S = t.a[7], Hq = t.a[8] & 0xFFFF, Hkv = t.a[8] >> 16;
  const bool causal = (t.a[6] >> 30) & 1;
  const int q_ts = (t.a[9] & 0xFFFF) * D, kv_ts = (t.a[9] >> 16) * D, o_ts = Hq * D;
  const float scale = __int_as_float(t.a[10]), softcap = __int_as_float(t.a[11]);
  uint32_t* Ksm = reinterpret_cast<uint32_t*>(smem);
  uint32_t* Vt = Ksm + BKV * KS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const int kvh = h / (Hq / Hkv);
  const int q0 = qb * BQ, r0 = q0 + warp * 16 + g;
  uint32_t qf[32];
#pragma unroll
  for (int hi = 0; hi < 2; ++hi) {
    const int row = r0 + 8 * hi;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 qv = make_uint4(0, 0, 0, 0);
      if (row < S) qv = *reinterpret_cast<const uint4*>(q + ((static_cast<size_t>(b) * S + row) * q_ts + h * D + c * 32 + tig * 8));
      qf[(2 * c) * 4 + hi] = qv.x; qf[(2 * c) * 4 + 2 + hi] = qv.y; qf[(2 * c + 1) * 4 + hi] = qv.z; qf[(2 * c + 1) * 4 + 2 + hi] = qv.w;
    }
  }
  float o[64], s[32];
#pragma unroll
  for (int e = 0; e < 64; ++e) o[e] = 0.f;
  float m0 = -1.0e30f, m1 = -1.0e30f, l0 = 0.f, l1 = 0.f;
  const int kv_end = causal ? min(S, q0 + BQ) : S;
  const __nv_bfloat16* kbase = k + (static_cast<size_t>(b) * S * kv_ts + kvh * D);
  const __nv_bfloat16* vbase = v + (static_cast<size_t>(b) * S * kv_ts + kvh * D);
  for (int kv0 = 0; kv0 < kv_end; kv0 += BKV) {
    __syncthreads();
    for (int i = tid; i < BKV * 16; i += kMKThreads) {
      const int key = i >> 4, ch = i & 15;
      uint4 kq = make_uint4(0, 0, 0, 0);
      if (kv0 + key < S) kq = *reinterpret_cast<const uint4*>(kbase + (static_cast<size_t>(kv0 + key) * kv_ts + ch * 8));
      *reinterpret_cast<uint4*>(Ksm + key * KS + ch * 4) = kq;
    }
    for (int i = tid; i < 32 * 16; i += kMKThreads) {
      const int pj = i & 31, ch = i >> 5;
      uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
      if (kv0 + 2 * pj < S) va = *reinterpret_cast<const uint4*>(vbase + (static_cast<size_t>(kv0 + 2 * pj) * kv_ts + ch * 8));
      if (kv0 + 2 * pj + 1 < S) vb = *reinterpret_cast<const uint4*>(vbase + (static_cast<size_t>(kv0 + 2 * pj + 1) * kv_ts + ch * 8));
      uint32_t* dst = Vt + (ch * 8) * VS + pj;
      dst[0 * VS] = (va.x & 0xFFFFu) | (vb.x << 16); dst[1 * VS] = (va.x >> 16) | (vb.x & 0xFFFF0000u);
      dst[2 * VS] = (va.y & 0xFFFFu) | (vb.y << 16); dst[3 * VS] = (va.y >> 16) | (vb.y & 0xFFFF0000u);
      dst[4 * VS] = (va.z & 0xFFFFu) | (vb.z << 16); dst[5 * VS] = (va.z >> 16) | (vb.z & 0xFFFF0000u);
      dst[6 * VS] = (va.w & 0xFFFFu) | (vb.w << 16); dst[7 * VS] = (va.w >> 16) | (vb.w & 0xFFFF0000u);
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 32; ++e) s[e] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 kw = *reinterpret_cast<const uint4*>(Ksm + (nt * 8 + g) * KS + c * 16 + tig * 4);
        mma_bf16_16816p(s + nt * 4, qf[(2 * c) * 4 + 0], qf[(2 * c) * 4 + 1], qf[(2 * c) * 4 + 2], qf[(2 * c) * 4 + 3], kw.x, kw.y);
        mma_bf16_16816p(s + nt * 4, qf[(2 * c + 1) * 4 + 0], qf[(2 * c + 1) * 4 + 1], qf[(2 * c + 1) * 4 + 2], qf[(2 * c + 1) * 4 + 3], kw.z, kw.w);
      }
    }
    float mx0 = -1.0e30f, mx1 = -1.0e30f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = s[nt * 4 + j] * scale;
        if (softcap > 0.f) x = softcap * tanhf(x / softcap);
        x *= 1.4426950408889634f;
        const int kj = kv0 + nt * 8 + tig * 2 + (j & 1), qi = r0 + 8 * (j >> 1);
        if (kj >= S || (causal && kj > qi)) x = -1.0e30f;
        s[nt * 4 + j] = x;
        if (j < 2) mx0 = fmaxf(mx0, x); else mx1 = fmaxf(mx1, x);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float corr0 = exp2f(m0 - mn0), corr1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float pv = 0.f;                                     // masked scores contribute exactly 0 (also when a whole row is masked)
        if (s[nt * 4 + j] > -1.0e29f) pv = exp2f(s[nt * 4 + j] - (j < 2 ? mn0 : mn1));
        s[nt * 4 + j] = pv;
        if (j < 2) rs0 += pv; else rs1 += pv;
      }
    }
    l0 = l0 * corr0 + rs0; l1 = l1 * corr1 + rs1;         // per-lane partial row sums; reduced over the quad once, at the end
#pragma unroll
    for (int dt = 0; dt < 16; ++dt) { o[dt * 4 + 0] *= corr0; o[dt * 4 + 1] *= corr0; o[dt * 4 + 2] *= corr1; o[dt * 4 + 3] *= corr1; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t pa0 = ptx::pack_bf16x2(s[(2 * kk) * 4 + 0], s[(2 * kk) * 4 + 1]), pa1 = ptx::pack_bf16x2(s[(2 * kk) * 4 + 2], s[(2 * kk) * 4 + 3]);
      const uint32_t pa2 = ptx::pack_bf16x2(s[(2 * kk + 1) * 4 + 0], s[(2 * kk + 1) * 4 + 1]), pa3 = ptx::pack_bf16x2(s[(2 * kk + 1) * 4 + 2], s[(2 * kk + 1) * 4 + 3]);
#pragma unroll
      for (int dt = 0; dt < 16; ++dt) {
        const uint32_t* vw = Vt + (dt * 8 + g) * VS + kk * 8 + tig;
        mma_bf16_16816p(o + dt * 4, pa0, pa1, pa2, pa3, vw[0], vw[4]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
  for (int dt = 0; dt < 16; ++dt) {
    const int col = h * D + dt * 8 + tig * 2;
    if (r0 < S) ow[((static_cast<size_t>(b) * S + r0) * o_ts + col) >> 1] = ptx::pack_bf16x2(o[dt * 4 + 0] * inv0, o[dt * 4 + 1] * inv0);
    if (r0 + 8 < S) ow[((static_cast<size_t>(b) * S + r0 + 8) * o_ts + col) >> 1] = ptx::pack_bf16x2(o[dt * 4 + 2] * inv1, o[dt * 4 + 3] * inv1);
  }
}

// ---- stand-alone element-wise tasks of the reference's builder (the dense model fuses them into neighbouring tasks) ----
// SILU_MUL: out[b, i] = silu(x[b, i]) * x[b, I + i].  a: x, out, I, v0, v1 (16-byte vector range of the [B, I] output)
TD_DEVICE void task_silu_mul(const MKParams& p, const Task& t) {
  const uint4* x = (const uint4*)p.ptrs[t.a[0]];
  uint4* out = (uint4*)p.ptrs[t.a[1]];
  const int ivec = t.a[2] / 8;
  for (int v = t.a[3] + threadIdx.x; v < t.a[4]; v += kMKThreads) {
    const int b = v / ivec, i = v % ivec;
    float g[8], u[8];
    unpack8(x[static_cast<size_t>(b) * 2 * ivec + i], g);
    unpack8(x[static_cast<size_t>(b) * 2 * ivec + ivec + i], u);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = g[e] / (1.f + __expf(-g[e])) * u[e];
    out[v] = pack8(g);
  }
}
// ADD: out = lhs + rhs over 16-byte vectors [v0, v1).  a: lhs, rhs, out, v0, v1
TD_DEVICE void task_add(const MKParams& p, const Task& t) {
  const uint4* a = (const uint4*)p.ptrs[t.a[0]];
  const uint4* b = (const uint4*)p.ptrs[t.a[1]];
  uint4* out = (uint4*)p.ptrs[t.a[2]];
  for (int v = t.a[3] + threadIdx.x; v < t.a[4]; v += kMKThreads) {
    float x[8], y[8];
    unpack8(a[v], x); unpack8(b[v], y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] += y[e];
    out[v] = pack8(x);
  }
}
// PREFETCH: pull bytes [off, off + n) of a weight into L2 ahead of the tasks that stream it.  a: w, off_kb, n_kb
TD_DEVICE void task_prefetch(const MKParams& p, const Task& t) {
  const char* w = reinterpret_cast<const char*>(p.ptrs[t.a[0]]) + static_cast<size_t>(t.a[1]) * 1024;
  const size_t bytes = static_cast<size_t>(t.a[2]) * 1024;
  for (size_t off = static_cast<size_t>(threadIdx.x) * 16384; off < bytes; off += static_cast<size_t>(kMKThreads) * 16384)
    ptx::prefetch_l2_bulk(w + off, static_cast<uint32_t>(min(bytes - off, size_t(16384))));
}

// ---- ATTN_COMBINE: LSE-merge of the split-KV partials of (b, kv head).  a: part, out, b, kvh, Hq, Hkv, n_splits ----
TD_DEVICE void task_attn_combine(const MKParams& p, const Task& t) {
  const float* part = (const float*)p.ptrs[t.a[0]];
  uint32_t* out = (uint32_t*)p.ptrs[t.a[1]];
  const int b = t.a[2], kvh = t.a[3], Hq = t.a[4], Hkv = t.a[5], n_splits = t.a[6];
  const int G = Hq / Hkv;
  for (int idx = threadIdx.x; idx < G * 64; idx += kMKThreads) {
    const int g = idx / 64, d2 = idx % 64;
    const float* base = part + (static_cast<size_t>(b * Hkv + kvh) * n_splits * 8 + g) * 130;
    float mm = -INFINITY;
    for (int s = 0; s < n_splits; ++s) mm = fmaxf(mm, base[static_cast<size_t>(s) * 8 * 130]);
    float ll = 0.f, o0 = 0.f, o1 = 0.f;
    for (int s = 0; s < n_splits; ++s) {
      const float* ps = base + static_cast<size_t>(s) * 8 * 130;
      const float c = (ps[0] == -INFINITY) ? 0.f : __expf(ps[0] - mm);
      ll += ps[1] * c; o0 += ps[2 + 2 * d2] * c; o1 += ps[3 + 2 * d2] * c;
    }
    const float inv = ll > 0.f ? 1.f / ll : 0.f;
    out[(static_cast<size_t>(b) * Hq + kvh * G + g) * 64 + d2] = ptx::pack_bf16x2(o0 * inv, o1 * inv);
  }
}

// ---- ALLREDUCE slice (+ residual):  res_out[v] = res[v] + sum_r part_r[v].  a: part(symm), flags(symm), res, res_out, v0, v1, op_id, n_slices, slice ----
TD_DEVICE void task_allreduce(const MKParams& p, const Task& t, uint32_t epoch) {
  uint4* part = (uint4*)p.ptrs[t.a[0]];
  uint32_t* flags = (uint32_t*)p.ptrs[t.a[1]];
  const uint4* res = (const uint4*)p.ptrs[t.a[2]];
  uint4* res_out = (uint4*)p.ptrs[t.a[3]];
  const int v0 = t.a[4], v1 = t.a[5], slice = t.a[8];
  const SymmCtx& c = p.symm;
  const int W = c.world;
  uint32_t* my_flags = flags + slice * W;                     // [W] one word per source rank
  if (W > 1) {
    // my partial slice is complete (scoreboard dependency) -> tell every rank, then wait for theirs
    if (threadIdx.x < W) {
      ptx::fence_acq_rel_sys();
      ptx::st_release_sys(symm_at(c, my_flags + c.rank, threadIdx.x), epoch);
      uint32_t v;
      do { v = ptx::ld_acquire_sys(my_flags + threadIdx.x); } while (static_cast<int32_t>(v - epoch) < 0);
    }
    __syncthreads();
  }
  for (int v = v0 + threadIdx.x; v < v1; v += kMKThreads) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (W > 1 && c.mc_base) {
      float f[8];
      unpack8(ptx::multimem_ld_reduce_bf16x8(symm_mc(c, part) + v), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = f[e];
    } else {
      for (int r = 0; r < W; ++r) {
        float f[8];
        unpack8(ptx::ld_relaxed_sys_v4(symm_at(c, part + v, (c.rank + r) % W)), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    }
    if (res) {
      float r[8];
      unpack8(res[v], r);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += r[e];
    }
    res_out[v] = pack8(acc);
  }
}

TD_DEVICE void prefetch_linear_weights(const MKParams& p, const Task& t) {
  if (t.type != T_LINEAR) return;
  const char* W = reinterpret_cast<const char*>(p.ptrs[t.a[1]]) + static_cast<size_t>(t.a[5]) * t.a[3] * 2;
  size_t bytes = static_cast<size_t>(t.a[6]) * t.a[3] * 2;
  for (size_t off = 0; off < bytes; off += 65536) ptx::prefetch_l2_bulk(W + off, static_cast<uint32_t>(min(bytes - off, size_t(65536))));
}

// Instantiations of the interpreter (kVariant bit 0: prefill task types FLASH_ATTN / QKROPE_SPLIT; bit 1: tensor-core LINEAR body for
// 9..64 tokens; bit 2: paged-KV task types).  Variant 0 -- decode with 1..8 tokens -- contains exactly the task bodies that have run on hardware; the prefill
// task's 96 accumulators and the mma.sync LINEAR body would otherwise shape the register allocation of the whole switch.
template <int kVariant>
__global__ void __launch_bounds__(kMKThreads, 1) mega_kernel(const MKParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float red[32];
  const uint32_t epoch = p.epoch[0] + 1;
  __shared__ int s_next;
  int qi = p.dynamic ? 0 : p.queue_off[blockIdx.x];
  const int q1 = p.dynamic ? p.num_tasks : p.queue_off[blockIdx.x + 1];
  while (true) {
    if (p.dynamic) {
      // runtime scheduler (reference core/scheduler.py: global queue + atomic_add): tasks are stored in topological
      // order, so whoever holds task k only ever waits for tasks < k, which are done or held by a running CTA
      if (threadIdx.x == 0) s_next = static_cast<int>(atomicAdd(p.epoch + 2, 1u));
      __syncthreads();
      qi = s_next;
      __syncthreads();
    }
    if (qi >= q1) break;
    const Task& t = p.tasks[qi];
    // weight prefetch (reference model_builder "prefetch" tasks): the W tile of a LINEAR task is one contiguous region;
    // pull it (and the tile of this CTA's next task) DRAM -> L2 before sitting on the dependency, so the stream of weight
    // bytes does not stop at phase boundaries
    if (threadIdx.x == 0) {
      prefetch_linear_weights(p, t);
      if (!p.dynamic && qi + 1 < q1) prefetch_linear_weights(p, p.tasks[qi + 1]);
    }
    if (t.dep_idx >= 0) {
      if (threadIdx.x == 0) {
        const uint32_t target = epoch * static_cast<uint32_t>(t.dep_count);
        while (static_cast<int32_t>(ptx::ld_acquire_gpu(p.sb + t.dep_idx) - target) < 0) {
        }
      }
      __syncthreads();
    }
    switch (t.type) {
      case T_RMSNORM: task_rmsnorm(p, t, red); break;
      case T_LINEAR: task_linear<(kVariant & 2) != 0>(p, t, smem); break;
      case T_QKROPE: task_qkrope<0>(p, t); break;
      case T_ATTN: task_attn<false>(p, t, smem); break;
      case T_QKROPE_PAGED: if constexpr ((kVariant & 4) != 0) task_qkrope<1>(p, t); break;
      case T_QKROPE_SPLIT: if constexpr ((kVariant & 1) != 0) task_qkrope<2>(p, t); break;
      case T_ATTN_PAGED: if constexpr ((kVariant & 4) != 0) task_attn<true>(p, t, smem); break;
      case T_FLASH_ATTN: if constexpr ((kVariant & 1) != 0) task_flash_attn(p, t, smem); break;
      case T_ATTN_COMBINE: task_attn_combine(p, t); break;
      case T_SILU_MUL: task_silu_mul(p, t); break;
      case T_ADD: task_add(p, t); break;
      case T_PREFETCH: task_prefetch(p, t); break;
      case T_ALLREDUCE: task_allreduce(p, t, epoch); break;
      default: break;
    }
    __syncthreads();
    if (t.sig_idx >= 0 && threadIdx.x == 0) {
      __threadfence();
      ptx::red_release_gpu_add(p.sb + t.sig_idx, 1u);
    }
    if (!p.dynamic) ++qi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.epoch + 1, 1u) == gridDim.x - 1) { p.epoch[1] = 0; p.epoch[2] = 0; __threadfence(); p.epoch[0] = epoch; }
  }
}

struct TdSymmArgs { long long rank, world; unsigned long long base, stride, mc_base; };

}  // namespace

struct TdMegaArgs {
  TdSymmArgs symm;
  const void* tasks; const void* queue_off; const void* ptrs; void* sb; void* epoch;
  long long B, grid, smem_bytes, dynamic, num_tasks;
};

TD_API int td_mega_task_size() { return (int)sizeof(Task); }

TD_API int td_mega_launch(const TdMegaArgs* a, void* stream) {
  if (a->B < 1 || a->B > kMmaMaxB) { td::drv::set_error("megakernel: batch must be 1..64 (1..8: GEMV tasks, 9..64: tensor-core tasks)"); return -1; }
  if (a->B > kMaxB && a->smem_bytes < kMmaStageBytes + 256) { td::drv::set_error("megakernel: tensor-core linears need 128 KB + 256 B of dynamic shared memory"); return -1; }
  MKParams p;
  p.tasks = (const Task*)a->tasks; p.queue_off = (const int*)a->queue_off; p.ptrs = (void* const*)a->ptrs;
  p.sb = (uint32_t*)a->sb; p.epoch = (uint32_t*)a->epoch;
  p.symm.rank = (int)a->symm.rank; p.symm.world = (int)a->symm.world; p.symm.base = a->symm.base; p.symm.stride = a->symm.stride; p.symm.mc_base = a->symm.mc_base;
  p.B = (int)a->B; p.dynamic = (int)(a->dynamic & 1); p.num_tasks = (int)a->num_tasks;
  // bit 1 of `dynamic`: the task list contains prefill task types; batches above 8 tokens need the tensor-core LINEAR body
  const int variant = ((a->dynamic & 2) ? 1 : 0) | (a->B > kMaxB ? 2 : 0) | ((a->dynamic & 4) ? 4 : 0);      // bit 2 of `dynamic`: paged-KV tasks
  using KernelFn = void (*)(const MKParams);
  static const KernelFn kernels[8] = {mega_kernel<0>, mega_kernel<1>, mega_kernel<2>, mega_kernel<3>,
                                      mega_kernel<4>, mega_kernel<5>, mega_kernel<6>, mega_kernel<7>};
  static long long smem_set[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (a->smem_bytes > smem_set[variant]) {
    TD_CUDA_CHECK(cudaFuncSetAttribute(kernels[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a->smem_bytes));
    smem_set[variant] = a->smem_bytes;
  }
  kernels[variant]<<<(int)a->grid, kMKThreads, (size_t)a->smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
