// Expert-parallel dispatch / combine, throughput ("normal") mode with token saving, intra-node over NVLink.
//
// Reference: kernels/nvidia/ep_a2a_intra_node.py (kernel_dispatch_token_intra_node :39-125 -- one warp per token, a token
// that goes to several experts of the SAME rank is sent once plus an index per expert; kernel_combine_token_intra_node
// :219-289, local pre-combine :174), ep_a2a.py / layers/nvidia/ep_a2a_layer.py.  B200-first differences:
//   * the payload row lands in   rx[parity][src rank][slot]   on the destination and is never moved again: the expert GEMM
//     fetches rows with TMA tile::gather4 through an index list (csrc/gemm_sm100.cuh, a_gather), so there is no
//     receive-side compaction pass;
//   * per (token, destination rank): one payload row + one row descriptor {token, first pair, #pairs}; per (token, k): one
//     16-byte pair descriptor {row slot, local expert, flat source pair id, routing weight} -- the pairs of a token are
//     contiguous, which lets the combine side reduce them locally;
//   * combine = local weighted pre-reduce of a token's expert outputs on the expert rank, ONE row per (token, rank) sent
//     back, final sum over the <= topk ranks on the token's owner;
//   * arrival = 64-bit release stores (phase << 32 | count), parity double buffering, device-resident phase counter:
//     no barrier, no reset, CUDA-graph replayable.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "runtime/driver.h"
#include "td/primitives.cuh"
#include "td/ptx.cuh"

using namespace td;

namespace {

constexpr int kThreads = 256;

struct DispParams {
  SymmCtx symm;
  int T, H, topk, E, epr, T_max, P_max;
  const uint4* x;                  // [T, H] bf16
  const int* topk_idx;             // [T, topk]
  const float* topk_w;             // [T, topk]
  char* rx;                        // symmetric [2][W][T_max][H] bf16 payload rows
  long long rx_buf_bytes;
  int4* rp;                        // symmetric [2][W][P_max] pair descriptors
  int4* rmeta;                     // symmetric [2][W][T_max] row descriptors
  unsigned long long* rflag;       // symmetric [2][W][2]: (phase << 32 | rows), (phase << 32 | pairs)
  int* send_rows;                  // local [W], zero between calls
  int* send_pairs;                 // local [W]
  uint32_t* phase;                 // local [0] calls, [1] send counter, [2] exit counter
  // local outputs
  int* pair_expert;                // [W * P_max] local expert of every received pair, -1 = empty slot
  int* pair_row;                   // [W * P_max] row of rx (src * T_max + slot), -1 = empty slot
  int* rcnt;                       // [W][2] rows / pairs received from every source
};

__global__ void __launch_bounds__(kThreads, 1) ep_dispatch_normal_kernel(const DispParams p) {
  const SymmCtx& c = p.symm;
  const int W = c.world, me = c.rank;
  const uint32_t ph = p.phase[0] + 1;
  const uint32_t par = ph & 1u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = kThreads / 32;
  const int vec_per_row = p.H / 8;
  char* rx_par = p.rx + par * p.rx_buf_bytes;
  int4* rp_par = p.rp + static_cast<size_t>(par) * W * p.P_max;
  int4* rm_par = p.rmeta + static_cast<size_t>(par) * W * p.T_max;

  // ---------------- send: one warp per token ----------------
  for (int t = blockIdx.x * wpc + warp; t < p.T; t += gridDim.x * wpc) {
    int e = -1; float w = 0.f;
    if (lane < p.topk) { e = p.topk_idx[t * p.topk + lane]; w = p.topk_w[t * p.topk + lane]; }
    const bool valid = (e >= 0 && e < p.E);
    const int d = valid ? e / p.epr : -1 - lane;                  // invalid lanes get unique keys (no grouping)
    const unsigned same = __match_any_sync(0xffffffffu, d);       // lanes of this token that go to the same rank
    const int leader_lane = __ffs(same) - 1;
    const bool leader = valid && lane == leader_lane;
    int slot = 0, pbase = 0;
    if (leader) {
      slot = atomicAdd(p.send_rows + d, 1);
      pbase = atomicAdd(p.send_pairs + d, __popc(same));
    }
    slot = __shfl_sync(0xffffffffu, slot, leader_lane);
    pbase = __shfl_sync(0xffffffffu, pbase, leader_lane);
    if (valid) {
      const int r = __popc(same & ((1u << lane) - 1u));
      int4* dst = symm_at(c, rp_par, d) + static_cast<size_t>(me) * p.P_max + pbase + r;
      ptx::st_v4(dst, make_uint4(static_cast<uint32_t>(slot), static_cast<uint32_t>(e % p.epr),
                                 static_cast<uint32_t>(t * p.topk + lane), __float_as_uint(w)));
    }
    if (leader) {
      int4* dst = symm_at(c, rm_par, d) + static_cast<size_t>(me) * p.T_max + slot;
      ptx::st_v4(dst, make_uint4(static_cast<uint32_t>(t), static_cast<uint32_t>(pbase), static_cast<uint32_t>(__popc(same)), 0u));
    }
    // payload: once per distinct destination rank
    unsigned leaders = __ballot_sync(0xffffffffu, leader);
    const uint4* row = p.x + static_cast<size_t>(t) * vec_per_row;
    while (leaders) {
      const int L = __ffs(leaders) - 1;
      leaders &= leaders - 1;
      const int dL = __shfl_sync(0xffffffffu, d, L), sL = __shfl_sync(0xffffffffu, slot, L);
      uint4* out = reinterpret_cast<uint4*>(symm_at(c, rx_par, dL) + (static_cast<size_t>(me) * p.T_max + sL) * p.H * 2);
      for (int v = lane; v < vec_per_row; v += 32) ptx::st_v4(out + v, row[v]);
    }
  }
  __syncthreads();
  // ---------------- publish counts (last CTA to finish sending) ----------------
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    ptx::fence_acq_rel_sys();
    s_last = (atomicAdd(p.phase + 1, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (static_cast<int>(threadIdx.x) < W) {
      const int d = threadIdx.x;
      const int rows = p.send_rows[d], pairs = p.send_pairs[d];
      p.send_rows[d] = 0; p.send_pairs[d] = 0;
      unsigned long long* f = symm_at(c, p.rflag + (static_cast<size_t>(par) * W + me) * 2, d);
      ptx::fence_acq_rel_sys();
      ptx::st_release_sys(reinterpret_cast<uint64_t*>(f), (static_cast<uint64_t>(ph) << 32) | static_cast<uint32_t>(rows));
      ptx::st_release_sys(reinterpret_cast<uint64_t*>(f + 1), (static_cast<uint64_t>(ph) << 32) | static_cast<uint32_t>(pairs));
    }
    if (threadIdx.x == 0) p.phase[1] = 0;
  }
  // ---------------- receive: per source, the index lists the expert GEMM gathers through ----------------
  __shared__ int s_rows, s_pairs;
  for (int src = blockIdx.x; src < W; src += gridDim.x) {
    if (threadIdx.x == 0) {
      const unsigned long long* f = p.rflag + (static_cast<size_t>(par) * W + src) * 2;
      uint64_t a, b;
      do { a = ptx::ld_acquire_sys(reinterpret_cast<const uint64_t*>(f)); } while ((a >> 32) != ph);
      do { b = ptx::ld_acquire_sys(reinterpret_cast<const uint64_t*>(f + 1)); } while ((b >> 32) != ph);
      s_rows = static_cast<int>(a & 0xffffffffu); s_pairs = static_cast<int>(b & 0xffffffffu);
      p.rcnt[src * 2] = s_rows; p.rcnt[src * 2 + 1] = s_pairs;
    }
    __syncthreads();
    const int n_pairs = s_pairs;
    const int4* pr = rp_par + static_cast<size_t>(src) * p.P_max;
    for (int i = threadIdx.x; i < p.P_max; i += kThreads) {
      int le = -1, row = -1;
      if (i < n_pairs) {
        const uint4 d4 = ptx::ld_relaxed_sys_v4(pr + i);
        le = static_cast<int>(d4.y); row = src * p.T_max + static_cast<int>(d4.x);
      }
      p.pair_expert[src * p.P_max + i] = le;
      p.pair_row[src * p.P_max + i] = row;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.phase + 2, 1u) == gridDim.x - 1) { p.phase[2] = 0; __threadfence(); p.phase[0] = ph; }
  }
}

struct CombParams {
  SymmCtx symm;
  int T, H, topk, E, epr, T_max, P_max;
  const uint4* y;                  // [W * P_max, H] bf16 expert outputs in received-pair order
  const int4* rp;                  // local view of the pair descriptors of the dispatch being combined: [W][P_max]
  const int4* rmeta;               // [W][T_max]
  const int* rcnt;                 // [W][2]
  const int* topk_idx;             // [T, topk] (this rank's tokens)
  char* comb;                      // symmetric [2][W (expert rank)][T_max][H] bf16
  long long comb_buf_bytes;
  uint32_t* comb_flag;             // symmetric [2][W]
  uint32_t* phase;                 // local [0] calls, [1] counter, [2] exit
  uint4* out;                      // [T, H]
};

__global__ void __launch_bounds__(kThreads, 1) ep_combine_normal_kernel(const CombParams p) {
  const SymmCtx& c = p.symm;
  const int W = c.world, me = c.rank;
  const uint32_t ph = p.phase[0] + 1;
  const uint32_t par = ph & 1u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = kThreads / 32;
  const int vec_per_row = p.H / 8;
  char* comb_par = p.comb + par * p.comb_buf_bytes;

  // ---------------- expert side: weighted pre-reduce of every received token, one row back per (token, rank) ----------------
  for (int u = blockIdx.x * wpc + warp; u < W * p.T_max; u += gridDim.x * wpc) {
    const int src = u / p.T_max, slot = u % p.T_max;
    if (slot >= p.rcnt[src * 2]) continue;
    const int4 meta = p.rmeta[static_cast<size_t>(src) * p.T_max + slot];
    const int t = meta.x, first = meta.y, n = meta.z;
    uint4* dst = reinterpret_cast<uint4*>(symm_at(c, comb_par, src) + (static_cast<size_t>(me) * p.T_max + t) * p.H * 2);
    for (int v = lane; v < vec_per_row; v += 32) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < n; ++i) {
        const size_t pair = static_cast<size_t>(src) * p.P_max + first + i;
        const float w = __int_as_float(p.rp[pair].w);
        const uint4 xv = p.y[pair * vec_per_row + v];
        const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[2 * q] += w * ptx::bf16_lo(xs[q]); acc[2 * q + 1] += w * ptx::bf16_hi(xs[q]); }
      }
      uint4 o;
      o.x = ptx::pack_bf16x2(acc[0], acc[1]); o.y = ptx::pack_bf16x2(acc[2], acc[3]);
      o.z = ptx::pack_bf16x2(acc[4], acc[5]); o.w = ptx::pack_bf16x2(acc[6], acc[7]);
      ptx::st_v4(dst + v, o);
    }
  }
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    ptx::fence_acq_rel_sys();
    s_last = (atomicAdd(p.phase + 1, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (static_cast<int>(threadIdx.x) < W) {
      ptx::fence_acq_rel_sys();
      ptx::st_release_sys(symm_at(c, p.comb_flag + par * W + me, threadIdx.x), ph);
    }
    if (threadIdx.x == 0) p.phase[1] = 0;
  }
  // ---------------- token owner: sum of the rows returned by the (<= topk) ranks that hold its experts ----------------
  if (warp == 0) td::wait<true, true>(p.comb_flag + par * W, W, ph);
  __syncthreads();
  for (int t = blockIdx.x * wpc + warp; t < p.T; t += gridDim.x * wpc) {
    int e = -1;
    if (lane < p.topk) e = p.topk_idx[t * p.topk + lane];
    const bool valid = (e >= 0 && e < p.E);
    const int d = valid ? e / p.epr : -1 - lane;
    const unsigned same = __match_any_sync(0xffffffffu, d);
    const unsigned leaders = __ballot_sync(0xffffffffu, valid && lane == __ffs(same) - 1);
    for (int v = lane; v < vec_per_row; v += 32) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      unsigned rest = leaders;
      while (rest) {
        const int L = __ffs(rest) - 1;
        rest &= rest - 1;
        const int dL = __shfl_sync(0xffffffffu, d, L);
        const uint4 xv = ptx::ld_relaxed_sys_v4(comb_par + (static_cast<size_t>(dL) * p.T_max + t) * p.H * 2 + v * 16);
        const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[2 * q] += ptx::bf16_lo(xs[q]); acc[2 * q + 1] += ptx::bf16_hi(xs[q]); }
      }
      uint4 o;
      o.x = ptx::pack_bf16x2(acc[0], acc[1]); o.y = ptx::pack_bf16x2(acc[2], acc[3]);
      o.z = ptx::pack_bf16x2(acc[4], acc[5]); o.w = ptx::pack_bf16x2(acc[6], acc[7]);
      p.out[static_cast<size_t>(t) * vec_per_row + v] = o;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.phase + 2, 1u) == gridDim.x - 1) { p.phase[2] = 0; __threadfence(); p.phase[0] = ph; }
  }
}

struct TdSymmArgsN { long long rank, world; unsigned long long base, stride, mc_base; };
inline SymmCtx make_ctx(const TdSymmArgsN& s) {
  SymmCtx c; c.rank = (int)s.rank; c.world = (int)s.world; c.base = s.base; c.stride = s.stride; c.mc_base = s.mc_base;
  return c;
}

}  // namespace

struct TdEPNDispatchArgs {
  TdSymmArgsN symm;
  long long T, H, topk, E, T_max, P_max, grid;
  const void* x; const void* topk_idx; const void* topk_w;
  void* rx; long long rx_buf_bytes; void* rp; void* rmeta; void* rflag; void* send_rows; void* send_pairs; void* phase;
  void* pair_expert; void* pair_row; void* rcnt;
};

TD_API int td_ep_dispatch_normal(const TdEPNDispatchArgs* a, void* stream) {
  if (a->H % 8) { td::drv::set_error("ep_dispatch_normal: hidden size must be a multiple of 8"); return -1; }
  if (a->topk > 32) { td::drv::set_error("ep_dispatch_normal: topk <= 32"); return -1; }
  DispParams p;
  p.symm = make_ctx(a->symm);
  if (a->E % p.symm.world) { td::drv::set_error("ep_dispatch_normal: experts must divide evenly over ranks"); return -1; }
  if (a->T > a->T_max) { td::drv::set_error("ep_dispatch_normal: more tokens than the context was created for"); return -1; }
  p.T = (int)a->T; p.H = (int)a->H; p.topk = (int)a->topk; p.E = (int)a->E; p.epr = p.E / p.symm.world;
  p.T_max = (int)a->T_max; p.P_max = (int)a->P_max;
  p.x = (const uint4*)a->x; p.topk_idx = (const int*)a->topk_idx; p.topk_w = (const float*)a->topk_w;
  p.rx = (char*)a->rx; p.rx_buf_bytes = a->rx_buf_bytes; p.rp = (int4*)a->rp; p.rmeta = (int4*)a->rmeta;
  p.rflag = (unsigned long long*)a->rflag; p.send_rows = (int*)a->send_rows; p.send_pairs = (int*)a->send_pairs;
  p.phase = (uint32_t*)a->phase; p.pair_expert = (int*)a->pair_expert; p.pair_row = (int*)a->pair_row; p.rcnt = (int*)a->rcnt;
  ep_dispatch_normal_kernel<<<(int)a->grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

struct TdEPNCombineArgs {
  TdSymmArgsN symm;
  long long T, H, topk, E, T_max, P_max, grid;
  const void* y; const void* rp; const void* rmeta; const void* rcnt; const void* topk_idx;
  void* comb; long long comb_buf_bytes; void* comb_flag; void* phase; void* out;
};

TD_API int td_ep_combine_normal(const TdEPNCombineArgs* a, void* stream) {
  CombParams p;
  p.symm = make_ctx(a->symm);
  p.T = (int)a->T; p.H = (int)a->H; p.topk = (int)a->topk; p.E = (int)a->E; p.epr = p.E / p.symm.world;
  p.T_max = (int)a->T_max; p.P_max = (int)a->P_max;
  p.y = (const uint4*)a->y; p.rp = (const int4*)a->rp; p.rmeta = (const int4*)a->rmeta; p.rcnt = (const int*)a->rcnt;
  p.topk_idx = (const int*)a->topk_idx; p.comb = (char*)a->comb; p.comb_buf_bytes = a->comb_buf_bytes;
  p.comb_flag = (uint32_t*)a->comb_flag; p.phase = (uint32_t*)a->phase; p.out = (uint4*)a->out;
  ep_combine_normal_kernel<<<(int)a->grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
