// Expert-parallel low-latency dispatch / combine over NVLink (intra-node), bf16 or online-quantised fp8.
//
// Reference: kernels/nvidia/low_latency_all_to_all_v2.py (dispatch_kernel_v2 :156-356, combine_kernel_v2 :360-489,
// message = 16 B meta + payload + per-128 scales; 2-phase double buffers), layers/nvidia/ep_ll_a2a_layer.py, and the
// DeepEP low-latency semantics it follows (README.md:98-185).  Differences (B200-first):
//   * peers are addressed by pointer arithmetic on the fixed-stride symmetric heap; every message is written with
//     16-byte vector stores straight into the destination rank's staging slot  [local_expert][src_rank][slot];
//   * per-(expert, source) arrival is ONE 64-bit release store  (phase << 32 | count)  -- no signal reset, no
//     barrier; staging is double buffered by call parity, phase numbers are device resident (graph replayable);
//   * the receive side compacts arrivals into the packed per-expert layout the grouped GEMM consumes
//     (recv_x[le, start:start+count]) in the same kernel -- no separate post-process launch;
//   * NVLink needs many SMs to fill (measured ~6-10 GB/s per SM), so the grid is the whole chip and one warp moves
//     one token at a time.
#include <cuda_fp8.h>
#include "td/primitives.cuh"
#include "runtime/driver.h"

using namespace td;

namespace {

constexpr int kEPThreads = 256;
constexpr int kHdr = 16;   // bytes of per-message header (flat (token, k) index of the source)

struct EPParams {
  SymmCtx symm;
  int T, H, topk, E, epr, max_m;        // tokens on this rank, hidden, top-k, experts (global), experts per rank, slots per (expert, src)
  int msg_bytes;                        // kHdr + payload (+ scales), multiple of 16
  int use_fp8;
  const uint4* x;                       // [T, H] bf16
  const int* topk_idx;                  // [T, topk]
  char* staging;                        // symmetric: [2][epr][W][max_m][msg_bytes]
  long long staging_buf_bytes;
  unsigned long long* recv_flag;        // symmetric: [2][epr][W]   (phase << 32 | count)
  int* send_count;                      // local: [E]  (zero between calls; reset by the kernel)
  uint32_t* phase;                      // local: [0] completed calls, [1] CTA counter A, [2] exit counter
  // outputs (local)
  char* recv_x;                         // [epr][W * max_m][H] bf16 or fp8
  float* recv_scale;                    // [epr][W * max_m][H / 128] (fp8 only)
  int* recv_src_info;                   // [epr][W * max_m] flat (token * topk + k) index at the source rank
  long long* recv_range;                // [epr][W] (count, start) packed as (count << 32 | start)
  int* recv_count;                      // [epr]  (zeroed by the host before the launch)
};

TD_DEVICE float warp_max16(float v) {   // max over aligned groups of 16 lanes
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <bool kFP8>
__global__ void __launch_bounds__(kEPThreads, 1) ep_dispatch_kernel(const EPParams p) {
  const SymmCtx& c = p.symm;
  const int W = c.world, me = c.rank;
  const uint32_t ph = p.phase[0] + 1;
  const uint32_t par = ph & 1u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_cta = kEPThreads / 32;
  const int gw = blockIdx.x * warps_per_cta + warp, nw = gridDim.x * warps_per_cta;
  char* stage_local = p.staging + par * p.staging_buf_bytes;
  const int vec_per_row = p.H / 8;                       // 16-byte vectors of bf16 per token

  // ---------------- phase A: send ----------------
  for (int pair = gw; pair < p.T * p.topk; pair += nw) {
    const int e = p.topk_idx[pair];
    if (e < 0 || e >= p.E) continue;
    const int t = pair / p.topk;
    const int dst = e / p.epr, le = e % p.epr;
    int slot = 0;
    if (lane == 0) slot = atomicAdd(p.send_count + e, 1);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= p.max_m) continue;                       // capacity overflow: dropped (host asserts max_m >= T)
    char* msg = symm_at(c, stage_local, dst) + ((static_cast<size_t>(le) * W + me) * p.max_m + slot) * p.msg_bytes;
    if (lane == 0) ptx::st_v4(msg, make_uint4(static_cast<uint32_t>(pair), static_cast<uint32_t>(t), 0u, 0u));
    const uint4* row = p.x + static_cast<size_t>(t) * vec_per_row;
    if constexpr (!kFP8) {
      uint4* out = reinterpret_cast<uint4*>(msg + kHdr);
      for (int v = lane; v < vec_per_row; v += 32) ptx::st_v4(out + v, row[v]);
    } else {
      // 8 bf16 per lane; a 128-element quantisation group = 16 consecutive lanes
      uint2* out = reinterpret_cast<uint2*>(msg + kHdr);
      float* sc = reinterpret_cast<float*>(msg + kHdr + p.H);
      for (int v0 = 0; v0 < vec_per_row; v0 += 32) {
        const int v = v0 + lane;
        float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (v < vec_per_row) {
          const uint4 raw = row[v];
          const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) { f[2 * i] = ptx::bf16_lo(w[i]); f[2 * i + 1] = ptx::bf16_hi(w[i]); }
        }
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(f[i]));
        amax = warp_max16(amax);
        const float scale = fmaxf(amax, 1e-4f) * (1.f / 448.f);
        const float inv = 1.f / scale;
        if (v < vec_per_row) {
          __nv_fp8x4_e4m3 lo(make_float4(f[0] * inv, f[1] * inv, f[2] * inv, f[3] * inv));
          __nv_fp8x4_e4m3 hi(make_float4(f[4] * inv, f[5] * inv, f[6] * inv, f[7] * inv));
          out[v] = make_uint2(lo.__x, hi.__x);
          if ((lane & 15) == 0) sc[v / 16] = scale;
        }
      }
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
    for (int e = threadIdx.x; e < p.E; e += kEPThreads) {
      const int cnt = min(p.send_count[e], p.max_m);
      p.send_count[e] = 0;
      const int dst = e / p.epr, le = e % p.epr;
      unsigned long long* f = p.recv_flag + (static_cast<size_t>(par) * p.epr + le) * W + me;
      ptx::st_release_sys(reinterpret_cast<uint64_t*>(symm_at(c, f, dst)),
                          (static_cast<uint64_t>(ph) << 32) | static_cast<uint32_t>(cnt));
    }
    if (threadIdx.x == 0) p.phase[1] = 0;
  }

  // ---------------- phase B: receive + compact ----------------
  __shared__ int s_cnt, s_start;
  for (int unit = blockIdx.x; unit < p.epr * W; unit += gridDim.x) {
    const int le = unit / W, src = unit % W;
    if (threadIdx.x == 0) {
      const unsigned long long* f = p.recv_flag + (static_cast<size_t>(par) * p.epr + le) * W + src;
      uint64_t v;
      do { v = ptx::ld_acquire_sys(reinterpret_cast<const uint64_t*>(f)); } while ((v >> 32) != ph);
      const int cnt = static_cast<int>(v & 0xffffffffu);
      const int start = cnt ? atomicAdd(p.recv_count + le, cnt) : 0;
      p.recv_range[le * W + src] = (static_cast<long long>(cnt) << 32) | static_cast<uint32_t>(start);
      s_cnt = cnt; s_start = start;
    }
    __syncthreads();
    const int cnt = s_cnt, start = s_start;
    const char* src_msgs = stage_local + ((static_cast<size_t>(le) * W + src) * p.max_m) * p.msg_bytes;
    const size_t cap = static_cast<size_t>(W) * p.max_m;
    for (int i = warp; i < cnt; i += warps_per_cta) {
      const char* msg = src_msgs + static_cast<size_t>(i) * p.msg_bytes;
      const size_t row = static_cast<size_t>(le) * cap + start + i;
      if (lane == 0) p.recv_src_info[row] = static_cast<int>(ptx::ld_relaxed_sys_v4(msg).x);
      if constexpr (!kFP8) {
        uint4* dst = reinterpret_cast<uint4*>(p.recv_x + row * p.H * 2);
        for (int v = lane; v < vec_per_row; v += 32) dst[v] = ptx::ld_relaxed_sys_v4(msg + kHdr + v * 16);
      } else {
        uint4* dst = reinterpret_cast<uint4*>(p.recv_x + row * p.H);
        for (int v = lane; v < p.H / 16; v += 32) dst[v] = ptx::ld_relaxed_sys_v4(msg + kHdr + v * 16);
        const int ng = p.H / 128;
        for (int g = lane; g < ng; g += 32) p.recv_scale[row * ng + g] = *reinterpret_cast<const volatile float*>(msg + kHdr + p.H + g * 4);
      }
    }
    __syncthreads();
  }
  // ---------------- phase bookkeeping ----------------
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.phase + 2, 1u) == gridDim.x - 1) { p.phase[2] = 0; __threadfence(); p.phase[0] = ph; }
  }
}

struct EPCombineParams {
  SymmCtx symm;
  int T, H, topk, epr, max_m;
  const uint4* y;                 // [epr][W * max_m][H] bf16 expert outputs in the packed dispatch layout
  const int* src_info;            // [epr][W * max_m]
  const long long* recv_range;    // [epr][W]
  const int* topk_idx;            // [T, topk]
  const float* topk_w;            // [T, topk]
  char* comb;                     // symmetric: [2][T_max * topk][H] bf16
  long long comb_buf_bytes;
  uint32_t* comb_flag;            // symmetric: [2][W]
  uint32_t* phase;                // local: [0] calls, [1] counter, [2] exit
  uint4* out;                     // [T, H] bf16
};

__global__ void __launch_bounds__(kEPThreads, 1) ep_combine_kernel(const EPCombineParams p) {
  const SymmCtx& c = p.symm;
  const int W = c.world, me = c.rank;
  const uint32_t ph = p.phase[0] + 1;
  const uint32_t par = ph & 1u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_cta = kEPThreads / 32;
  const int vec_per_row = p.H / 8;
  char* comb_local = p.comb + par * p.comb_buf_bytes;
  const size_t cap = static_cast<size_t>(W) * p.max_m;

  // ---------------- phase A: return every expert output row to its source rank ----------------
  for (int unit = blockIdx.x; unit < p.epr * W; unit += gridDim.x) {
    const int le = unit / W, src = unit % W;
    const long long rg = p.recv_range[le * W + src];
    const int cnt = static_cast<int>(rg >> 32), start = static_cast<int>(rg & 0xffffffff);
    char* dst_base = symm_at(c, comb_local, src);
    for (int i = warp; i < cnt; i += warps_per_cta) {
      const size_t row = static_cast<size_t>(le) * cap + start + i;
      const int flat = p.src_info[row];
      const uint4* yr = p.y + row * vec_per_row;
      uint4* dst = reinterpret_cast<uint4*>(dst_base + static_cast<size_t>(flat) * p.H * 2);
      for (int v = lane; v < vec_per_row; v += 32) ptx::st_v4(dst + v, yr[v]);
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
    if (threadIdx.x < W) {
      ptx::fence_acq_rel_sys();
      ptx::st_release_sys(symm_at(c, p.comb_flag + par * W + me, threadIdx.x), ph);
    }
    if (threadIdx.x == 0) p.phase[1] = 0;
  }
  // ---------------- phase B: weighted top-k sum of what came back ----------------
  if (warp == 0) td::wait<true, true>(p.comb_flag + par * W, W, ph);
  __syncthreads();
  const long long total = static_cast<long long>(p.T) * vec_per_row;
  for (long long idx = blockIdx.x * (long long)kEPThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kEPThreads) {
    const int t = static_cast<int>(idx / vec_per_row), v = static_cast<int>(idx % vec_per_row);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < p.topk; ++k) {
      const int e = p.topk_idx[t * p.topk + k];
      if (e < 0 || e >= p.epr * W) continue;      // same validity predicate as dispatch: a pair that was never sent has no row to add
      const float w = p.topk_w[t * p.topk + k];
      const uint4 x = ptx::ld_relaxed_sys_v4(comb_local + (static_cast<size_t>(t) * p.topk + k) * p.H * 2 + v * 16);
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[2 * i] += w * ptx::bf16_lo(xs[i]); acc[2 * i + 1] += w * ptx::bf16_hi(xs[i]); }
    }
    uint4 o;
    o.x = ptx::pack_bf16x2(acc[0], acc[1]); o.y = ptx::pack_bf16x2(acc[2], acc[3]);
    o.z = ptx::pack_bf16x2(acc[4], acc[5]); o.w = ptx::pack_bf16x2(acc[6], acc[7]);
    p.out[idx] = o;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.phase + 2, 1u) == gridDim.x - 1) { p.phase[2] = 0; __threadfence(); p.phase[0] = ph; }
  }
}

struct TdSymmArgs { long long rank, world; unsigned long long base, stride, mc_base; };
inline SymmCtx make_ctx(const TdSymmArgs& s) {
  SymmCtx c; c.rank = (int)s.rank; c.world = (int)s.world; c.base = s.base; c.stride = s.stride; c.mc_base = s.mc_base;
  return c;
}

}  // namespace

struct TdEPDispatchArgs {
  TdSymmArgs symm;
  long long T, H, topk, E, max_m, use_fp8, grid;
  const void* x; const void* topk_idx; void* staging; long long staging_buf_bytes; void* recv_flag; void* send_count; void* phase;
  void* recv_x; void* recv_scale; void* recv_src_info; void* recv_range; void* recv_count;
};

TD_API long long td_ep_msg_bytes(long long H, int use_fp8) {
  long long payload = use_fp8 ? (H + (H / 128) * 4) : H * 2;
  return (kHdr + payload + 15) / 16 * 16;
}

TD_API int td_ep_dispatch(const TdEPDispatchArgs* a, void* stream) {
  if (a->H % 128) { td::drv::set_error("ep_dispatch: hidden size must be a multiple of 128"); return -1; }
  EPParams p;
  p.symm = make_ctx(a->symm);
  if (a->E % p.symm.world) { td::drv::set_error("ep_dispatch: experts must divide evenly over ranks"); return -1; }
  p.T = (int)a->T; p.H = (int)a->H; p.topk = (int)a->topk; p.E = (int)a->E; p.epr = p.E / p.symm.world; p.max_m = (int)a->max_m;
  p.use_fp8 = (int)a->use_fp8; p.msg_bytes = (int)td_ep_msg_bytes(a->H, p.use_fp8);
  p.x = (const uint4*)a->x; p.topk_idx = (const int*)a->topk_idx; p.staging = (char*)a->staging; p.staging_buf_bytes = a->staging_buf_bytes;
  p.recv_flag = (unsigned long long*)a->recv_flag; p.send_count = (int*)a->send_count; p.phase = (uint32_t*)a->phase;
  p.recv_x = (char*)a->recv_x; p.recv_scale = (float*)a->recv_scale; p.recv_src_info = (int*)a->recv_src_info;
  p.recv_range = (long long*)a->recv_range; p.recv_count = (int*)a->recv_count;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  TD_CUDA_CHECK(cudaMemsetAsync(p.recv_count, 0, sizeof(int) * p.epr, s));
  const int grid = (int)a->grid;
  if (p.use_fp8) ep_dispatch_kernel<true><<<grid, kEPThreads, 0, s>>>(p);
  else ep_dispatch_kernel<false><<<grid, kEPThreads, 0, s>>>(p);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

struct TdEPCombineArgs {
  TdSymmArgs symm;
  long long T, H, topk, E, max_m, grid;
  const void* y; const void* src_info; const void* recv_range; const void* topk_idx; const void* topk_w;
  void* comb; long long comb_buf_bytes; void* comb_flag; void* phase; void* out;
};

TD_API int td_ep_combine(const TdEPCombineArgs* a, void* stream) {
  EPCombineParams p;
  p.symm = make_ctx(a->symm);
  p.T = (int)a->T; p.H = (int)a->H; p.topk = (int)a->topk; p.epr = (int)(a->E / p.symm.world); p.max_m = (int)a->max_m;
  p.y = (const uint4*)a->y; p.src_info = (const int*)a->src_info; p.recv_range = (const long long*)a->recv_range;
  p.topk_idx = (const int*)a->topk_idx; p.topk_w = (const float*)a->topk_w;
  p.comb = (char*)a->comb; p.comb_buf_bytes = a->comb_buf_bytes; p.comb_flag = (uint32_t*)a->comb_flag; p.phase = (uint32_t*)a->phase;
  p.out = (uint4*)a->out;
  ep_combine_kernel<<<(int)a->grid, kEPThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
