// Low-latency ("LL", flag-in-data) all-reduce for tiny messages -- decode-step activations (bs x hidden x 2 B = 8 KB).
//
// NCCL's LL idea, as the reference uses it for its low-latency all-gather (kernels/nvidia/low_latency_allgather.py:531-567
// _pack_ll_block / _recv_ll_block): every 8-byte atom carries 4 bytes of payload and the 4-byte phase number of the call, written
// with ONE 8-byte store, so "the flag is set" and "the data is there" are the same event: no fence, no separate flag, no barrier.
// Each rank writes its packed message into slot [me] of every peer's buffer; every rank then spins on the atoms of all W slots and
// sums them in fp32.  Buffers are parity double-buffered and owned exclusively by this kernel (a stale atom can never carry the
// current phase), the phase counter lives on the device (CUDA-graph replayable).
//
// STATUS: written after this round's GPU budget was spent (opt-in via AllReduceMethod.OneShot_LL); the send / receive loops are the
// hardware-validated loops of allgather_kernel<2> (csrc/comm_kernels.cu) with the copy-out replaced by the sum.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "runtime/driver.h"
#include "td/primitives.cuh"
#include "td/ptx.cuh"

using namespace td;

namespace {

constexpr int kLLThreads = 256;

struct LLParams {
  SymmCtx symm;
  const uint32_t* in;      // nwords 32-bit words (2 x 16-bit or 1 x fp32 each)
  uint32_t* out;
  uint2* buf;              // symmetric [2][W][nwords] atoms {payload, phase}
  long long buf_atoms;     // atoms per parity half (= W * max_words)
  long long max_words;     // slot stride
  long long nwords;
  uint32_t* phase;         // local [0] completed calls, [1] exit counter
  int dtype;               // 0 bf16, 1 fp16, 2 fp32 (same codes as csrc/comm_kernels.cu)
};

__global__ void __launch_bounds__(kLLThreads, 1) allreduce_ll_kernel(const LLParams p) {
  const SymmCtx& c = p.symm;
  const int W = c.world;
  const uint32_t ph = p.phase[0] + 1;
  uint2* buf = p.buf + (ph & 1u) * p.buf_atoms;
  const long long stride = static_cast<long long>(gridDim.x) * kLLThreads;
  // send: my words, tagged with the phase, into slot [me] of every rank (mine included: one code path)
  for (long long v = blockIdx.x * static_cast<long long>(kLLThreads) + threadIdx.x; v < p.nwords; v += stride) {
    const uint32_t w = p.in[v];
    for (int r = 0; r < W; ++r) {
      uint2* dst = symm_at(c, buf + c.rank * p.max_words + v, (c.rank + r) % W);
      asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(w), "r"(ph) : "memory");
    }
  }
  // receive + reduce: an atom is valid once its flag half equals this call's phase
  for (long long v = blockIdx.x * static_cast<long long>(kLLThreads) + threadIdx.x; v < p.nwords; v += stride) {
    float a0 = 0.f, a1 = 0.f;
    for (int s = 0; s < W; ++s) {
      const uint2* src = buf + s * p.max_words + v;
      uint32_t d, f;
      do {
        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(d), "=r"(f) : "l"(src) : "memory");
      } while (f != ph);
      if (p.dtype == 0) { a0 += ptx::bf16_lo(d); a1 += ptx::bf16_hi(d); }
      else if (p.dtype == 1) { const __half2 h = *reinterpret_cast<const __half2*>(&d); a0 += __low2float(h); a1 += __high2float(h); }
      else a0 += __uint_as_float(d);
    }
    p.out[v] = p.dtype == 0 ? ptx::pack_bf16x2(a0, a1) : p.dtype == 1 ? ptx::pack_f16x2(a0, a1) : __float_as_uint(a0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.phase + 1, 1u) == gridDim.x - 1) { p.phase[1] = 0; __threadfence(); p.phase[0] = ph; }
  }
}

}  // namespace

struct TdARLLArgs {
  long long rank, world; unsigned long long base, stride, mc_base;
  const void* in; void* out; void* buf; long long max_words; long long nbytes; void* phase; long long dtype; long long grid;
};

TD_API int td_allreduce_ll(const TdARLLArgs* a, void* stream) {
  if (a->nbytes % 4) { td::drv::set_error("allreduce_ll: nbytes must be a multiple of 4"); return -1; }
  const long long nwords = a->nbytes / 4;
  if (nwords > a->max_words) { td::drv::set_error("allreduce_ll: message larger than the LL buffer"); return -1; }
  LLParams p;
  p.symm.rank = (int)a->rank; p.symm.world = (int)a->world; p.symm.base = a->base; p.symm.stride = a->stride; p.symm.mc_base = a->mc_base;
  p.in = reinterpret_cast<const uint32_t*>(a->in); p.out = reinterpret_cast<uint32_t*>(a->out);
  p.buf = reinterpret_cast<uint2*>(a->buf); p.max_words = a->max_words; p.buf_atoms = a->max_words * a->world;
  p.nwords = nwords; p.phase = reinterpret_cast<uint32_t*>(a->phase); p.dtype = (int)a->dtype;
  int grid = (int)a->grid;
  if (grid <= 0) grid = (int)((nwords + kLLThreads - 1) / kLLThreads);
  if (grid > 32) grid = 32;
  if (grid < 1) grid = 1;
  allreduce_ll_kernel<<<grid, kLLThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  TD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
