// Device-side distributed primitives over the fixed-stride symmetric heap.
//
// These are the B200-native equivalents of the reference's MLIR `distributed` dialect ops and their
// NVIDIA lowering (spec of semantics: /root/reference/lib/Conversion/TritonDistributedToLLVM/NVIDIA/
// DistributedOpToLLVM.cpp:156-352 -- wait = per-lane `ld.acquire` spin + warp sync; notify = fence +
// scoped store / atomic add issued by one thread; symm_at = nvshmem_ptr), and of the NVSHMEM wrapper
// thunks (/root/reference/shmem/nvshmem_bind/runtime/nvshmem_wrapper.cu:28-695).  There is no dialect,
// no bitcode link step and no pointer table: every rank maps every peer's heap segment at
//     base + peer * stride
// in one contiguous VA reservation, so symm_at() is pure address arithmetic.
#pragma once
#include "ptx.cuh"

namespace td {

// Passed by value to every distributed kernel (fits in registers / constant bank).
struct SymmCtx {
  int rank;            // my rank in the symmetric team
  int world;           // team size
  uint64_t base;       // VA of rank 0's segment in *this* process
  uint64_t stride;     // bytes between consecutive ranks' segments
  uint64_t mc_base;    // VA of the NVLS multicast mapping of the heap (0 if unavailable)
};

enum class SignalOp : int { SET = 1, ADD = 2 };          // DistributedAttrDefs.td:36-44
enum class CommScope : int { GPU = 1, INTRA_NODE = 2 };  // DistributedAttrDefs.td:46-53 (INTER_NODE out of scope)

TD_DEVICE int rank(const SymmCtx& c) { return c.rank; }
TD_DEVICE int num_ranks(const SymmCtx& c) { return c.world; }

// Translate a pointer into my own segment to the same offset in `peer`'s segment.
template <typename T>
TD_DEVICE T* symm_at(const SymmCtx& c, T* local_ptr, int peer) {
  return reinterpret_cast<T*>(reinterpret_cast<uint64_t>(local_ptr) +
                              (static_cast<int64_t>(peer) - static_cast<int64_t>(c.rank)) * static_cast<int64_t>(c.stride));
}
// Translate a pointer into my own segment to the NVLS multicast alias of the same offset.
template <typename T>
TD_DEVICE T* symm_mc(const SymmCtx& c, T* local_ptr) {
  uint64_t off = reinterpret_cast<uint64_t>(local_ptr) - (c.base + static_cast<uint64_t>(c.rank) * c.stride);
  return reinterpret_cast<T*>(c.mc_base + off);
}

// ---- notify ------------------------------------------------------------------------------------
// One thread signals; callers make the CTA's prior stores happen-before it (e.g. __syncthreads()).
// The release store/red is cumulative over everything ordered before it by that barrier.
TD_DEVICE void notify(const SymmCtx& c, uint32_t* flag_local_addr, int peer, uint32_t value, SignalOp op = SignalOp::SET,
                      CommScope scope = CommScope::INTRA_NODE) {
  uint32_t* dst = (peer == c.rank) ? flag_local_addr : symm_at(c, flag_local_addr, peer);
  if (scope == CommScope::GPU && peer == c.rank) {
    if (op == SignalOp::SET) ptx::st_release_gpu(dst, value);
    else ptx::red_release_gpu_add(dst, value);
  } else {
    if (op == SignalOp::SET) ptx::st_release_sys(dst, value);
    else ptx::red_release_sys_add(dst, value);
  }
}

// ---- wait --------------------------------------------------------------------------------------
// Warp-cooperative: lane i < n spins on flags[i] until it equals (or exceeds, for monotone phase
// counters) `value`; then the warp reconverges.  Returns a token to thread through consume_token().
template <bool kGreaterEqual = false, bool kSysScope = true>
TD_DEVICE uint32_t wait(const uint32_t* flags, int n, uint32_t value) {
  const uint32_t lane = ptx::lane_id();
  if (static_cast<int>(lane) < n) {
    uint32_t v; TD_SPIN_GUARD(guard)
    do {
      v = kSysScope ? ptx::ld_acquire_sys(flags + lane) : ptx::ld_acquire_gpu(flags + lane); TD_SPIN_POLL(guard, "td::wait", flags + lane, v, value)
    } while (kGreaterEqual ? (static_cast<int32_t>(v - value) < 0) : (v != value));
  }
  __syncwarp();
  return value;
}
// Single-thread wait for one flag (used by TMA-issuing elected threads: the acquiring thread is the
// one that subsequently issues the async-proxy loads, followed by a proxy fence).
template <bool kSysScope = true>
TD_DEVICE void wait_ge(const uint32_t* flag, uint32_t value) {
  uint32_t v; TD_SPIN_GUARD(guard)
  do {
    v = kSysScope ? ptx::ld_acquire_sys(flag) : ptx::ld_acquire_gpu(flag); TD_SPIN_POLL(guard, "td::wait_ge", flag, v, value)
  } while (static_cast<int32_t>(v - value) < 0);
}
// consume_token: identity in hand-written CUDA -- kept so device code reads like the reference's kernels.
template <typename T>
TD_DEVICE T consume_token(T v, uint32_t /*token*/) { return v; }

// ---- cross-GPU barrier (whole CTA participates; flag-flip, no atomics) ---------------------------
// slots: uint32 [2][world] in the symmetric heap, zero-initialised; `epoch` is a monotonically
// increasing per-barrier counter kept by the caller (device memory, so CUDA graphs can replay it).
// Round-parity selects one of two slot arrays so back-to-back barriers cannot alias
// (same idea as /root/reference/python/triton_dist/kernels/nvidia/common_ops.py:172-224).
TD_DEVICE void barrier_all_block(const SymmCtx& c, uint32_t* slots, uint32_t epoch) {
  __syncthreads();
  uint32_t* arr = slots + (epoch & 1u) * c.world;
  const int t = threadIdx.x;
  if (t < c.world) {
    ptx::fence_acq_rel_sys();
    ptx::st_release_sys(symm_at(c, arr + c.rank, t), epoch);   // write my arrival into peer t's slot[me]
    uint32_t v;
    TD_SPIN_GUARD(guard) do { v = ptx::ld_acquire_sys(arr + t); TD_SPIN_POLL(guard, "td::barrier_all_block (peer = thread)", arr + t, v, epoch) } while (static_cast<int32_t>(v - epoch) < 0);  // peer t arrived at me
  }
  __syncthreads();
}

// ---- grid barrier (all CTAs co-resident) ---------------------------------------------------------
// counter: one uint32 in global memory, zero-initialised, monotone.  `gen` = number of barriers this
// launch has already executed on this counter + base generation (caller-tracked).
TD_DEVICE void grid_barrier(uint32_t* counter, uint32_t target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::red_release_gpu_add(counter, 1u);
    TD_SPIN_GUARD(guard) while (static_cast<int32_t>(ptx::ld_acquire_gpu(counter) - target) < 0) { TD_SPIN_POLL(guard, "td::grid_barrier", counter, 0ull, target)
    }
  }
  __syncthreads();
}

// 16 independent 16-byte loads per thread before the first store: the copy loop of a comm CTA is bound by the latency of
// its (L2) reads -- 4 loads in flight per thread moved 23 GB/s per SM over NVLink (intra-kernel profile, 8xB200), the
// port sustains ~45 GB/s per SM.
TD_DEVICE void copy16_strided_deep(void* dst, const void* src, size_t bytes, int tid, int nthreads) {
  const size_t n = bytes >> 4;
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  constexpr int U = 16;
  size_t i = tid;
  for (; i + (U - 1) * static_cast<size_t>(nthreads) < n; i += U * static_cast<size_t>(nthreads)) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ptx::ld_nc_v4(s + i + u * static_cast<size_t>(nthreads));
#pragma unroll
    for (int u = 0; u < U; ++u) ptx::st_na_v4(d + i + u * static_cast<size_t>(nthreads), v[u]);
  }
  for (; i < n; i += nthreads) ptx::st_na_v4(d + i, ptx::ld_nc_v4(s + i));
}

// ---- put / get (block / warp / thread scopes) over peer pointers ----------------------------------
// 16-byte vectorised; src/dst/bytes must be 16 B aligned.  Equivalent of nvshmem putmem/getmem
// (nvshmem_wrapper.cu: putmem_block / getmem_block ...), intra-node only.
TD_DEVICE void copy16_strided(void* dst, const void* src, size_t bytes, int tid, int nthreads) {
  const size_t n = bytes >> 4;
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  size_t i = tid;
  // 4-deep unroll: all loads first, then all stores (memory-level parallelism)
  for (; i + 3 * (size_t)nthreads < n; i += 4 * (size_t)nthreads) {
    uint4 v0 = ptx::ld_nc_v4(s + i);
    uint4 v1 = ptx::ld_nc_v4(s + i + nthreads);
    uint4 v2 = ptx::ld_nc_v4(s + i + 2 * (size_t)nthreads);
    uint4 v3 = ptx::ld_nc_v4(s + i + 3 * (size_t)nthreads);
    ptx::st_na_v4(d + i, v0);
    ptx::st_na_v4(d + i + nthreads, v1);
    ptx::st_na_v4(d + i + 2 * (size_t)nthreads, v2);
    ptx::st_na_v4(d + i + 3 * (size_t)nthreads, v3);
  }
  for (; i < n; i += nthreads) ptx::st_na_v4(d + i, ptx::ld_nc_v4(s + i));
}
TD_DEVICE void putmem_block(const SymmCtx& c, void* dst_local_addr, const void* src, size_t bytes, int peer) {
  copy16_strided(symm_at(c, reinterpret_cast<char*>(dst_local_addr), peer), src, bytes, threadIdx.x, blockDim.x);
}
TD_DEVICE void getmem_block(const SymmCtx& c, void* dst, const void* src_local_addr, size_t bytes, int peer) {
  copy16_strided(dst, symm_at(c, reinterpret_cast<const char*>(src_local_addr), peer), bytes, threadIdx.x, blockDim.x);
}
TD_DEVICE void putmem_warp(const SymmCtx& c, void* dst_local_addr, const void* src, size_t bytes, int peer) {
  copy16_strided(symm_at(c, reinterpret_cast<char*>(dst_local_addr), peer), src, bytes, ptx::lane_id(), 32);
}
TD_DEVICE void getmem_warp(const SymmCtx& c, void* dst, const void* src_local_addr, size_t bytes, int peer) {
  copy16_strided(dst, symm_at(c, reinterpret_cast<const char*>(src_local_addr), peer), bytes, ptx::lane_id(), 32);
}
// put + signal: data first, then a release-scoped flag write on the same peer (nvshmem putmem_signal).
TD_DEVICE void putmem_signal_block(const SymmCtx& c, void* dst_local_addr, const void* src, size_t bytes,
                                   uint32_t* sig_local_addr, uint32_t sig_val, SignalOp op, int peer) {
  putmem_block(c, dst_local_addr, src, bytes, peer);
  __syncthreads();
  if (threadIdx.x == 0) notify(c, sig_local_addr, peer, sig_val, op);
}
TD_DEVICE void putmem_signal_warp(const SymmCtx& c, void* dst_local_addr, const void* src, size_t bytes,
                                  uint32_t* sig_local_addr, uint32_t sig_val, SignalOp op, int peer) {
  putmem_warp(c, dst_local_addr, src, bytes, peer);
  __syncwarp();
  if (ptx::lane_id() == 0) notify(c, sig_local_addr, peer, sig_val, op);
}
TD_DEVICE void signal_wait_until_ge(const uint32_t* sig, uint32_t value) { wait_ge<true>(sig, value); }

}  // namespace td
