// NVSHMEM-style device API over the symmetric heap of this framework (no NVSHMEM library).
//
// The reference exposes ~80 `libshmem_device.*` functions to its kernels (python/triton_dist/language/extra/cuda/
// libnvshmem_device.py:102-990: my_pe / n_pes / team_* / remote_ptr / barrier* / sync* / quiet / fence / getmem* / putmem* /
// putmem_signal* / signal_op / signal_wait_until / broadcast* / fcollect* / putmem_rma*), thunked through
// shmem/nvshmem_bind/runtime/nvshmem_wrapper.cu.  On one NVSwitch domain every peer's heap segment is mapped into every
// GPU's address space at `base + pe * stride`, so all of them reduce to address arithmetic + ordinary (scoped) loads and
// stores: a put IS a store, "nbi" and blocking variants coincide (completion = visibility after quiet / a release),
// the *_rma variants (IBGDA in the reference) are the same NVLink stores.
//
// Scopes follow NVSHMEM: no suffix = calling thread, _warp = all 32 lanes call with identical arguments, _block = all
// threads of the CTA call with identical arguments.  Collectives (barrier / sync / broadcast / fcollect) take a `Sync`
// object: `slots` = uint32 [2][world] on the symmetric heap (zeroed once), `epoch` = a local device counter the
// collective itself advances (CUDA-graph replayable: no host-side state).
#pragma once
#include "primitives.cuh"

namespace td {
namespace shmem {

// ---- constants (values as in NVSHMEM so ported kernels keep their literals) ----------------------------------------
enum Cmp : int { CMP_EQ = 0, CMP_NE = 1, CMP_GT = 2, CMP_LE = 3, CMP_LT = 4, CMP_GE = 5 };
enum SigOp : int { SIGNAL_SET = 9, SIGNAL_ADD = 10 };

// A team is an arithmetic progression of PEs (what nvshmem_team_split_strided produces).
struct Team {
  int start, stride, size;
};
struct Sync {
  uint32_t* slots;   // symmetric: uint32 [2][world]
  uint32_t* epoch;   // local: one uint32 (device memory), advanced by each collective
};

TD_DEVICE int my_pe(const SymmCtx& c) { return c.rank; }
TD_DEVICE int n_pes(const SymmCtx& c) { return c.world; }
TD_DEVICE Team team_world(const SymmCtx& c) { return Team{0, 1, c.world}; }
TD_DEVICE int team_n_pes(const Team& t) { return t.size; }
// index of PE `pe` inside the team, -1 when it is not a member
TD_DEVICE int team_index_of(const Team& t, int pe) {
  const int d = pe - t.start;
  if (d < 0 || t.stride <= 0 || d % t.stride != 0) return -1;
  const int i = d / t.stride;
  return i < t.size ? i : -1;
}
TD_DEVICE int team_my_pe(const SymmCtx& c, const Team& t) { return team_index_of(t, c.rank); }
TD_DEVICE int team_pe(const Team& t, int idx) { return t.start + idx * t.stride; }
// nvshmem_team_translate_pe: index `src_pe` of src_team -> index in dest_team (-1 when absent)
TD_DEVICE int team_translate_pe(const Team& src_team, int src_pe, const Team& dest_team) {
  if (src_pe < 0 || src_pe >= src_team.size) return -1;
  return team_index_of(dest_team, team_pe(src_team, src_pe));
}

template <typename T>
TD_DEVICE T* remote_ptr(const SymmCtx& c, T* local_ptr, int pe) { return symm_at(c, local_ptr, pe); }
template <typename T>
TD_DEVICE T* remote_mc_ptr(const SymmCtx& c, T* local_ptr) { return c.mc_base ? symm_mc(c, local_ptr) : nullptr; }

// ---- ordering -------------------------------------------------------------------------------------------------------
// fence: order my prior puts before my later puts (per destination); quiet: complete all my prior puts.  Both are the
// system-scope acq_rel fence here -- NVLink stores are ordinary memory operations of the issuing thread.
TD_DEVICE void fence() { ptx::fence_acq_rel_sys(); }
TD_DEVICE void quiet() { ptx::fence_acq_rel_sys(); }

// ---- point-to-point -------------------------------------------------------------------------------------------------
TD_DEVICE void int_p(const SymmCtx& c, int* dst_local_addr, int value, int pe) { *symm_at(c, dst_local_addr, pe) = value; }

// calling thread moves all bytes (16-byte chunks when everything is aligned, bytes otherwise)
TD_DEVICE void copy_thread(void* dst, const void* src, size_t bytes) {
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15u) == 0) {
    copy16_strided(dst, src, bytes, 0, 1);
  } else {
    char* d = reinterpret_cast<char*>(dst);
    const char* s = reinterpret_cast<const char*>(src);
    for (size_t i = 0; i < bytes; ++i) d[i] = s[i];
  }
}
// group-cooperative byte copy with an unaligned tail (the 16-byte fast path of primitives.cuh needs aligned operands)
TD_DEVICE void copy_group(void* dst, const void* src, size_t bytes, int tid, int nthreads) {
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0) {
    const size_t body = bytes & ~size_t(15);
    copy16_strided(dst, src, body, tid, nthreads);
    char* d = reinterpret_cast<char*>(dst);
    const char* s = reinterpret_cast<const char*>(src);
    for (size_t i = body + tid; i < bytes; i += nthreads) d[i] = s[i];
  } else {
    char* d = reinterpret_cast<char*>(dst);
    const char* s = reinterpret_cast<const char*>(src);
    for (size_t i = tid; i < bytes; i += nthreads) d[i] = s[i];
  }
}

TD_DEVICE void putmem(const SymmCtx& c, void* dst_local_addr, const void* src, size_t bytes, int pe) {
  copy_thread(symm_at(c, reinterpret_cast<char*>(dst_local_addr), pe), src, bytes);
}
TD_DEVICE void getmem(const SymmCtx& c, void* dst, const void* src_local_addr, size_t bytes, int pe) {
  copy_thread(dst, symm_at(c, reinterpret_cast<const char*>(src_local_addr), pe), bytes);
}
TD_DEVICE void putmem_warp(const SymmCtx& c, void* dst_local_addr, const void* src, size_t bytes, int pe) {
  copy_group(symm_at(c, reinterpret_cast<char*>(dst_local_addr), pe), src, bytes, ptx::lane_id(), 32);
  __syncwarp();
}
TD_DEVICE void getmem_warp(const SymmCtx& c, void* dst, const void* src_local_addr, size_t bytes, int pe) {
  copy_group(dst, symm_at(c, reinterpret_cast<const char*>(src_local_addr), pe), bytes, ptx::lane_id(), 32);
  __syncwarp();
}
TD_DEVICE void putmem_block(const SymmCtx& c, void* dst_local_addr, const void* src, size_t bytes, int pe) {
  copy_group(symm_at(c, reinterpret_cast<char*>(dst_local_addr), pe), src, bytes, threadIdx.x, blockDim.x);
  __syncthreads();
}
TD_DEVICE void getmem_block(const SymmCtx& c, void* dst, const void* src_local_addr, size_t bytes, int pe) {
  copy_group(dst, symm_at(c, reinterpret_cast<const char*>(src_local_addr), pe), bytes, threadIdx.x, blockDim.x);
  __syncthreads();
}
// non-blocking-interface and RMA spellings: same operation (see the header comment)
#define TD_SHMEM_ALIAS(name, target)                                                                      \
  TD_DEVICE void name(const SymmCtx& c, void* d, const void* s, size_t b, int pe) { shmem::target(c, d, s, b, pe); }
TD_SHMEM_ALIAS(putmem_nbi, putmem)
TD_SHMEM_ALIAS(getmem_nbi, getmem)
TD_SHMEM_ALIAS(putmem_nbi_warp, putmem_warp)
TD_SHMEM_ALIAS(getmem_nbi_warp, getmem_warp)
TD_SHMEM_ALIAS(putmem_nbi_block, putmem_block)
TD_SHMEM_ALIAS(getmem_nbi_block, getmem_block)
TD_SHMEM_ALIAS(putmem_rma, putmem)
TD_SHMEM_ALIAS(putmem_rma_warp, putmem_warp)
TD_SHMEM_ALIAS(putmem_rma_block, putmem_block)
TD_SHMEM_ALIAS(putmem_rma_nbi, putmem)
TD_SHMEM_ALIAS(putmem_rma_nbi_warp, putmem_warp)
TD_SHMEM_ALIAS(putmem_rma_nbi_block, putmem_block)
#undef TD_SHMEM_ALIAS

// ---- signals ----------------------------------------------------------------------------------------------------------
// 64-bit signal words as in NVSHMEM (the 32-bit flags of primitives.cuh are this framework's native form)
TD_DEVICE void signal_op(const SymmCtx& c, uint64_t* sig_local_addr, uint64_t value, int op, int pe) {
  uint64_t* dst = symm_at(c, sig_local_addr, pe);
  if (op == SIGNAL_ADD) {
    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(dst), "l"(value) : "memory");
  } else {
    ptx::st_release_sys(dst, value);
  }
}
TD_DEVICE bool cmp_holds(int cmp, uint64_t v, uint64_t ref) {
  switch (cmp) {
    case CMP_EQ: return v == ref;
    case CMP_NE: return v != ref;
    case CMP_GT: return v > ref;
    case CMP_LE: return v <= ref;
    case CMP_LT: return v < ref;
    default: return v >= ref;
  }
}
// spins (acquire, system scope) until `*sig cmp value` holds; returns the value that satisfied the condition
TD_DEVICE uint64_t signal_wait_until(const uint64_t* sig, int cmp, uint64_t value) {
  uint64_t v;
  TD_SPIN_GUARD(guard)
  do {
    v = ptx::ld_acquire_sys(sig);
    TD_SPIN_POLL(guard, "shmem::signal_wait_until", sig, v, value)
  } while (!cmp_holds(cmp, v, value));
  return v;
}

// put + signal: data first, then the signal with release semantics (the receiver's acquire of the signal makes the data
// visible).  Thread scope: the calling thread does both; warp / block: everyone copies, one thread signals.
TD_DEVICE void putmem_signal(const SymmCtx& c, void* dst_local_addr, const void* src, size_t bytes, uint64_t* sig_local_addr,
                             uint64_t sig_val, int sig_op, int pe) {
  shmem::putmem(c, dst_local_addr, src, bytes, pe);
  signal_op(c, sig_local_addr, sig_val, sig_op, pe);
}
TD_DEVICE void putmem_signal_warp(const SymmCtx& c, void* dst_local_addr, const void* src, size_t bytes, uint64_t* sig_local_addr,
                                  uint64_t sig_val, int sig_op, int pe) {
  shmem::putmem_warp(c, dst_local_addr, src, bytes, pe);
  if (ptx::lane_id() == 0) signal_op(c, sig_local_addr, sig_val, sig_op, pe);
  __syncwarp();
}
TD_DEVICE void putmem_signal_block(const SymmCtx& c, void* dst_local_addr, const void* src, size_t bytes, uint64_t* sig_local_addr,
                                   uint64_t sig_val, int sig_op, int pe) {
  shmem::putmem_block(c, dst_local_addr, src, bytes, pe);
  if (threadIdx.x == 0) signal_op(c, sig_local_addr, sig_val, sig_op, pe);
  __syncthreads();
}
#define TD_SHMEM_ALIAS_SIG(name, target)                                                                               \
  TD_DEVICE void name(const SymmCtx& c, void* d, const void* s, size_t b, uint64_t* sig, uint64_t v, int op, int pe) { \
    shmem::target(c, d, s, b, sig, v, op, pe);                                                                                \
  }
TD_SHMEM_ALIAS_SIG(putmem_signal_nbi, putmem_signal)
TD_SHMEM_ALIAS_SIG(putmem_signal_nbi_warp, putmem_signal_warp)
TD_SHMEM_ALIAS_SIG(putmem_signal_nbi_block, putmem_signal_block)
TD_SHMEM_ALIAS_SIG(putmem_signal_rma, putmem_signal)
TD_SHMEM_ALIAS_SIG(putmem_signal_rma_warp, putmem_signal_warp)
TD_SHMEM_ALIAS_SIG(putmem_signal_rma_block, putmem_signal_block)
TD_SHMEM_ALIAS_SIG(putmem_signal_rma_nbi, putmem_signal)
TD_SHMEM_ALIAS_SIG(putmem_signal_rma_nbi_warp, putmem_signal_warp)
TD_SHMEM_ALIAS_SIG(putmem_signal_rma_nbi_block, putmem_signal_block)
#undef TD_SHMEM_ALIAS_SIG

// ---- barriers ---------------------------------------------------------------------------------------------------------
// One thread per team member exchanges arrival flags (flag-flip on the epoch parity, as barrier_all_block of primitives.cuh).
// kGroup: 0 = thread (the caller alone loops over the members), 1 = warp, 2 = block.
template <int kGroup>
TD_DEVICE void team_sync_impl(const SymmCtx& c, const Team& t, const Sync& s) {
  if (kGroup == 2) __syncthreads();
  if (kGroup == 1) __syncwarp();
  const int tid = kGroup == 2 ? static_cast<int>(threadIdx.x) : (kGroup == 1 ? static_cast<int>(ptx::lane_id()) : 0);
  const int nthr = kGroup == 2 ? static_cast<int>(blockDim.x) : (kGroup == 1 ? 32 : 1);
  const uint32_t epoch = *s.epoch + 1u;
  uint32_t* arr = s.slots + (epoch & 1u) * c.world;
  if (team_index_of(t, c.rank) >= 0) {
    ptx::fence_acq_rel_sys();
    for (int i = tid; i < t.size; i += nthr) ptx::st_release_sys(symm_at(c, arr + c.rank, team_pe(t, i)), epoch);
    for (int i = tid; i < t.size; i += nthr) {
      uint32_t v;
      TD_SPIN_GUARD(guard)
      do {
        v = ptx::ld_acquire_sys(arr + team_pe(t, i));
        TD_SPIN_POLL(guard, "shmem team barrier (slot of a member)", arr + team_pe(t, i), v, epoch)
      } while (static_cast<int32_t>(v - epoch) < 0);
    }
  }
  if (kGroup == 2) __syncthreads();
  if (kGroup == 1) __syncwarp();
  if (tid == 0) *s.epoch = epoch;
  if (kGroup == 2) __syncthreads();
  if (kGroup == 1) __syncwarp();
}
// sync = arrival exchange; barrier = quiet + sync (all prior puts of the caller are complete and visible afterwards).
// team_sync_impl already fences before signalling, so both names share one implementation.
TD_DEVICE void team_sync(const SymmCtx& c, const Team& t, const Sync& s) { team_sync_impl<0>(c, t, s); }
TD_DEVICE void team_sync_warp(const SymmCtx& c, const Team& t, const Sync& s) { team_sync_impl<1>(c, t, s); }
TD_DEVICE void team_sync_block(const SymmCtx& c, const Team& t, const Sync& s) { team_sync_impl<2>(c, t, s); }
TD_DEVICE void barrier(const SymmCtx& c, const Team& t, const Sync& s) { team_sync_impl<0>(c, t, s); }
TD_DEVICE void barrier_warp(const SymmCtx& c, const Team& t, const Sync& s) { team_sync_impl<1>(c, t, s); }
TD_DEVICE void barrier_block(const SymmCtx& c, const Team& t, const Sync& s) { team_sync_impl<2>(c, t, s); }
TD_DEVICE void sync_all(const SymmCtx& c, const Sync& s) { team_sync_impl<0>(c, team_world(c), s); }
TD_DEVICE void sync_all_warp(const SymmCtx& c, const Sync& s) { team_sync_impl<1>(c, team_world(c), s); }
TD_DEVICE void sync_all_block(const SymmCtx& c, const Sync& s) { team_sync_impl<2>(c, team_world(c), s); }
TD_DEVICE void barrier_all(const SymmCtx& c, const Sync& s) { team_sync_impl<0>(c, team_world(c), s); }
TD_DEVICE void barrier_all_warp(const SymmCtx& c, const Sync& s) { team_sync_impl<1>(c, team_world(c), s); }
TD_DEVICE void barrier_all_block(const SymmCtx& c, const Sync& s) { team_sync_impl<2>(c, team_world(c), s); }

// ---- collectives ------------------------------------------------------------------------------------------------------
// broadcast: `nbytes` of the root's `src` arrive in `dst` (symmetric) of every team member (root included).
// The root pushes (NVLink stores), then the team synchronises.  root = index inside the team.
template <int kGroup>
TD_DEVICE void broadcastmem_impl(const SymmCtx& c, const Team& t, const Sync& s, void* dst_local_addr, const void* src, size_t nbytes,
                                 int root) {
  const int tid = kGroup == 2 ? static_cast<int>(threadIdx.x) : (kGroup == 1 ? static_cast<int>(ptx::lane_id()) : 0);
  const int nthr = kGroup == 2 ? static_cast<int>(blockDim.x) : (kGroup == 1 ? 32 : 1);
  if (team_my_pe(c, t) == root) {
    for (int i = 0; i < t.size; ++i)
      copy_group(symm_at(c, reinterpret_cast<char*>(dst_local_addr), team_pe(t, i)), src, nbytes, tid, nthr);
  }
  team_sync_impl<kGroup>(c, t, s);
}
TD_DEVICE void broadcastmem(const SymmCtx& c, const Team& t, const Sync& s, void* d, const void* src, size_t n, int root) {
  broadcastmem_impl<0>(c, t, s, d, src, n, root);
}
TD_DEVICE void broadcastmem_warp(const SymmCtx& c, const Team& t, const Sync& s, void* d, const void* src, size_t n, int root) {
  broadcastmem_impl<1>(c, t, s, d, src, n, root);
}
TD_DEVICE void broadcastmem_block(const SymmCtx& c, const Team& t, const Sync& s, void* d, const void* src, size_t n, int root) {
  broadcastmem_impl<2>(c, t, s, d, src, n, root);
}
// typed spellings (nvshmem_<type>_broadcast): element counts instead of bytes
template <typename T>
TD_DEVICE void broadcast(const SymmCtx& c, const Team& t, const Sync& s, T* dst, const T* src, size_t nelems, int root) {
  broadcastmem_impl<0>(c, t, s, dst, src, nelems * sizeof(T), root);
}
template <typename T>
TD_DEVICE void broadcast_warp(const SymmCtx& c, const Team& t, const Sync& s, T* dst, const T* src, size_t nelems, int root) {
  broadcastmem_impl<1>(c, t, s, dst, src, nelems * sizeof(T), root);
}
template <typename T>
TD_DEVICE void broadcast_block(const SymmCtx& c, const Team& t, const Sync& s, T* dst, const T* src, size_t nelems, int root) {
  broadcastmem_impl<2>(c, t, s, dst, src, nelems * sizeof(T), root);
}

// fcollect (all-gather of equal contributions): member i's `src` lands at dst + i * nbytes on every member.
template <int kGroup>
TD_DEVICE void fcollectmem_impl(const SymmCtx& c, const Team& t, const Sync& s, void* dst_local_addr, const void* src, size_t nbytes) {
  const int tid = kGroup == 2 ? static_cast<int>(threadIdx.x) : (kGroup == 1 ? static_cast<int>(ptx::lane_id()) : 0);
  const int nthr = kGroup == 2 ? static_cast<int>(blockDim.x) : (kGroup == 1 ? 32 : 1);
  const int me = team_my_pe(c, t);
  if (me >= 0) {
    for (int q = 0; q < t.size; ++q) {
      const int i = (me + q) % t.size;          // every member starts with a different destination
      copy_group(symm_at(c, reinterpret_cast<char*>(dst_local_addr), team_pe(t, i)) + static_cast<size_t>(me) * nbytes, src, nbytes, tid, nthr);
    }
  }
  team_sync_impl<kGroup>(c, t, s);
}
template <typename T>
TD_DEVICE void fcollect(const SymmCtx& c, const Team& t, const Sync& s, T* dst, const T* src, size_t nelems) {
  fcollectmem_impl<0>(c, t, s, dst, src, nelems * sizeof(T));
}
template <typename T>
TD_DEVICE void fcollect_warp(const SymmCtx& c, const Team& t, const Sync& s, T* dst, const T* src, size_t nelems) {
  fcollectmem_impl<1>(c, t, s, dst, src, nelems * sizeof(T));
}
template <typename T>
TD_DEVICE void fcollect_block(const SymmCtx& c, const Team& t, const Sync& s, T* dst, const T* src, size_t nelems) {
  fcollectmem_impl<2>(c, t, s, dst, src, nelems * sizeof(T));
}

}  // namespace shmem
}  // namespace td
