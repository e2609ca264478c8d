// Hang detection for the spin loops of the device primitives (debug library variant: -DTD_WAIT_TIMEOUT_NS=<ns>, selected with
// TD_DEBUG_WAITS=<milliseconds>).  Every guarded loop checks the global timer once per 1024 polls and, past the deadline, prints what
// it was waiting for (block, thread, address, observed / expected value) and traps: the launch fails with a diagnostic instead of
// spinning forever.  The reference's waits -- like the default build's -- have no timeout.
//
// Without the macro both hooks expand to nothing.  ptx.cuh / primitives.cuh place them on EXISTING source lines and include this file
// in place of a blank line, so the default build is bit-identical to the one that was validated on hardware (checked: the SASS of
// gemm_sm100.cu, flash_attn_sm100.cu, comm_kernels.cu, megakernel.cu does not change).
#pragma once
#ifdef TD_WAIT_TIMEOUT_NS
#include <cstdio>
namespace td {
namespace ptx {
struct SpinGuard {
  unsigned long long t0;
  unsigned int n;
  __device__ __forceinline__ static unsigned long long now() {
    unsigned long long r;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(r));
    return r;
  }
  __device__ __forceinline__ SpinGuard() : t0(now()), n(0) {}
  __device__ __forceinline__ void poll(const char* what, const void* addr, unsigned long long observed, unsigned long long expected) {
    if ((++n & 0x3FFu) == 0u && now() - t0 > static_cast<unsigned long long>(TD_WAIT_TIMEOUT_NS)) {
      printf("[td] %s timed out after %llu ms: block (%d,%d,%d) thread %d, address %p, observed %llu, expected %llu\n", what,
             static_cast<unsigned long long>(TD_WAIT_TIMEOUT_NS) / 1000000ull, static_cast<int>(blockIdx.x), static_cast<int>(blockIdx.y),
             static_cast<int>(blockIdx.z), static_cast<int>(threadIdx.x), addr, observed, expected);
      __trap();
    }
  }
};
}  // namespace ptx
}  // namespace td
#define TD_SPIN_GUARD(g) ::td::ptx::SpinGuard g;
#define TD_SPIN_POLL(g, what, addr, observed, expected) g.poll(what, addr, observed, expected);
#else
#define TD_SPIN_GUARD(g)
#define TD_SPIN_POLL(g, what, addr, observed, expected)
#endif
