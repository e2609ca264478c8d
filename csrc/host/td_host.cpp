// CPU emulation runtime: the same symmetric-heap / signal model as the GPU runtime, on POSIX shared
// memory with C++11 atomics.  It lets every protocol (flag lifecycles, phase counters, ring orders,
// barriers, dispatch/combine slot accounting) run under pytest with `gloo` and world_size > 1 on a box
// without GPUs.  The reference has no such backend (its tests need GPUs + NVSHMEM:
// /root/reference/python/triton_dist/utils.py:51-148); SURVEY.md section 4 asks for one.
//
// Layout mirrors csrc/runtime/symm_heap.cu: rank r's segment is mapped at  base + r * stride  inside one
// PROT_NONE reservation, so symm_at() is the same arithmetic on both backends.
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

#define TDH_API extern "C" __attribute__((visibility("default")))

namespace {
thread_local char g_err[512] = {0};

struct HostHeap {
  std::string prefix;
  int rank = 0, world = 1;
  size_t bytes = 0, stride = 0;
  uint8_t* base = nullptr;
  int local_fd = -1;
};

std::string seg_name(const std::string& prefix, int r) { return "/" + prefix + "_r" + std::to_string(r); }

inline std::atomic<uint32_t>* a32(void* p) { return reinterpret_cast<std::atomic<uint32_t>*>(p); }
inline std::atomic<uint64_t>* a64(void* p) { return reinterpret_cast<std::atomic<uint64_t>*>(p); }
}  // namespace

TDH_API const char* tdh_last_error() { return g_err; }

// Phase 1 (all ranks): create my own segment.
TDH_API void* tdh_heap_create(const char* prefix, int rank, int world, unsigned long long bytes) {
  const size_t page = 1u << 16;
  HostHeap* h = new HostHeap();
  h->prefix = prefix; h->rank = rank; h->world = world;
  h->bytes = ((bytes + page - 1) / page) * page;
  h->stride = h->bytes;
  const std::string name = seg_name(h->prefix, rank);
  shm_unlink(name.c_str());
  h->local_fd = shm_open(name.c_str(), O_CREAT | O_RDWR | O_EXCL, 0600);
  if (h->local_fd < 0) { snprintf(g_err, sizeof(g_err), "shm_open(%s): %s", name.c_str(), strerror(errno)); delete h; return nullptr; }
  if (ftruncate(h->local_fd, static_cast<off_t>(h->bytes)) != 0) {
    snprintf(g_err, sizeof(g_err), "ftruncate: %s", strerror(errno)); close(h->local_fd); shm_unlink(name.c_str()); delete h; return nullptr;
  }
  return h;
}

// Phase 2 (after a host barrier): map every rank's segment at base + r * stride.
TDH_API int tdh_heap_map(void* hp) {
  HostHeap* h = static_cast<HostHeap*>(hp);
  void* res = mmap(nullptr, h->stride * h->world, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (res == MAP_FAILED) { snprintf(g_err, sizeof(g_err), "reserve: %s", strerror(errno)); return -1; }
  h->base = static_cast<uint8_t*>(res);
  for (int r = 0; r < h->world; ++r) {
    int fd = (r == h->rank) ? h->local_fd : shm_open(seg_name(h->prefix, r).c_str(), O_RDWR, 0600);
    if (fd < 0) { snprintf(g_err, sizeof(g_err), "shm_open peer %d: %s", r, strerror(errno)); return -1; }
    void* m = mmap(h->base + r * h->stride, h->bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);
    if (m == MAP_FAILED) { snprintf(g_err, sizeof(g_err), "map peer %d: %s", r, strerror(errno)); return -1; }
    if (r != h->rank) close(fd);
  }
  return 0;
}
// Phase 3 (after another host barrier): names can be unlinked, the mappings keep the memory alive.
TDH_API int tdh_heap_unlink(void* hp) {
  HostHeap* h = static_cast<HostHeap*>(hp);
  shm_unlink(seg_name(h->prefix, h->rank).c_str());
  return 0;
}
TDH_API unsigned long long tdh_heap_base(void* hp) { return reinterpret_cast<unsigned long long>(static_cast<HostHeap*>(hp)->base); }
TDH_API unsigned long long tdh_heap_stride(void* hp) { return static_cast<HostHeap*>(hp)->stride; }
TDH_API unsigned long long tdh_heap_bytes(void* hp) { return static_cast<HostHeap*>(hp)->bytes; }
TDH_API int tdh_heap_destroy(void* hp) {
  HostHeap* h = static_cast<HostHeap*>(hp);
  if (!h) return 0;
  if (h->base) munmap(h->base, h->stride * h->world);
  if (h->local_fd >= 0) close(h->local_fd);
  shm_unlink(seg_name(h->prefix, h->rank).c_str());
  delete h;
  return 0;
}

// ---- chaos mode: TD_HOST_CHAOS_US=N makes every notify / first wait poll sleep a random 0..N microseconds, i.e. ranks
// drift apart by arbitrary amounts between protocol steps.  Parity double-buffering, phase counters and slot reuse must
// survive any such schedule (tests/test_dist_cpu.py::test_cpu_chaos); the reference only has hand-placed stragglers.
static long long chaos_us() {
  static long long v = -1;
  if (v < 0) { const char* e = getenv("TD_HOST_CHAOS_US"); v = e ? atoll(e) : 0; }
  return v;
}
static void chaos() {
  const long long n = chaos_us();
  if (n <= 0) return;
  static thread_local unsigned long long st = 0x9E3779B97F4A7C15ull ^ static_cast<unsigned long long>(getpid()) * 0xD1B54A32D192ED03ull;
  st ^= st << 13; st ^= st >> 7; st ^= st << 17;                       // xorshift64
  std::this_thread::sleep_for(std::chrono::microseconds(static_cast<long long>(st % static_cast<unsigned long long>(n + 1))));
}

// ---- signal primitives (same semantics as td::notify / td::wait on the device) ----------------------
// op: 1 = SET, 2 = ADD  (DistributedAttrDefs.td:36-44)
TDH_API void tdh_notify32(void* addr, unsigned int value, int op) {
  chaos();
  if (op == 2) a32(addr)->fetch_add(value, std::memory_order_release);
  else a32(addr)->store(value, std::memory_order_release);
}
TDH_API void tdh_notify64(void* addr, unsigned long long value, int op) {
  chaos();
  if (op == 2) a64(addr)->fetch_add(value, std::memory_order_release);
  else a64(addr)->store(value, std::memory_order_release);
}
TDH_API unsigned int tdh_ld_acquire32(void* addr) { return a32(addr)->load(std::memory_order_acquire); }
TDH_API unsigned long long tdh_ld_acquire64(void* addr) { return a64(addr)->load(std::memory_order_acquire); }
TDH_API unsigned int tdh_atomic_add32(void* addr, unsigned int v) { return a32(addr)->fetch_add(v, std::memory_order_acq_rel); }
TDH_API unsigned int tdh_atomic_cas32(void* addr, unsigned int cmp, unsigned int val) {
  a32(addr)->compare_exchange_strong(cmp, val, std::memory_order_acq_rel);
  return cmp;
}
// cmp: 0 = EQ, 1 = GE (signed distance, wrap-safe).  Returns 0 on success, 1 on timeout (hang detection).
TDH_API int tdh_wait32(void* addr, unsigned int value, int cmp, long long timeout_us) {
  chaos();
  const auto t0 = std::chrono::steady_clock::now();
  unsigned spins = 0;
  while (true) {
    const uint32_t v = a32(addr)->load(std::memory_order_acquire);
    if (cmp == 0 ? (v == value) : (static_cast<int32_t>(v - value) >= 0)) return 0;
    if (++spins > 64) {
      std::this_thread::yield();
      if (timeout_us > 0 && (spins & 0x3ff) == 0) {
        const auto dt = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        if (dt > timeout_us) return 1;
      }
    }
  }
}
// wait on n consecutive flags (the warp-cooperative dl.wait of the reference: distributed_ops.py:53-70)
TDH_API int tdh_wait32_n(void* addr, int n, unsigned int value, int cmp, long long timeout_us) {
  for (int i = 0; i < n; ++i) {
    int r = tdh_wait32(static_cast<uint32_t*>(addr) + i, value, cmp, timeout_us);
    if (r) return r;
  }
  return 0;
}
TDH_API void tdh_fence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

// Flag-flip barrier over the heap: slots = uint32 [2][world] at the same offset in every segment.
// Identical algorithm to td::barrier_all_block (csrc/td/primitives.cuh).
TDH_API int tdh_barrier_all(void* hp, unsigned long long slots_off, unsigned int epoch, long long timeout_us) {
  HostHeap* h = static_cast<HostHeap*>(hp);
  const size_t arr = slots_off + static_cast<size_t>(epoch & 1u) * h->world * 4;
  for (int t = 0; t < h->world; ++t) {
    uint8_t* peer_slot = h->base + t * h->stride + arr + h->rank * 4;
    a32(peer_slot)->store(epoch, std::memory_order_release);
  }
  for (int t = 0; t < h->world; ++t) {
    uint8_t* my_slot = h->base + h->rank * h->stride + arr + t * 4;
    if (tdh_wait32(my_slot, epoch, 1, timeout_us)) { snprintf(g_err, sizeof(g_err), "barrier timeout waiting for rank %d", t); return 1; }
  }
  return 0;
}

TDH_API void tdh_memcpy(void* dst, const void* src, unsigned long long n) { memcpy(dst, src, n); }

// ------------------------------------------------------------------------------------------------
// DLPack capsule payloads built and destroyed in C: the deleter of a tensor that aliases heap memory can run at any time (another
// thread, the garbage collector, interpreter shutdown) -- it must not be a Python callback.
// ------------------------------------------------------------------------------------------------
namespace {
struct DlDevice { int32_t device_type; int32_t device_id; };
struct DlDataType { uint8_t code; uint8_t bits; uint16_t lanes; };
struct DlTensor { void* data; DlDevice device; int32_t ndim; DlDataType dtype; int64_t* shape; int64_t* strides; uint64_t byte_offset; };
struct DlManagedTensor { DlTensor dl_tensor; void* manager_ctx; void (*deleter)(DlManagedTensor*); };
struct DlOwned { DlManagedTensor mt; int64_t shape[8]; };
void dl_delete(DlManagedTensor* m) { delete reinterpret_cast<DlOwned*>(m); }
}  // namespace

TDH_API void* tdh_dl_make(void* data, int ndim, const int64_t* shape, int code, int bits, int device_type, int device_id) {
  if (ndim < 0 || ndim > 8) return nullptr;
  DlOwned* o = new DlOwned();
  for (int i = 0; i < ndim; ++i) o->shape[i] = shape[i];
  o->mt.dl_tensor.data = data;
  o->mt.dl_tensor.device = {device_type, device_id};
  o->mt.dl_tensor.ndim = ndim;
  o->mt.dl_tensor.dtype = {static_cast<uint8_t>(code), static_cast<uint8_t>(bits), 1};
  o->mt.dl_tensor.shape = o->shape;
  o->mt.dl_tensor.strides = nullptr;
  o->mt.dl_tensor.byte_offset = 0;
  o->mt.manager_ctx = nullptr;
  o->mt.deleter = dl_delete;
  return &o->mt;
}
