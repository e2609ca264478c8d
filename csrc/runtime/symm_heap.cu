// Symmetric heap on CUDA VMM -- the NVSHMEM replacement for a single NVSwitch domain.
//
// Reference behaviour being replaced: nvshmem.core.init + nvshmem.core.tensor / get_peer_tensor
// (/root/reference/python/triton_dist/utils.py:223-287) and nvshmem_ptr / nvshmemx_mc_ptr on the device.
//
// Design: every rank cuMemCreate()s ONE physical segment of `bytes` (POSIX-fd shareable), the fds are
// exchanged by the Python layer over a unix socket (SCM_RIGHTS), and every rank maps all `world`
// segments into ONE virtual-address reservation at  base + r * stride .  Peer translation on the device
// is then pure arithmetic (td::symm_at), there is no pointer table to load.  If the fabric supports it,
// a multicast object is bound over all segments and mapped once more (NVLS: multimem.ld_reduce/st/red).
#include "driver.h"
#include <unistd.h>
#include <vector>

namespace {

struct Heap {
  int dev = 0;
  int rank = 0, world = 1;
  size_t bytes = 0;       // usable bytes per rank
  size_t stride = 0;      // VA distance between ranks (== bytes rounded to granularity)
  CUdeviceptr base = 0;   // VA of rank 0
  CUmemGenericAllocationHandle local = 0;
  std::vector<CUmemGenericAllocationHandle> peers;
  bool mapped = false;
  // multicast
  CUmemGenericAllocationHandle mc = 0;
  CUdeviceptr mc_base = 0;
  bool mc_mapped = false;
};

CUmemAllocationProp make_prop(int dev) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

int set_access(CUdeviceptr p, size_t n, int dev) {
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  TD_CU_CHECK(td::drv::cuMemSetAccess_fn()(p, n, &acc, 1));
  return 0;
}

}  // namespace

using namespace td::drv;

// Returns an opaque heap handle (or null).  `bytes` is rounded up to the allocation granularity.
TD_API void* td_heap_create(int dev, int rank, int world, unsigned long long bytes) {
  if (!cuMemCreate_fn()) { set_error("CUDA driver VMM API unavailable"); return nullptr; }
  if (cudaSetDevice(dev) != cudaSuccess || cudaFree(0) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
  Heap* h = new Heap();
  h->dev = dev; h->rank = rank; h->world = world;
  CUmemAllocationProp prop = make_prop(dev);
  size_t gran = 0;
  if (cuMemGetAllocationGranularity_fn()(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) gran = 2u << 20;
  // multicast binding wants a (possibly larger) granularity: take the max so the same segment can be bound
  if (cuMulticastGetGranularity_fn()) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = world > 1 ? world : 2;
    mp.size = gran;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (cuMulticastGetGranularity_fn()(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > gran) gran = mg;
  }
  h->bytes = ((bytes + gran - 1) / gran) * gran;
  h->stride = h->bytes;
  CUresult r = cuMemCreate_fn()(&h->local, h->bytes, &prop, 0);
  if (r != CUDA_SUCCESS) { set_error("cuMemCreate failed: %s", err_str(r)); delete h; return nullptr; }
  h->peers.assign(world, 0);
  h->peers[rank] = h->local;
  return h;
}

// POSIX fd for my segment (caller sends it to the peers with SCM_RIGHTS and closes it afterwards).
TD_API int td_heap_export_fd(void* hp) {
  Heap* h = static_cast<Heap*>(hp);
  int fd = -1;
  CUresult r = cuMemExportToShareableHandle_fn()(&fd, h->local, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) { set_error("cuMemExportToShareableHandle failed: %s", err_str(r)); return -1; }
  return fd;
}

// fds[r] = fd received from rank r (ignored for r == my rank).  Maps all segments; zero-fills mine.
TD_API int td_heap_map(void* hp, const int* fds) {
  Heap* h = static_cast<Heap*>(hp);
  for (int r = 0; r < h->world; ++r) {
    if (r == h->rank) continue;
    TD_CU_CHECK(cuMemImportFromShareableHandle_fn()(&h->peers[r], reinterpret_cast<void*>(static_cast<uintptr_t>(fds[r])),
                                                    CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  }
  TD_CU_CHECK(cuMemAddressReserve_fn()(&h->base, h->stride * h->world, h->stride < (1ull << 30) ? 0 : 0, 0, 0));
  for (int r = 0; r < h->world; ++r) {
    TD_CU_CHECK(cuMemMap_fn()(h->base + r * h->stride, h->bytes, 0, h->peers[r], 0));
  }
  if (set_access(h->base, h->stride * h->world, h->dev)) return -1;
  h->mapped = true;
  TD_CUDA_CHECK(cudaMemset(reinterpret_cast<void*>(h->base + h->rank * h->stride), 0, h->bytes));
  TD_CUDA_CHECK(cudaDeviceSynchronize());
  return 0;
}

TD_API unsigned long long td_heap_base(void* hp) { return static_cast<Heap*>(hp)->base; }
TD_API unsigned long long td_heap_stride(void* hp) { return static_cast<Heap*>(hp)->stride; }
TD_API unsigned long long td_heap_bytes(void* hp) { return static_cast<Heap*>(hp)->bytes; }
TD_API unsigned long long td_heap_mc_base(void* hp) { return static_cast<Heap*>(hp)->mc_base; }

// ---- multicast (NVLS) ----------------------------------------------------------------------------
TD_API int td_multicast_supported(int dev) {
  auto f = cuDeviceGetAttribute_fn();
  if (!f || !cuMulticastCreate_fn()) return 0;
  int v = 0;
  if (f(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return 0;
  return v;
}
// rank 0: create the multicast object and return its fd (others pass the received fd to td_heap_mc_import)
TD_API int td_heap_mc_create(void* hp) {
  Heap* h = static_cast<Heap*>(hp);
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = h->world;
  mp.size = h->bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUresult r = cuMulticastCreate_fn()(&h->mc, &mp);
  if (r != CUDA_SUCCESS) { set_error("cuMulticastCreate failed: %s", err_str(r)); return -1; }
  int fd = -1;
  r = cuMemExportToShareableHandle_fn()(&fd, h->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) { set_error("export multicast handle failed: %s", err_str(r)); return -1; }
  return fd;
}
TD_API int td_heap_mc_import(void* hp, int fd) {
  Heap* h = static_cast<Heap*>(hp);
  TD_CU_CHECK(cuMemImportFromShareableHandle_fn()(&h->mc, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                                  CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  return 0;
}
// every rank, after all ranks hold the handle
TD_API int td_heap_mc_add_device(void* hp) {
  Heap* h = static_cast<Heap*>(hp);
  TD_CU_CHECK(cuMulticastAddDevice_fn()(h->mc, h->dev));
  return 0;
}
// every rank, after ALL ranks have added their device (caller barriers in between)
TD_API int td_heap_mc_bind_and_map(void* hp) {
  Heap* h = static_cast<Heap*>(hp);
  TD_CU_CHECK(cuMulticastBindMem_fn()(h->mc, 0, h->local, 0, h->bytes, 0));
  TD_CU_CHECK(cuMemAddressReserve_fn()(&h->mc_base, h->bytes, 0, 0, 0));
  TD_CU_CHECK(cuMemMap_fn()(h->mc_base, h->bytes, 0, h->mc, 0));
  if (set_access(h->mc_base, h->bytes, h->dev)) return -1;
  h->mc_mapped = true;
  return 0;
}

TD_API int td_heap_destroy(void* hp) {
  Heap* h = static_cast<Heap*>(hp);
  if (!h) return 0;
  cudaDeviceSynchronize();
  if (h->mc_mapped) {
    cuMemUnmap_fn()(h->mc_base, h->bytes);
    cuMemAddressFree_fn()(h->mc_base, h->bytes);
  }
  if (h->mc) {
    if (h->mc_mapped && cuMulticastUnbind_fn()) cuMulticastUnbind_fn()(h->mc, h->dev, 0, h->bytes);
    cuMemRelease_fn()(h->mc);
  }
  if (h->mapped) {
    for (int r = 0; r < h->world; ++r) cuMemUnmap_fn()(h->base + r * h->stride, h->bytes);
    cuMemAddressFree_fn()(h->base, h->stride * h->world);
  }
  for (int r = 0; r < h->world; ++r)
    if (h->peers[r]) cuMemRelease_fn()(h->peers[r]);
  delete h;
  return 0;
}

// ---- stream-ordered memory operations (copy-engine style signalling) --------------------------------
// Reference: _wait_eq_cuda / _set_signal_cuda (/root/reference/python/triton_dist/kernels/nvidia/common_ops.py:364-414)
TD_API int td_stream_write_value32(void* stream, unsigned long long addr, unsigned int value) {
  TD_CU_CHECK(cuStreamWriteValue32_fn()(reinterpret_cast<CUstream>(stream), addr, value, CU_STREAM_WRITE_VALUE_DEFAULT));
  return 0;
}
TD_API int td_stream_wait_value32(void* stream, unsigned long long addr, unsigned int value, int geq) {
  TD_CU_CHECK(cuStreamWaitValue32_fn()(reinterpret_cast<CUstream>(stream), addr, value,
                                       geq ? CU_STREAM_WAIT_VALUE_GEQ : CU_STREAM_WAIT_VALUE_EQ));
  return 0;
}
TD_API int td_memcpy_async(void* dst, const void* src, unsigned long long bytes, void* stream) {
  TD_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

// ---- device / topology queries ------------------------------------------------------------------------
TD_API int td_device_info(int dev, int* out /* [8]: sms, cc_major, cc_minor, multicast, smem_optin, l2_bytes, clock_khz, mem_clock_khz */) {
  cudaDeviceProp prop;
  TD_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  out[0] = prop.multiProcessorCount; out[1] = prop.major; out[2] = prop.minor;
  out[3] = td_multicast_supported(dev);
  out[4] = static_cast<int>(prop.sharedMemPerBlockOptin);
  out[5] = prop.l2CacheSize;
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev); out[6] = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrMemoryClockRate, dev); out[7] = v;
  return 0;
}
TD_API int td_can_access_peer(int dev, int peer) {
  int v = 0;
  if (cudaDeviceCanAccessPeer(&v, dev, peer) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return v;
}
TD_API int td_p2p_native_atomics(int dev, int peer) {
  int v = 0;
  if (cudaDeviceGetP2PAttribute(&v, cudaDevP2PAttrNativeAtomicSupported, dev, peer) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return v;
}
