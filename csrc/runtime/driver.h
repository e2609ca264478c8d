// Lazy CUDA driver-API entry points resolved through the runtime (cudaGetDriverEntryPoint), so that
// libtd_b200.so has no link-time dependency on libcuda.so.1 and can be dlopen'ed on a CPU-only box
// (the "does it build / import" check) while still using VMM, multicast, stream mem-ops and TMA descriptors
// on a GPU box.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#define TD_API extern "C" __attribute__((visibility("default")))

namespace td {
namespace drv {

inline thread_local char g_last_error[1024] = {0};
inline void set_error(const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_last_error, sizeof(g_last_error), fmt, a, b);
}

template <typename Fn>
inline Fn resolve(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    (void)cudaGetLastError();
    return nullptr;
  }
  return reinterpret_cast<Fn>(fn);
}

#define TD_DRV_FN(name)                                               \
  inline decltype(&::name) name##_fn() {                              \
    static decltype(&::name) f = resolve<decltype(&::name)>(#name);   \
    return f;                                                         \
  }

TD_DRV_FN(cuTensorMapEncodeTiled)
TD_DRV_FN(cuMemCreate)
TD_DRV_FN(cuMemRelease)
TD_DRV_FN(cuMemMap)
TD_DRV_FN(cuMemUnmap)
TD_DRV_FN(cuMemAddressReserve)
TD_DRV_FN(cuMemAddressFree)
TD_DRV_FN(cuMemSetAccess)
TD_DRV_FN(cuMemGetAllocationGranularity)
TD_DRV_FN(cuMemExportToShareableHandle)
TD_DRV_FN(cuMemImportFromShareableHandle)
TD_DRV_FN(cuMulticastCreate)
TD_DRV_FN(cuMulticastAddDevice)
TD_DRV_FN(cuMulticastBindMem)
TD_DRV_FN(cuMulticastGetGranularity)
TD_DRV_FN(cuMulticastUnbind)
TD_DRV_FN(cuStreamWriteValue32)
TD_DRV_FN(cuStreamWaitValue32)
TD_DRV_FN(cuStreamWriteValue64)
TD_DRV_FN(cuStreamWaitValue64)
TD_DRV_FN(cuGetErrorString)
TD_DRV_FN(cuDeviceGetAttribute)
TD_DRV_FN(cuCtxGetDevice)

inline const char* err_str(CUresult r) {
  const char* s = nullptr;
  auto f = cuGetErrorString_fn();
  if (f && f(r, &s) == CUDA_SUCCESS && s) return s;
  return "unknown CUresult";
}

#define TD_CU_CHECK(expr)                                                        \
  do {                                                                           \
    CUresult _r = (expr);                                                        \
    if (_r != CUDA_SUCCESS) {                                                    \
      td::drv::set_error("%s failed: %s", #expr, td::drv::err_str(_r));          \
      return -1;                                                                 \
    }                                                                            \
  } while (0)

#define TD_CUDA_CHECK(expr)                                                      \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      td::drv::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));        \
      return -1;                                                                 \
    }                                                                            \
  } while (0)

}  // namespace drv
}  // namespace td
