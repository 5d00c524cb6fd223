// Shared host/device helpers for libeemflow_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>

#include "../../include/eemflow_b200.h"

namespace eem {

// Thread-local last-error text (eem_last_error_string).  Defined in api.cu.
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

inline cudaStream_t as_stream(eem_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device, cached per device.  Defined in api.cu.
int sm_count();

// Kernel launches issued by this library in this process (eem_launch_count).  Defined in api.cu.
void count_launch();

#define EEM_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return ::eem::fail(EEM_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

#define EEM_CHECK_ALIGNED(ptr, bytes)                                                           \
  do {                                                                                          \
    if ((reinterpret_cast<uintptr_t>(ptr) % (bytes)) != 0)                                      \
      return ::eem::fail(EEM_ERR_MISALIGNED, "%s must be %d-byte aligned", #ptr, (int)(bytes)); \
  } while (0)

// Call after every launch: picks up launch-configuration errors without synchronising.
#define EEM_CHECK_LAUNCH(name)                                                              \
  do {                                                                                      \
    ::eem::count_launch();                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                    \
    if (e_ != cudaSuccess)                                                                  \
      return ::eem::fail(EEM_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e_));             \
  } while (0)

#define EEM_CHECK_CUDA(expr)                                                                \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return ::eem::fail(EEM_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));            \
  } while (0)

// Opt-in for more than 48 KiB of dynamic shared memory.  cudaFuncSetAttribute applies to the CURRENT device's
// context only, so the "already configured" cache is kept per device: a process that touches a second GPU
// (nn.DataParallel as in train_EEMFlow_HREM.py:117, a voxelizer with gpu_nr=1, a model moved to cuda:1) sets the
// attribute there as well.  One instance (a function-local static) per kernel instantiation.
constexpr int kMaxDevices = 64;
struct DynSmemOptIn {
  std::mutex mu;
  size_t configured[kMaxDevices] = {};
  template <class Kernel>
  cudaError_t ensure(Kernel kernel, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices)
      return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    std::lock_guard<std::mutex> lock(mu);
    if (bytes > configured[dev]) {
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e == cudaSuccess) configured[dev] = bytes;
    }
    return e;
  }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers -----------------------------------------------------------------------

__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace eem
