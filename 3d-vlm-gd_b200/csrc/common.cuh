// Shared host/device helpers for lib3dgd (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>

#include "../../include/gd3.h"

#include <cstdint>
#include <cstdio>
#include <cstdarg>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "lib3dgd is written for sm_100a (B200) only"
#endif

namespace gd3 {

// ---------------------------------------------------------------------------------------------
// error reporting (thread-local message, C-ABI returns a negative code)
// ---------------------------------------------------------------------------------------------
// error codes are the GD3_ERR_* macros of include/gd3.h

void set_error(const char* fmt, ...);
const char* last_error();

#define GD3_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::gd3::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                       __LINE__);                                                         \
      return GD3_ERR_CUDA;                                                         \
    }                                                                                     \
  } while (0)

#define GD3_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::gd3::set_error(__VA_ARGS__);           \
      return GD3_ERR_INVALID;           \
    }                                          \
  } while (0)

// every kernel launch of the library goes through GD3_CHECK_LAUNCH: it also feeds gd3_launch_count()
void count_launch();
#define GD3_CHECK_LAUNCH()                \
  do {                                    \
    ::gd3::count_launch();                \
    GD3_CHECK_CUDA(cudaGetLastError());   \
  } while (0)

// Optional per-kernel timing with CUDA events on the launching stream (gd3_profile_enable).  A scope
// brackets one launch; durations are resolved lazily by gd3_profile_read after a synchronise.
struct ProfScope {
  const char* name;
  cudaStream_t stream;
  int slot;
  ProfScope(const char* name, cudaStream_t stream);
  ~ProfScope();
};
#define GD3_PROF(name, stream) ::gd3::ProfScope _gd3_prof_scope(name, stream)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Opt-in dynamic shared memory of one kernel.  cudaFuncSetAttribute is per device and the size may grow between calls,
// so remember the largest size configured on each device (one static instance per kernel at the launch site).
struct SmemOptIn {
  static constexpr int kMaxDevices = 64;
  size_t bytes[kMaxDevices] = {};
  template <class Kernel>
  cudaError_t ensure(Kernel kernel, size_t smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    const bool tracked = dev >= 0 && dev < kMaxDevices;
    if (tracked && bytes[dev] >= smem) return cudaSuccess;
    if (smem > 48 * 1024 || !tracked) {
      const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    if (tracked) bytes[dev] = smem;
    return cudaSuccess;
  }
};

template <class T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <class T>
__host__ __device__ constexpr T round_up(T a, T b) { return ceil_div(a, b) * b; }

// workspace carving: every sub-buffer 256-byte aligned
struct Carver {
  uint8_t* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<uint8_t*>(p)), off(0) {}
  template <class T>
  T* take(size_t count) {
    off = round_up<size_t>(off, 256);
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return r;
  }
  size_t total() const { return round_up<size_t>(off, 256); }
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum for blockDim.x <= 1024 (result valid in every thread)
template <class T>
__device__ __forceinline__ T block_sum(T v, T* red /* >= 32 entries of smem */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  T r = (lane < nw) ? red[lane] : T(0);
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ float bf16_bits_to_float(uint32_t hi16) { return __uint_as_float(hi16 << 16); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

#endif  // __CUDACC__

}  // namespace gd3
