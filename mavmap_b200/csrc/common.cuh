// common.cuh — error handling, launch accounting and small device helpers shared by
// every translation unit of libmavmap_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include "../../include/mavmap_b200.h"

namespace mm {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int ensure_device();   // MM_OK or MM_ERR_NO_DEVICE

#define MM_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      mm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MM_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define MM_LAUNCH_CHECK()                                                               \
  do {                                                                                  \
    mm::count_launch();                                                                 \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      mm::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MM_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

// RAII device buffer (plain cudaMalloc; sized for 180 GB HBM, no pooling needed here)
template <typename T>
struct DevBuf {
  T* p = nullptr; size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  cudaError_t alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
};

static inline int num_sms() {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  return sms;
}

// ---- device helpers --------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; result valid in thread 0.  smem: >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  return v;
}
__device__ __forceinline__ double block_max(double v, double* smem) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.0;
  if (w == 0) v = warp_max(v);
  return v;
}
// atomic max for non-negative doubles (IEEE order == integer order for x >= 0)
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}

}  // namespace mm
