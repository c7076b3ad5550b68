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

// RAII device buffer on the device's stream-ordered memory pool.  A bundle-adjustment call allocates ~70 buffers; with
// plain cudaMalloc/cudaFree the map/unmap work dominated the end-to-end time of mm_ba_solve (measured: 230 of 360 ms on
// cfg2).  The pool keeps freed blocks (release threshold = never), so repeated calls reuse them.  Semantics are those of
// cudaMalloc/cudaFree: the allocation is complete when alloc() returns, and release() waits for the device to be idle.
inline cudaStream_t alloc_stream() {
  static cudaStream_t st = [] {
    cudaStream_t t = nullptr; cudaStreamCreateWithFlags(&t, cudaStreamNonBlocking);
    int dev = 0; cudaGetDevice(&dev);
    cudaMemPool_t pool; if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) { uint64_t thr = UINT64_MAX; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    return t; }();
  return st;
}
template <typename T>
struct DevBuf {
  T* p = nullptr; size_t n = 0; bool owned = true;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() { if (p && owned) { cudaDeviceSynchronize(); cudaFreeAsync(p, alloc_stream()); } p = nullptr; n = 0; owned = true; }
  void view(T* q, size_t count) { release(); p = q; n = count; owned = false; }      // non-owning window into another buffer
  cudaError_t alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    cudaError_t e = cudaMallocAsync((void**)&p, count * sizeof(T), alloc_stream());
    if (e == cudaSuccess) e = cudaStreamSynchronize(alloc_stream());
    if (e == cudaSuccess) n = count; else p = nullptr;
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
