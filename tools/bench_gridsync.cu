// micro-benchmark: cost of a grid-wide barrier (cooperative_groups vs hand-rolled) on this GPU
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void k_cg(int n, double* out) {
  cg::grid_group g = cg::this_grid();
  double v = 0;
  for (int i = 0; i < n; ++i) { v += 1.0; g.sync(); }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void my_barrier(unsigned* ctr, unsigned* gen, unsigned nb, unsigned& lg) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned g = lg;
    __threadfence();
    if (atomicAdd(ctr, 1u) == nb - 1) { atomicExch(ctr, 0u); __threadfence(); atomicAdd(gen, 1u); }
    else { while (ld_acquire(gen) == g) {} }
    lg = g + 1;
  }
  __syncthreads();
}
__global__ void k_my(int n, unsigned* bar, double* out) {
  unsigned lg = 0; double v = 0;
  for (int i = 0; i < n; ++i) { v += 1.0; my_barrier(bar, bar + 32, gridDim.x, lg); }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = v;
}
int main() {
  double* out; unsigned* bar; cudaMalloc(&out, 8); cudaMalloc(&bar, 256); 
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int n = 2000;
  for (int grid : {1, 8, 63, 148, 296}) for (int block : {256}) {
    void* args[] = { &n, &out };
    cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(block), args, 0, 0); cudaDeviceSynchronize();
    cudaEventRecord(e0); cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(block), args, 0, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemset(bar, 0, 256);
    void* args2[] = { &n, &bar, &out };
    cudaLaunchCooperativeKernel((void*)k_my, dim3(grid), dim3(block), args2, 0, 0); cudaDeviceSynchronize();
    cudaMemset(bar, 0, 256);
    cudaEventRecord(e0); cudaLaunchCooperativeKernel((void*)k_my, dim3(grid), dim3(block), args2, 0, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms2; cudaEventElapsedTime(&ms2, e0, e1);
    printf("grid %3d block %3d: cg.sync %.3f us  hand-rolled %.3f us  (%s)\n", grid, block, 1e3 * ms / n, 1e3 * ms2 / n, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
