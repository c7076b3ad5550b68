// Scratch micro-benchmark: time of one 48x48x48 fp64 tile product from shared memory with the fragment layout of k_tc_factor
// (6 x 3 per thread, 128 threads), at one and two CTAs per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. tools/bench_tcgemm.cu -o tools/bench_tcgemm
#include <cstdio>
#include "../mavmap_b200/csrc/tilechol.cuh"
using namespace mm;
__global__ void __launch_bounds__(128) k_gemm_loop(int iters, double* out) {
  extern __shared__ __align__(128) double sm[];
  double* A = sm; double* B = sm + TC_TT;
  for (int i = threadIdx.x; i < 2 * TC_TT; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int tid = threadIdx.x, tr = tid >> 4, tcn = tid & 15;
  double acc[6][3];
  for (int a = 0; a < 6; ++a) for (int b = 0; b < 3; ++b) acc[a][b] = 0.0;
  for (int it = 0; it < iters; ++it) { tc_frag_gemm<true>(acc, A, B, 6 * tr, 3 * tcn); asm volatile("" ::: "memory"); }
  double s = 0; for (int a = 0; a < 6; ++a) for (int b = 0; b < 3; ++b) s += acc[a][b];
  out[blockIdx.x * 128 + tid] = s;
}
int main() {
  double* out; cudaMalloc(&out, sizeof(double) * 128 * 2048);
  cudaFuncSetAttribute(k_gemm_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  for (int cfg = 0; cfg < 4; ++cfg) {
    const int grid = cfg == 0 ? 148 : (cfg == 1 ? 296 : (cfg == 2 ? 592 : 1184));
    const size_t smem = cfg <= 1 ? 111 * 1024 : 2 * TC_TT * sizeof(double);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0); k_gemm_loop<<<grid, 128, smem>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("grid %4d smem %6zu: %.3f ms -> %.3f us per tile product per CTA, %.1f TFLOP/s aggregate\n", grid, smem, ms, 1e3 * ms / iters,
                           2.0 * TC_T * TC_T * TC_T * iters * grid / (ms * 1e-3) / 1e12);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
