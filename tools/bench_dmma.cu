// Scratch micro-benchmark: one 48x48x48 fp64 tile product C -= A B' from shared memory, 128 threads per CTA,
//   (a) the 6 x 3 register fragment on the fp64 FMA pipe (what k_tc_factor runs today),
//   (b) mma.sync.m8n8k4.f64 (DMMA), each warp a 24 x 24 quadrant = 3 x 3 MMA tiles, operands column-major with leading dimension 48
//       (the layout the tiles have in global memory: 4-way bank conflicts on the fragment loads),
//   (c) DMMA with the rows of every column XOR-swizzled by 4 * (column & 3) (conflict-free fragment loads).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. tools/bench_dmma.cu -o tools/bench_dmma
#include <cstdio>
#include "../mavmap_b200/csrc/tilechol.cuh"
using namespace mm;

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <bool SWZ> __device__ __forceinline__ int elem(int r, int q) { return q * TC_T + (SWZ ? (r ^ ((q & 3) << 2)) : r); }

// acc[i][j][2]: MMA tile (i, j) of the warp's 24 x 24 quadrant; C(r, c) -= sum_q A(r, q) B(c, q)
template <bool SWZ>
__device__ __forceinline__ void dmma_tile_product(double (&acc)[3][3][2], const double* A, const double* B, int r0, int c0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll 2
  for (int q0 = 0; q0 < TC_T; q0 += 4) {
    double a[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { a[i] = -A[elem<SWZ>(r0 + 8 * i + g, q0 + t)]; b[i] = B[elem<SWZ>(c0 + 8 * i + g, q0 + t)]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

template <int MODE>
__global__ void __launch_bounds__(128) k_loop(int iters, double* out) {
  extern __shared__ __align__(128) double sm[];
  double* A = sm; double* B = sm + TC_TT;
  for (int i = threadIdx.x; i < 2 * TC_TT; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  double s = 0;
  if (MODE == 0) {
    const int tid = threadIdx.x, tr = tid >> 4, tcn = tid & 15;
    double acc[6][3];
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 3; ++b) acc[a][b] = 0.0;
    for (int it = 0; it < iters; ++it) { tc_frag_gemm<true>(acc, A, B, 6 * tr, 3 * tcn); asm volatile("" ::: "memory"); }
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 3; ++b) s += acc[a][b];
  } else {
    const int w = threadIdx.x >> 5;
    double acc[3][3][2];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int it = 0; it < iters; ++it) { dmma_tile_product<MODE == 2>(acc, A, B, 24 * (w >> 1), 24 * (w & 1)); asm volatile("" ::: "memory"); }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s += acc[i][j][0] + acc[i][j][1];
  }
  out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, double* out) {
  cudaFuncSetAttribute(k_loop<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int grid = cfg == 0 ? 148 : (cfg == 1 ? 296 : 592);
    const size_t smem = cfg <= 1 ? 111 * 1024 : 2 * TC_TT * sizeof(double);       // 1, 2 and 4 CTAs per SM
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); k_loop<MODE><<<grid, 128, smem>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    printf("%-28s %d CTA/SM: %.3f us per tile product per CTA, %.1f TFLOP/s aggregate\n", name, grid / 148, 1e3 * ms / iters, 2.0 * TC_T * TC_T * TC_T * iters * grid / (ms * 1e-3) / 1e12);
  }
}
int main() {
  double* out; cudaMalloc(&out, sizeof(double) * 128 * 2048);
  run<0>("FMA 6x3 fragment", out); run<1>("DMMA, ld 48 (conflicts)", out); run<2>("DMMA, swizzled rows", out);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
