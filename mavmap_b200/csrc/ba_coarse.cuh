// ba_coarse.cuh — coarse level of the two-level PCG preconditioner on the reduced camera system.
//
// Why: with block-Jacobi alone the reduced camera system of a survey-type sequence needs 500-2000+ PCG iterations at the
// tolerance that reproduces the reference's direct (sparse Cholesky, bundle_adjustment.cc:555) solve, because its low end of
// the spectrum is made of smooth, near-gauge deformations (locally a similarity transform of a patch of cameras).  The coarse
// space spans exactly those: images are grouped into aggregates (host, greedy over the block graph), each aggregate carries the
// 7 similarity modes (3 translations, 3 rotations, 1 scale about the aggregate centre) expressed in the (rvec, tvec)
// parameters of its images, and the preconditioner becomes   M^-1 = blockdiag(S)^-1 + P (P' S P)^-1 P'   (additive two-level).
// Measured on the oracle's cfg2 system: 880 -> 110 iterations (aggregates of 8), independent of the number of images.
//
// This file: assembly of the coarse matrix Ac = P' S P from the stored 6x6 blocks, and its explicit inverse by a blocked
// in-place Gauss-Jordan sweep in ONE cooperative launch (Ac is symmetric positive definite: no pivoting).  The inverse is
// applied inside the persistent PCG kernels (ba_kernels.cuh) as a dense mat-vec spread over all warps of the grid.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace mm {

constexpr int CM = 7;           // coarse unknowns per aggregate
constexpr int PCS = 6 * CM;     // doubles of one image's prolongation block, row-major [6][7]

// one warp per stored block of S (n_img diagonal blocks, then the off-diagonal blocks (a < b))
__global__ void __launch_bounds__(128) k_coarse_assemble(
    int n_img, int64_t nblk, const int* __restrict__ blk_a, const int* __restrict__ blk_b, const double* __restrict__ S,
    const int* __restrict__ agg, const double* __restrict__ Pc, int m, double* __restrict__ Ac) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nblk) return;
  const int ia = b < n_img ? (int)b : blk_a[b - n_img], ib = b < n_img ? (int)b : blk_b[b - n_img];
  const double* Sb = S + 36 * (size_t)b;
  const double* Pa = Pc + PCS * (size_t)ia;
  const double* Pb = Pc + PCS * (size_t)ib;
  const int ga = CM * agg[ia], gb = CM * agg[ib];
  for (int o = lane; o < CM * CM; o += 32) {
    const int i = o / CM, j = o - CM * i;
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double u = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) u += Sb[6 * r + c] * Pb[CM * c + j];
      t += Pa[CM * r + i] * u;
    }
    if (t != 0.0) {
      atomicAdd(Ac + (size_t)(ga + i) * m + gb + j, t);
      if (ia != ib) atomicAdd(Ac + (size_t)(gb + j) * m + ga + i, t);
    }
  }
}

// a zero diagonal (aggregate without the mode: a single image has no scale mode, fixed images have none at all) decouples
// that unknown; otherwise a relative ridge keeps the unpivoted elimination away from exact singularity
__global__ void k_coarse_ridge(int m, double* __restrict__ Ac) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double d = Ac[(size_t)i * m + i];
  Ac[(size_t)i * m + i] = (d > 0.0) ? d * (1.0 + 1e-10) : 1.0;
}

// In-place inverse of a symmetric positive definite m x m matrix (row-major) by blocked Gauss-Jordan, panel width 32.
// Per panel:  (A) every CTA inverts the 32 x 32 pivot block in shared memory (redundantly: cheaper than a barrier);
// the grid forms the scaled row panel Rk = inv * A[k,:] (with inv itself in the pivot columns) and
// copies the column panel Ck = A[:,k];  grid barrier;  (B) 64 x 64 tiles of A get  A -= Ck Rk  (pivot rows <- Rk, pivot
// columns <- -Ck inv);  grid barrier.  2 m^3 flop, m/32 * 2 barriers.
constexpr int GJ_NB = 32;
__global__ void __launch_bounds__(256, 2) k_spd_inverse(int m, double* __restrict__ A, double* __restrict__ Ck, double* __restrict__ Rk, unsigned long long* dbg) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_inv[GJ_NB][GJ_NB + 1];
  __shared__ double s_c[64][GJ_NB + 1];
  __shared__ double s_r[GJ_NB][64 + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // warp-interleaved global thread index: consecutive warps of work land on different SMs
  const int gtid = (warp * gridDim.x + blockIdx.x) * 32 + lane, gsize = gridDim.x * blockDim.x;
  const int nt = (m + 63) / 64;
#define GJ_STAMP(k) do { if (dbg && blockIdx.x == 0 && tid == 0 && k0 < 8 * GJ_NB) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); dbg[6 * (k0 / GJ_NB) + (k)] = t_; } } while (0)
  for (int k0 = 0; k0 < m; k0 += GJ_NB) {
    const int kb = min(GJ_NB, m - k0);
    GJ_STAMP(0);
    // pivot block -> shared memory (identity padding past the matrix edge), then 32 elimination steps with all threads
    for (int idx = tid; idx < GJ_NB * GJ_NB; idx += 256) {
      const int i = idx >> 5, j = idx & 31;
      s_inv[i][j] = (i < kb && j < kb) ? __ldcg(A + (size_t)(k0 + i) * m + k0 + j) : (i == j ? 1.0 : 0.0);
    }
    for (int c = 0; c < GJ_NB; ++c) {
      __syncthreads();
      const double ip = 1.0 / s_inv[c][c];
      double nv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + 256 * e, i = idx >> 5, j = idx & 31;
        const double own = s_inv[i][j], f = s_inv[i][c], r = s_inv[c][j];
        nv[e] = (i == c) ? (j == c ? ip : r * ip) : (j == c ? -f * ip : own - f * (r * ip));
      }
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e) { const int idx = tid + 256 * e; s_inv[idx >> 5][idx & 31] = nv[e]; }
    }
    __syncthreads();
    GJ_STAMP(1);
    // ---- (A) row panel (thread = column j x group of 4 panel rows) and column panel
    for (int idx = gtid; idx < 8 * m; idx += gsize) {
      const int cg8 = idx / m, j = idx - cg8 * m, c0 = 4 * cg8;
      if (j >= k0 && j < k0 + kb) {
#pragma unroll
        for (int c = 0; c < 4; ++c) Rk[(size_t)(c0 + c) * m + j] = s_inv[c0 + c][j - k0];
      } else {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 8
        for (int q = 0; q < GJ_NB; ++q) {
          const double v = q < kb ? __ldcg(A + (size_t)(k0 + q) * m + j) : 0.0;
          s0 += s_inv[c0][q] * v; s1 += s_inv[c0 + 1][q] * v; s2 += s_inv[c0 + 2][q] * v; s3 += s_inv[c0 + 3][q] * v;
        }
        Rk[(size_t)c0 * m + j] = s0; Rk[(size_t)(c0 + 1) * m + j] = s1; Rk[(size_t)(c0 + 2) * m + j] = s2; Rk[(size_t)(c0 + 3) * m + j] = s3;
      }
    }
    for (int64_t idx = gtid; idx < (int64_t)m * GJ_NB; idx += gsize) {
      const int i = (int)(idx >> 5), c = (int)(idx & 31);
      Ck[idx] = c < kb ? __ldcg(A + (size_t)i * m + k0 + c) : 0.0;
    }
    GJ_STAMP(2);
    grid.sync();
    GJ_STAMP(3);
    // ---- (B) rank-32 update of every tile
    for (int t = blockIdx.x; t < nt * nt; t += gridDim.x) {
      const int i0 = (t / nt) * 64, j0 = (t % nt) * 64;
      for (int idx = tid; idx < 64 * GJ_NB; idx += 256) { const int r = idx >> 5, c = idx & 31; s_c[r][c] = (i0 + r < m) ? __ldcg(Ck + (size_t)(i0 + r) * GJ_NB + c) : 0.0; }
      for (int idx = tid; idx < GJ_NB * 64; idx += 256) { const int c = idx >> 6, j = idx & 63; s_r[c][j] = (j0 + j < m) ? __ldcg(Rk + (size_t)c * m + j0 + j) : 0.0; }
      __syncthreads();
      const int tr = tid >> 4, tc = tid & 15;
      double acc[4][4], old[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {         // the tile's current values are in flight while the product is formed
          const int i = i0 + tr + 16 * a, j = j0 + tc + 16 * b;
          acc[a][b] = 0.0;
          old[a][b] = (i < m && j < m) ? __ldcg(A + (size_t)i * m + j) : 0.0;
        }
#pragma unroll 8
      for (int c = 0; c < GJ_NB; ++c) {
        double ca[4], rb[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) ca[a] = s_c[tr + 16 * a][c];
#pragma unroll
        for (int b = 0; b < 4; ++b) rb[b] = s_r[c][tc + 16 * b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] += ca[a] * rb[b];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = i0 + tr + 16 * a;
        if (i >= m) continue;
        const bool prow = i >= k0 && i < k0 + kb;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int j = j0 + tc + 16 * b;
          if (j >= m) continue;
          const bool pcol = j >= k0 && j < k0 + kb;
          double v;
          if (prow) v = s_r[i - k0][tc + 16 * b];
          else v = (pcol ? 0.0 : old[a][b]) - acc[a][b];
          __stcg(A + (size_t)i * m + j, v);
        }
      }
      __syncthreads();
    }
    GJ_STAMP(4);
    grid.sync();
    GJ_STAMP(5);
  }
#undef GJ_STAMP
}


// ---- small reduced systems (local BA): direct solve in ONE CTA ------------------------------------------
// Up to 26 images (156 unknowns) the whole reduced camera system fits in shared memory as a dense matrix.  A local-BA sized
// solve took 30-90 PCG iterations of 4-9 us each (latency of the vector exchange through L2); the dense Cholesky below takes
// tens of microseconds and is what the reference does anyway (SPARSE_SCHUR = Cholesky, bundle_adjustment.cc:555).
// Block pass: stored 6 x 6 blocks -> dense lower triangle; right-looking Cholesky with all threads (2 barriers per column);
// forward / backward substitution by one warp (no block barriers).
constexpr int DENSE_MAX_IMG = 26;
__global__ void __launch_bounds__(512, 1) k_dense_chol_solve(int n_img, int64_t nblk, const int* __restrict__ blk_a, const int* __restrict__ blk_b,
                                                             const double* __restrict__ S, const double* __restrict__ rhs, double* __restrict__ x,
                                                             int* __restrict__ fail) {
  extern __shared__ double dsm[];
  const int n = 6 * n_img, ld = n + 1;
  double* L = dsm;                 // [n][ld], lower triangle used
  double* v = dsm + (size_t)n * ld;   // [n]
  __shared__ int bad;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) bad = 0;
  for (int i = tid; i < n * ld; i += nt) L[i] = 0.0;
  __syncthreads();
  for (int64_t e = tid; e < nblk * 36; e += nt) {
    const int64_t b = e / 36; const int k = (int)(e - 36 * b), r = k / 6, c = k - 6 * r;
    const int ia = b < n_img ? (int)b : blk_a[b - n_img], ib = b < n_img ? (int)b : blk_b[b - n_img];      // ia <= ib: block (ia, ib) of the upper triangle
    const double val = S[36 * (size_t)b + k];
    const int gi = 6 * ia + r, gj = 6 * ib + c;
    if (gi >= gj) L[gi * ld + gj] = val;            // diagonal blocks: their lower half
    if (ia != ib) L[gj * ld + gi] = val;            // off-diagonal block, transposed into the lower triangle
  }
  for (int i = tid; i < n; i += nt) v[i] = rhs[i];
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    const double d = L[j * ld + j];
    if (!(d > 0.0) || !isfinite(d)) { if (tid == 0) bad = 1; break; }      // uniform: every thread reads the same value
    const double l = sqrt(d), il = 1.0 / l;
    __syncthreads();
    for (int i = j + tid; i < n; i += nt) L[i * ld + j] = (i == j) ? l : L[i * ld + j] * il;
    __syncthreads();
    // trailing update of the lower triangle: A[i][k] -= L[i][j] L[k][j], j < k <= i
    const int m = n - j - 1;
    for (int a = tid >> 5; a < m; a += (nt >> 5)) {       // warp per row, lanes along the row: no integer division, conflict-free
      const int i = j + 1 + a;
      const double lij = L[i * ld + j];
      for (int c = tid & 31; c <= a; c += 32) L[i * ld + j + 1 + c] -= lij * L[(j + 1 + c) * ld + j];
    }
    __syncthreads();
  }
  __syncthreads();
  if (bad) { if (tid == 0) *fail = 1; for (int i = tid; i < n; i += nt) x[i] = 0.0; return; }
  if (tid < 32) {                   // L y = b, then L' x = y; lane-strided rows, the pivot value is broadcast through shared memory
    for (int j = 0; j < n; ++j) {
      if (tid == 0) v[j] = v[j] / L[j * ld + j];
      __syncwarp();
      const double yj = v[j];
      for (int i = j + 1 + tid; i < n; i += 32) v[i] -= L[i * ld + j] * yj;
      __syncwarp();
    }
    for (int j = n - 1; j >= 0; --j) {
      if (tid == 0) v[j] = v[j] / L[j * ld + j];
      __syncwarp();
      const double xj = v[j];
      for (int i = tid; i < j; i += 32) v[i] -= L[j * ld + i] * xj;
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) x[i] = v[i];
}

}  // namespace mm
