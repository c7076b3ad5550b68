// match_tc.cu — tensor-core candidate selection for match_brute_force (K5).
//
// Replaces the two cv::BFMatcher::knnMatch(k=2) distance passes of mavmap/mavmap
// src/base2d/feature.cc:71-72 (OpenCV batchDistance): the N x M x K distance matrix is the one
// dense contraction of the hot path, so it runs on the 5th-generation tensor cores:
//
//   prep      each descriptor row is written twice, TF32-rounded (cvt.rna), K padded to K' = 32-multiple:
//               A'[i] = [ a_i            | 1      1      h1  h2 | 0.. ]   h1+h2 = |a_i|^2 + 1 (tf32 split)
//               B'[j] = [ -2 b_j         | g1     g2     1   1  | 0.. ]   g1+g2 = |b_j|^2     (tf32 split)
//             so that  A'[i] . B'[j] = 1 + |a_i - b_j|^2  up to the TF32 rounding of the cross term.
//   k_match_tc  persistent, warp-specialised: warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle; the 128-row A tile
//             of a work item stays resident, B tiles run through a 3-stage mbarrier ring), warp 1 = tcgen05.mma issuer
//             (kind::tf32, M=128, N=256, accumulators double-buffered in all 512 TMEM columns), warps 2-9 = epilogue (two
//             per TMEM lane quarter, each half of the columns, tcgen05.ld.32x32b.x32, one thread per query row; the
//             epilogue of tile t overlaps the MMAs of tile t+1).  Two passes:
//               <0> "bound":   a quarter of the column tiles; per row the minima of eight disjoint column groups (3-input
//                              minimum instructions); the second smallest of the eight bounds the row's true second-nearest value
//               <1> "collect": all tiles; every column with value <= bound + 2 eps sets its bit in the row's hit mask (one
//                              32-bit word per row and 32-column chunk, stored coalesced, no warp collective);
//                              eps = the rigorous TF32 error bound 2^-9 |a||b|
//   k_rerank  expands the row's hit masks into a candidate list (global scratch, slot-major), then exact fp64-accumulated
//             distances (same arithmetic as match.cu / the oracle) of the listed columns -> exact top-2 by (distance, index);
//             a row with more than TC_HCAP candidates is re-scanned exactly (k_rescan).
// The result is bit-identical to the exact SIMT path; the tensor cores only prune.
#include <cuda.h>
#include <float.h>
#include <vector>
#include <algorithm>
#include <mutex>
#include <unordered_map>
#include "common.cuh"
#include "match.cuh"

namespace mm {

constexpr int TC_M = 128, TC_N = 256, TC_KB = 32;           // tile rows, tile cols, K elements per stage (128 B)
constexpr int TC_STAGES = 3;                                  // B-operand ring (32 KB per stage)
constexpr int TC_MAX_KB = 6;                                  // A tile stays resident for a whole item: K' <= 192
constexpr int TC_A_BYTES = TC_M * TC_KB * 4, TC_B_BYTES = TC_N * TC_KB * 4;
constexpr int TC_HCAP = 64;                                   // candidates a row may have before it is re-scanned exactly
constexpr int TC_BOUND_DIV = 4;                               // the bound pass looks at 1/TC_BOUND_DIV of the column tiles
constexpr int TC_THREADS = 320;                               // warp0 TMA, warp1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr uint32_t SPIN_LIMIT = 1u << 24;                     // watchdog: trap instead of hanging the GPU

// one 128-row block of one direction of one pair; mask_off: first word of the block's hit masks (collect pass): word
// [mask_off + chunk * 128 + row] has bit e set iff column 32 chunk + e of the B image is a candidate of the block's row `row`
struct TcItem { int rowA0, nA, rowB0, nB; int64_t out_off; int64_t mask_off; float bmax; int pad; };

// ---------------------------------------------------------------- PTX wrappers (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) { if (++spins > SPIN_LIMIT) __trap(); }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(map)) : "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle: rows are 128 B apart, 8-row atoms
// are 1024 B apart (SBO); LBO is unused for swizzled K-major; descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // start address [0,14)
  d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                                  // version
  d |= (uint64_t)2 << 61;                                  // layout: SWIZZLE_128B
  return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- prep: fp32 descriptors -> TF32 operand rows
__device__ __forceinline__ float to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }

__global__ void k_tc_prep(const float* __restrict__ desc, int64_t rows, int K, int Kp, float* __restrict__ PA, float* __restrict__ PB,
                          float* __restrict__ norm_out /* [rows] |x|^2 in fp32 (rounded from fp64) */) {
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = desc + row * K;
  double n2 = 0.0;
  for (int k = lane; k < Kp; k += 32) {
    float a = 0.f, b = 0.f;
    if (k < K) { const float v = x[k]; n2 += (double)v * (double)v; a = to_tf32(v); b = -2.0f * a; }
    PA[row * Kp + k] = a; PB[row * Kp + k] = b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  __syncwarp();                          // the zero fill of columns >= K (other lanes) precedes lane 0's writes below
  if (lane == 0) {
    const float h1 = to_tf32((float)(n2 + 1.0)), h2 = to_tf32((float)(n2 + 1.0 - (double)h1));
    const float g1 = to_tf32((float)n2), g2 = to_tf32((float)(n2 - (double)g1));
    PA[row * Kp + K] = 1.f; PA[row * Kp + K + 1] = 1.f; PA[row * Kp + K + 2] = h1; PA[row * Kp + K + 3] = h2;
    PB[row * Kp + K] = g1; PB[row * Kp + K + 1] = g2; PB[row * Kp + K + 2] = 1.f; PB[row * Kp + K + 3] = 1.f;
    norm_out[row] = (float)n2;
  }
}

// ---------------------------------------------------------------- the tensor-core kernel
// TF32 error bound of the approximate value 1 + |a-b|^2 (see k_rerank): both operands RN-rounded to TF32
__device__ __forceinline__ float tc_eps(float na, float bmax) {
  return 0.001953125f * 1.02f * sqrtf(na) * bmax + 1e-5f * (na + bmax * bmax + 1.0f);
}

// MODE 0 ("bound"): only the first ~1/4 of the column tiles; every row keeps the two smallest approximate values seen
//          -> bound[row] = 2nd smallest over that column subset, an upper bound of the row's true 2nd-nearest value.
// MODE 1 ("collect"): all column tiles; the epilogue only compares against the per-row threshold
//          T = bound + 2 eps (+ rounding slack) and sets the bits of the (rare) columns below it in the row's hit masks.
//          Every column that can be in the exact top-2 is below T, so the hit masks are complete by construction.
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) k_match_tc(
    const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
    const TcItem* __restrict__ items, int n_items, int num_kb, const float* __restrict__ norms,
    float* __restrict__ bound, uint32_t* __restrict__ masks) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base, smem_b = smem_base + TC_MAX_KB * TC_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + TC_MAX_KB * TC_A_BYTES + TC_STAGES * TC_B_BYTES);
  // bars: [0..S) b_full, [S..2S) b_empty, 2S a_full, 2S+1 a_empty, [2S+2..2S+4) tmem_full, [2S+4..2S+6) tmem_empty
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 6);
  float* merge_f = reinterpret_cast<float*>(bars + 2 * TC_STAGES + 8);           // [128][4]
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + TC_STAGES);
  const uint32_t bar_afull = smem_u32(bars + 2 * TC_STAGES), bar_aempty = smem_u32(bars + 2 * TC_STAGES + 1);
  const uint32_t bar_tfull = smem_u32(bars + 2 * TC_STAGES + 2), bar_tempty = smem_u32(bars + 2 * TC_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB);
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_afull, 1); mbar_init(bar_aempty, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  auto tiles_of = [](int nB) { const int n = (nB + TC_N - 1) / TC_N; return MODE == 0 ? max(1, (n + TC_BOUND_DIV - 1) / TC_BOUND_DIV) : n; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, aphase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const TcItem w = items[it];
        const int n_tiles = tiles_of(w.nB);
        mbar_wait(bar_aempty, aphase ^ 1);                        // MMAs of the previous item have retired: A region is free
        mbar_expect_tx(bar_afull, (uint32_t)num_kb * TC_A_BYTES);
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(smem_a + kb * TC_A_BYTES, &mapA, bar_afull, kb * TC_KB, w.rowA0);
        aphase ^= 1;
        for (int nt = 0; nt < n_tiles; ++nt)
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            mbar_expect_tx(bar_full + 8 * stage, TC_B_BYTES);
            tma_load_2d(smem_b + stage * TC_B_BYTES, &mapB, bar_full + 8 * stage, kb * TC_KB, w.rowB0 + nt * TC_N);
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(TC_M, TC_N);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, aphase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const TcItem w = items[it];
        const int n_tiles = tiles_of(w.nB);
        mbar_wait(bar_afull, aphase); aphase ^= 1;
        tc_fence_after();
        for (int nt = 0; nt < n_tiles; ++nt) {
          mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);          // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TC_N;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint64_t da = umma_desc_sw128(smem_a + kb * TC_A_BYTES), db = umma_desc_sw128(smem_b + stage * TC_B_BYTES);
#pragma unroll
            for (int k = 0; k < TC_KB / 8; ++k)                    // UMMA_K = 8 tf32 = 32 bytes -> +2 in the (addr >> 4) field
              umma_tf32(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            umma_commit(bar_empty + 8 * stage);                    // frees the B stage when these MMAs retire
            if (kb == num_kb - 1) umma_commit(bar_tfull + 8 * acc);  // accumulator ready for the epilogue
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        umma_commit(bar_aempty);                                   // A region reusable once every MMA of this item retired
      }
    }
  } else {
    // ===================== epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row_in_tile = q * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const TcItem w = items[it];
      const int n_tiles = tiles_of(w.nB);
      const bool row_ok = row_in_tile < w.nA;
      // MODE 0: minima of four disjoint column groups (chunk index within the tile half).  The 2nd smallest of the row's
      // eight group minima (both halves) is >= the row's 2nd smallest value: a valid bound at one FMNMX3 per two columns.
      float g0 = FLT_MAX, g1 = FLT_MAX, g2 = FLT_MAX, g3 = FLT_MAX;
      float T = 0.f;                                                // MODE 1: the row's threshold
      uint32_t* my_masks = nullptr;
      if (MODE == 1) {
        const int64_t r = w.out_off + (row_ok ? row_in_tile : 0);
        const float bd = bound[r];
        const float na = norms[w.rowA0 + (row_ok ? row_in_tile : 0)];
        T = row_ok ? (bd + 2.0f * tc_eps(na, w.bmax)) * (1.0f + 4e-6f) : -1.f;    // FLT_MAX bound stays +inf
        my_masks = masks + w.mask_off + row_in_tile;
      }
      for (int nt = 0; nt < n_tiles; ++nt) {
        mbar_wait(bar_tfull + 8 * acc, acc_phase);
        tc_fence_after();
        const int col0 = nt * TC_N + half * (TC_N / 2);
        const int ncols = min(TC_N / 2, w.nB - col0);               // may be <= 0 for the second half of the last tile
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TC_N + half * (TC_N / 2);
        uint32_t va[32], vb[32];
        if (ncols > 0) tmem_ld32(tbase, va);
#pragma unroll
        for (int ch = 0; ch < TC_N / 64; ++ch) {
          if (ch * 32 >= ncols) break;                              // warp-uniform
          uint32_t (&v)[32] = (ch & 1) ? vb : va;
          tmem_ld_wait();
          if ((ch + 1) * 32 < ncols && ch + 1 < TC_N / 64) tmem_ld32(tbase + (ch + 1) * 32, (ch & 1) ? va : vb);
          const int lim = min(32, ncols - ch * 32);
          if (lim < 32) {                                           // ragged last chunk of an image: columns past the end read as FLT_MAX
#pragma unroll
            for (int e = 0; e < 32; ++e) if (e >= lim) v[e] = __float_as_uint(FLT_MAX);
          }
          if (MODE == 0) {
            // minima of the four 8-column groups of the chunk: 4 FMNMX3-class instructions each
            float s[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float t1 = fminf(fminf(__uint_as_float(v[8 * k]), __uint_as_float(v[8 * k + 1])), __uint_as_float(v[8 * k + 2]));
              const float t2 = fminf(fminf(__uint_as_float(v[8 * k + 3]), __uint_as_float(v[8 * k + 4])), __uint_as_float(v[8 * k + 5]));
              s[k] = fminf(fminf(fminf(t1, t2), __uint_as_float(v[8 * k + 6])), __uint_as_float(v[8 * k + 7]));
            }
            const float m = fminf(fminf(fminf(s[0], s[1]), s[2]), s[3]);
            if (ch == 0) g0 = fminf(g0, m); else if (ch == 1) g1 = fminf(g1, m); else if (ch == 2) g2 = fminf(g2, m); else g3 = fminf(g3, m);
          } else {
            // A 32-bit hit mask per (row, chunk of 32 columns), stored as it is: consecutive lanes = consecutive rows write one
            // 128-byte line, no warp collective, no walk over the set bits - the re-rank kernel expands the masks.
            // (Forms that did the expansion here, measured on B200 at 60 pairs of 5000 x 5000 x 64: one warp OR + a warp-uniform
            // walk over the set bits 1.17 - 1.26 ms; a 3-input-min tree with a voted slow path 1.79 ms; predicated appends 2.10 ms.)
            uint32_t mask = 0;
#pragma unroll
            for (int e = 0; e < 32; ++e) mask |= (__uint_as_float(v[e]) <= T) ? (1u << e) : 0u;
            if (lim < 32) mask &= (1u << lim) - 1u;
            my_masks[(size_t)(nt * (TC_N / 32) + half * (TC_N / 64) + ch) * TC_M] = mask;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (MODE == 0) {
        // merge the two column halves of each row (half 1 -> smem -> half 0): 2nd smallest of the eight group minima
        if (half == 1) { merge_f[row_in_tile * 4] = g0; merge_f[row_in_tile * 4 + 1] = g1; merge_f[row_in_tile * 4 + 2] = g2; merge_f[row_in_tile * 4 + 3] = g3; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 0) {
          float m1 = FLT_MAX, m2 = FLT_MAX;
          const float x[8] = { g0, g1, g2, g3, merge_f[row_in_tile * 4], merge_f[row_in_tile * 4 + 1], merge_f[row_in_tile * 4 + 2], merge_f[row_in_tile * 4 + 3] };
#pragma unroll
          for (int e = 0; e < 8; ++e) { m2 = fminf(m2, fmaxf(m1, x[e])); m1 = fminf(m1, x[e]); }
          if (row_ok) bound[w.out_off + row_in_tile] = m2;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------- exact re-rank of the candidates
__device__ __forceinline__ double exact_d2(const float* __restrict__ a, const float* __restrict__ b, int K) {
  double s = 0.0;
  for (int k = 0; k < K; ++k) { const double t = (double)a[k] - (double)b[k]; s = __dadd_rn(s, __dmul_rn(t, t)); }
  return s;
}
__device__ __forceinline__ void top2_insert_f(float d, int j, float& b0, int& i0, float& b1, int& i1) {
  // (dist, index) lexicographic order == OpenCV's strict-'<' insertion in ascending index
  if (d < b1 || (d == b1 && j < i1) || i1 < 0) {
    if (i0 < 0 || d < b0 || (d == b0 && j < i0)) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
    else { b1 = d; i1 = j; }
  }
}

// mask_off / item_stride: the hit masks of the job's row block m start at mask_off + m * item_stride (see TcItem)
struct RerankJob { int rowA0, nA, rowB0, nB; int64_t out_off; int64_t knn_off; int64_t mask_off; int item_stride; float bmax; };

// thread per query row: expands the row's hit masks (one word per chunk of 32 columns; the 128 threads of a block read one line
// per chunk) and computes the exact fp64-accumulated distances (strictly ascending k, separate multiply and add - the same
// arithmetic as match.cu and the oracle) of the hit columns, four independent chains at a time, then the exact top-2 by
// (distance, index).  A row with more than TC_HCAP candidates is re-scanned exactly by a whole CTA (k_rescan).
__global__ void __launch_bounds__(128) k_rerank(const RerankJob* __restrict__ jobs, const float* __restrict__ desc, int K,
                         const uint32_t* __restrict__ masks, int* __restrict__ cand_g, Knn2* __restrict__ knn12, Knn2* __restrict__ knn21,
                         int n12 /* jobs [0, n12) write knn12, the rest knn21 */, int* __restrict__ flagged, int* __restrict__ n_flagged, int flag_cap) {
  const RerankJob job = jobs[blockIdx.y];
  Knn2* __restrict__ knn = (int)blockIdx.y < n12 ? knn12 : knn21;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= job.nA) return;
  const float* a = desc + (size_t)(job.rowA0 + i) * K;
  const uint32_t* mrow = masks + job.mask_off + (size_t)(i >> 7) * job.item_stride + (i & 127);
  float b0 = FLT_MAX, b1 = FLT_MAX; int i0 = -1, i1 = -1;
  // phase 1: the set bits of the row's masks -> its candidate list in shared memory (slot-major: conflict-free).  Light and
  // divergent; the heavy part below then runs in lockstep over the lanes (expanding and computing in one loop made every lane
  // wait for the distance batches of all the others: 1.54 ms instead of 0.3 ms per direction).
  // (the list lives in global scratch, slot-major per block: as 32 KB of shared memory it took the L1 capacity that the sixteen
  // 16-byte reads of every candidate row rely on, and the kernel ran at 0.52 instead of 0.3 ms)
  int* cand = cand_g + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (TC_HCAP * 128);
  int n = 0;
  // (a one-byte-per-half-tile summary of the non-zero words, written by the collect pass and read first here, did not pay: the
  // scan is not what this kernel waits for, and the collect pass lost 1.4 points of tensor-pipe activity to the extra stores)
  const int n_chunks = (job.nB + 31) >> 5;
  for (int c = 0; c < n_chunks; c += 20) {                     // twenty independent loads in flight per thread: the scan is a stream, not a chain
    uint32_t wv[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) wv[k] = (c + k < n_chunks) ? __ldg(mrow + (size_t)(c + k) * TC_M) : 0u;
#pragma unroll
    for (int k = 0; k < 20; ++k) {
      uint32_t wbits = wv[k];
      while (wbits) { const int e = __ffs(wbits) - 1; wbits &= wbits - 1; if (n < TC_HCAP) cand[n * 128 + threadIdx.x] = 32 * (c + k) + e; ++n; }
    }
  }
  const bool overflow = n > TC_HCAP;
  const int nc = min(n, TC_HCAP);
  // phase 2: exact distances, four independent chains at a time.  The rows of the next four candidates are prefetched into
  // L1 while the current four are summed: every step of the serial k loop was a scattered 16-byte read at L2 latency, which
  // is what a call with few pairs (40 blocks per direction, nothing to overlap with) spent its time on.
  const int row_bytes = K * (int)sizeof(float);
  auto prefetch_rows = [&](int c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) if (c + k < nc) {
      const char* r = reinterpret_cast<const char*>(desc + (size_t)(job.rowB0 + cand[(c + k) * 128 + threadIdx.x]) * K);
      for (int o = 0; o < row_bytes; o += 128) asm volatile("prefetch.global.L1 [%0];" :: "l"(r + o));
    }
  };
  prefetch_rows(0);
  for (int c0 = 0; c0 < nc; c0 += 4) {
    int col[4]; const float* bp[4]; double e2[4];
    prefetch_rows(c0 + 4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      col[k] = (c0 + k < nc) ? cand[(c0 + k) * 128 + threadIdx.x] : -1;
      bp[k] = desc + (size_t)(job.rowB0 + (col[k] >= 0 ? col[k] : 0)) * K;
      e2[k] = 0.0;
    }
    for (int k0 = 0; k0 < K; k0 += 4) {                           // K % 4 == 0 on this path
      const float4 av = *reinterpret_cast<const float4*>(a + k0);
      float4 bv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) bv[k] = *reinterpret_cast<const float4*>(bp[k] + k0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double t;
        t = (double)av.x - (double)bv[k].x; e2[k] = __dadd_rn(e2[k], __dmul_rn(t, t));
        t = (double)av.y - (double)bv[k].y; e2[k] = __dadd_rn(e2[k], __dmul_rn(t, t));
        t = (double)av.z - (double)bv[k].z; e2[k] = __dadd_rn(e2[k], __dmul_rn(t, t));
        t = (double)av.w - (double)bv[k].w; e2[k] = __dadd_rn(e2[k], __dmul_rn(t, t));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) if (col[k] >= 0) top2_insert_f(__fsqrt_rn((float)e2[k]), col[k], b0, i0, b1, i1);
  }
  Knn2 out; out.d0 = b0; out.d1 = b1; out.i0 = i0; out.i1 = i1;
  knn[job.knn_off + i] = out;
  if (overflow) { const int slot = atomicAdd(n_flagged, 1); if (slot < flag_cap) { flagged[2 * slot] = blockIdx.y; flagged[2 * slot + 1] = i; } }
}

// exact full scan of one flagged row per warp (same arithmetic and tie rule as the SIMT path)
__global__ void __launch_bounds__(256) k_rescan(const RerankJob* __restrict__ jobs, const float* __restrict__ desc, int K,
                         const int* __restrict__ flagged, const int* __restrict__ n_flagged, int flag_cap, Knn2* __restrict__ knn12, Knn2* __restrict__ knn21, int n12) {
  // one CTA per flagged row (a warp per row took 0.1 ms per row at 5000 x 64: the fp64 chain of a distance is serial by design)
  __shared__ float sb0[8], sb1[8]; __shared__ int si0[8], si1[8];
  const int nf = min(*n_flagged, flag_cap);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int f = blockIdx.x; f < nf; f += gridDim.x) {
    const RerankJob job = jobs[flagged[2 * f]];
    Knn2* __restrict__ knn = flagged[2 * f] < n12 ? knn12 : knn21;
    const int i = flagged[2 * f + 1];
    const float* a = desc + (size_t)(job.rowA0 + i) * K;
    float b0 = FLT_MAX, b1 = FLT_MAX; int i0 = -1, i1 = -1;
    for (int j = threadIdx.x; j < job.nB; j += blockDim.x) {
      // same arithmetic as exact_d2 (ascending k, separate multiply and add); the row is fetched 16 floats ahead of the serial
      // fp64 chain so that the chain, not the loads, sets the pace (K % 4 == 0 on this path)
      const float* b = desc + (size_t)(job.rowB0 + j) * K;
      double s2 = 0.0;
      int k0 = 0;
      for (; k0 + 16 <= K; k0 += 16) {
        float4 av[4], bv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { av[q] = *reinterpret_cast<const float4*>(a + k0 + 4 * q); bv[q] = *reinterpret_cast<const float4*>(b + k0 + 4 * q); }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          double t;
          t = (double)av[q].x - (double)bv[q].x; s2 = __dadd_rn(s2, __dmul_rn(t, t));
          t = (double)av[q].y - (double)bv[q].y; s2 = __dadd_rn(s2, __dmul_rn(t, t));
          t = (double)av[q].z - (double)bv[q].z; s2 = __dadd_rn(s2, __dmul_rn(t, t));
          t = (double)av[q].w - (double)bv[q].w; s2 = __dadd_rn(s2, __dmul_rn(t, t));
        }
      }
      for (; k0 < K; ++k0) { const double t = (double)a[k0] - (double)b[k0]; s2 = __dadd_rn(s2, __dmul_rn(t, t)); }
      top2_insert_f(__fsqrt_rn((float)s2), j, b0, i0, b1, i1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob0 = __shfl_xor_sync(0xffffffffu, b0, o), ob1 = __shfl_xor_sync(0xffffffffu, b1, o);
      const int oi0 = __shfl_xor_sync(0xffffffffu, i0, o), oi1 = __shfl_xor_sync(0xffffffffu, i1, o);
      if (oi0 >= 0) top2_insert_f(ob0, oi0, b0, i0, b1, i1);
      if (oi1 >= 0) top2_insert_f(ob1, oi1, b0, i0, b1, i1);
    }
    if (lane == 0) { sb0[wib] = b0; sb1[wib] = b1; si0[wib] = i0; si1[wib] = i1; }
    __syncthreads();
    if (wib == 0) {
      const bool have = lane < (int)(blockDim.x >> 5);
      b0 = have ? sb0[lane] : FLT_MAX; b1 = have ? sb1[lane] : FLT_MAX; i0 = have ? si0[lane] : -1; i1 = have ? si1[lane] : -1;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const float ob0 = __shfl_xor_sync(0xffffffffu, b0, o), ob1 = __shfl_xor_sync(0xffffffffu, b1, o);
        const int oi0 = __shfl_xor_sync(0xffffffffu, i0, o), oi1 = __shfl_xor_sync(0xffffffffu, i1, o);
        if (oi0 >= 0) top2_insert_f(ob0, oi0, b0, i0, b1, i1);
        if (oi1 >= 0) top2_insert_f(ob1, oi1, b0, i0, b1, i1);
      }
      if (lane == 0) { Knn2 r; r.d0 = b0; r.d1 = b1; r.i0 = i0; r.i1 = i1; knn[job.knn_off + i] = r; }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- host side
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr; static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr; cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    else cudaGetLastError();
  });
  return fn;
}

struct Prepared {           // TF32 operand copies of one descriptor array
  const float* desc = nullptr; int64_t rows = 0; int K = 0, Kp = 0; bool valid = false; int64_t cap_rows = 0;
  DevBuf<float> PA, PB, norms; CUtensorMap mapA, mapB; std::vector<float> h_norm_max;   // per call computed lazily
  std::vector<float> h_norms;
  std::unordered_map<uint64_t, float> bmax_cache;     // (first row, rows) of an image -> largest |b|^2
};
std::mutex g_prep_mu;
std::vector<Prepared*> g_prepared;

int make_map(CUtensorMap* map, float* base, int64_t rows, int Kp, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return MM_ERR_UNSUPPORTED; }
  cuuint64_t gdim[2] = { (cuuint64_t)Kp, (cuuint64_t)rows };
  cuuint64_t gstride[1] = { (cuuint64_t)Kp * 4 };
  cuuint32_t box[2] = { (cuuint32_t)TC_KB, (cuuint32_t)box_rows };
  cuuint32_t estr[2] = { 1, 1 };
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return MM_ERR_CUDA; }
  return MM_OK;
}

int fill_prepared(Prepared* p, int64_t rows, cudaStream_t st) {
  const int64_t prow = rows + TC_N;       // slack rows so every TMA box starts in bounds
  p->rows = rows;
  MM_CUDA(cudaMemsetAsync(p->PA.p + (size_t)rows * p->Kp, 0, sizeof(float) * (size_t)TC_N * p->Kp, st));
  MM_CUDA(cudaMemsetAsync(p->PB.p + (size_t)rows * p->Kp, 0, sizeof(float) * (size_t)TC_N * p->Kp, st));
  k_tc_prep<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(p->desc, rows, p->K, p->Kp, p->PA.p, p->PB.p, p->norms.p);
  MM_LAUNCH_CHECK();
  p->h_norms.resize((size_t)rows); p->bmax_cache.clear();
  MM_CUDA(cudaMemcpyAsync(p->h_norms.data(), p->norms.p, sizeof(float) * (size_t)rows, cudaMemcpyDeviceToHost, st));
  MM_CUDA(cudaStreamSynchronize(st));
  int rc = make_map(&p->mapA, p->PA.p, prow, p->Kp, TC_M); if (rc) return rc;
  rc = make_map(&p->mapB, p->PB.p, prow, p->Kp, TC_N); if (rc) return rc;
  p->valid = true;
  return MM_OK;
}

int get_prepared(const float* desc, int64_t rows, int K, cudaStream_t st, Prepared** out) {
  std::lock_guard<std::mutex> lk(g_prep_mu);
  for (Prepared* p : g_prepared) if (p->desc == desc && p->K == K) {
    if (p->valid && p->rows >= rows) { *out = p; return MM_OK; }
    if (p->cap_rows >= rows) { int rc = fill_prepared(p, rows, st); if (rc) return rc; *out = p; return MM_OK; }
  }
  for (size_t i = 0; i < g_prepared.size(); ++i) if (g_prepared[i]->desc == desc) { delete g_prepared[i]; g_prepared.erase(g_prepared.begin() + i); --i; }
  Prepared* p = new Prepared(); p->desc = desc; p->K = K; p->Kp = (K + 4 + TC_KB - 1) / TC_KB * TC_KB;
  p->cap_rows = rows + rows / 2;
  const int64_t prow = p->cap_rows + TC_N;
  if (p->PA.alloc((size_t)prow * p->Kp) != cudaSuccess || p->PB.alloc((size_t)prow * p->Kp) != cudaSuccess || p->norms.alloc((size_t)p->cap_rows + 1) != cudaSuccess) {
    delete p; cudaGetLastError(); set_error("cudaMalloc failed for the TF32 operand copies"); return MM_ERR_ALLOC; }
  int rc = fill_prepared(p, rows, st); if (rc) { delete p; return rc; }
  g_prepared.push_back(p);
  *out = p;
  return MM_OK;
}

struct TcScratch { DevBuf<TcItem> items; DevBuf<RerankJob> jobs; DevBuf<float> bound; DevBuf<uint32_t> masks; DevBuf<int> flagged, n_flagged, lists; size_t items_cap = 0, jobs_cap = 0, cand_cap = 0, mask_cap = 0, lists_cap = 0; };
TcScratch g_scr;
std::atomic<uint64_t> g_tc_rows{0}, g_tc_flagged{0};

}  // namespace

void match_tc_release(const float* desc) {
  std::lock_guard<std::mutex> lk(g_prep_mu);
  for (size_t i = 0; i < g_prepared.size(); ++i) if (g_prepared[i]->desc == desc) { delete g_prepared[i]; g_prepared.erase(g_prepared.begin() + i); --i; }
}
void match_tc_invalidate(const float* desc) {
  std::lock_guard<std::mutex> lk(g_prep_mu);
  for (Prepared* p : g_prepared) if (p->desc == desc) p->valid = false;
}
int match_tc_refresh_rows(const float* desc, int K, const int64_t* r0, const int64_t* n, int count, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prep_mu);
  for (Prepared* p : g_prepared) if (p->desc == desc && p->K == K && p->valid) {
    bool any = false;
    for (int i = 0; i < count; ++i) {
      if (n[i] <= 0) continue;
      if (r0[i] < 0 || r0[i] + n[i] > p->rows) { p->valid = false; return MM_OK; }      // outside what was prepared: redo all of it on next use
      k_tc_prep<<<(unsigned)((n[i] + 7) / 8), 256, 0, st>>>(desc + r0[i] * K, n[i], K, p->Kp, p->PA.p + r0[i] * p->Kp, p->PB.p + r0[i] * p->Kp, p->norms.p + r0[i]);
      MM_LAUNCH_CHECK();
      MM_CUDA(cudaMemcpyAsync(p->h_norms.data() + r0[i], p->norms.p + r0[i], sizeof(float) * (size_t)n[i], cudaMemcpyDeviceToHost, st));
      any = true;
    }
    if (any) { MM_CUDA(cudaStreamSynchronize(st)); p->bmax_cache.clear(); }
    return MM_OK;
  }
  return MM_OK;
}
void match_tc_stats(uint64_t* rows, uint64_t* flagged) { *rows = g_tc_rows.load(); *flagged = g_tc_flagged.load(); }

// one range [p0, p1) of the pairs of a call: bound pass, collect pass, re-rank (+ re-scan of overflowing rows)
static int tc_run_range(Prepared* P, const float* desc, int K, const int64_t* offs, const int32_t* ia, const int32_t* ib,
                        const PairJob* jobs_host, int p0, int p1, Knn2* knn12, Knn2* knn21, cudaStream_t st) {
  // work items: (pair, direction, 128-row block); knn12 and knn21 live in different arrays
  std::vector<TcItem> items; std::vector<RerankJob> rjobs;
  int64_t cand_rows = 0, mask_words = 0;
  for (int p = p0; p < p1; ++p) {
    const PairJob& j = jobs_host[p];
    for (int dir = 0; dir < 2; ++dir) {
      const int imgA = dir == 0 ? ia[p] : ib[p], imgB = dir == 0 ? ib[p] : ia[p];
      const int nA = dir == 0 ? j.n1 : j.n2, nB = dir == 0 ? j.n2 : j.n1;
      if (nA == 0) continue;
      float bmax = 0.f;
      { const uint64_t key = ((uint64_t)offs[imgB] << 24) ^ (uint64_t)nB;
        auto hit = P->bmax_cache.find(key);
        if (hit != P->bmax_cache.end()) bmax = hit->second;
        else { for (int64_t r = offs[imgB]; r < offs[imgB] + nB; ++r) bmax = std::max(bmax, P->h_norms[(size_t)r]); P->bmax_cache[key] = bmax; } }
      RerankJob rj; rj.rowA0 = (int)offs[imgA]; rj.nA = nA; rj.rowB0 = (int)offs[imgB]; rj.nB = nB; rj.out_off = cand_rows;
      rj.knn_off = dir == 0 ? j.knn12_off : -(j.knn21_off + 1); rj.bmax = std::sqrt(bmax);
      // hit masks: per 128-row block one word per (chunk of 32 columns, row), chunks padded to whole column tiles
      rj.item_stride = (nB + TC_N - 1) / TC_N * (TC_N / 32) * TC_M; rj.mask_off = mask_words;
      rjobs.push_back(rj);
      if (nB > 0) for (int m = 0; m < nA; m += TC_M) { TcItem t; t.rowA0 = (int)offs[imgA] + m; t.nA = std::min(TC_M, nA - m); t.rowB0 = (int)offs[imgB]; t.nB = nB; t.out_off = cand_rows + m; t.mask_off = mask_words + (int64_t)(m / TC_M) * rj.item_stride; t.bmax = rj.bmax; t.pad = 0; items.push_back(t); }
      mask_words += (int64_t)((nA + TC_M - 1) / TC_M) * rj.item_stride;
      cand_rows += nA;
    }
  }
  if (items.size() > g_scr.items_cap) { MM_CUDA(g_scr.items.alloc(items.size())); g_scr.items_cap = items.size(); }
  if (rjobs.size() > g_scr.jobs_cap) { MM_CUDA(g_scr.jobs.alloc(rjobs.size())); g_scr.jobs_cap = rjobs.size(); }
  if ((size_t)cand_rows > g_scr.cand_cap) {
    const size_t c = (size_t)cand_rows + (size_t)cand_rows / 4;
    MM_CUDA(g_scr.bound.alloc(c)); MM_CUDA(g_scr.flagged.alloc(2 * c + 2)); g_scr.cand_cap = c; }
  if ((size_t)mask_words > g_scr.mask_cap) { const size_t c = (size_t)mask_words + (size_t)mask_words / 4; MM_CUDA(g_scr.masks.alloc(c)); g_scr.mask_cap = c; }
  if (!g_scr.n_flagged.p) MM_CUDA(g_scr.n_flagged.alloc(1));
  // knn21 jobs were tagged with a negative offset: split the re-rank into the two output arrays
  std::vector<RerankJob> j12, j21;
  for (auto& r : rjobs) { if (r.knn_off >= 0) j12.push_back(r); else { RerankJob t = r; t.knn_off = -r.knn_off - 1; j21.push_back(t); } }
  std::vector<RerankJob> all = j12; all.insert(all.end(), j21.begin(), j21.end());
  MM_CUDA(cudaMemcpyAsync(g_scr.jobs.p, all.data(), sizeof(RerankJob) * all.size(), cudaMemcpyHostToDevice, st));
  // (a job with nB == 0 has no chunks, a block's masks are written for every chunk the re-rank reads: nothing to clear)
  if (!items.empty()) {
    MM_CUDA(cudaMemcpyAsync(g_scr.items.p, items.data(), sizeof(TcItem) * items.size(), cudaMemcpyHostToDevice, st));
    const size_t smem = (size_t)TC_MAX_KB * TC_A_BYTES + (size_t)TC_STAGES * TC_B_BYTES + 1024 + 256 + (size_t)TC_M * 16;
    MM_CUDA(cudaFuncSetAttribute(k_match_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      // per device / context: set on every call
    MM_CUDA(cudaFuncSetAttribute(k_match_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::min((int)items.size(), num_sms());
    const int num_kb = P->Kp / TC_KB;
    k_match_tc<0><<<grid, TC_THREADS, smem, st>>>(P->mapA, P->mapB, g_scr.items.p, (int)items.size(), num_kb, P->norms.p, g_scr.bound.p, g_scr.masks.p);
    MM_LAUNCH_CHECK();
    k_match_tc<1><<<grid, TC_THREADS, smem, st>>>(P->mapA, P->mapB, g_scr.items.p, (int)items.size(), num_kb, P->norms.p, g_scr.bound.p, g_scr.masks.p);
    MM_LAUNCH_CHECK();
  }
  MM_CUDA(cudaMemsetAsync(g_scr.n_flagged.p, 0, sizeof(int), st));
  const int flag_cap = (int)cand_rows;
  int max_nA = 1; for (auto& r : all) max_nA = std::max(max_nA, r.nA);
  if (!all.empty()) {
    // both directions in one launch (the job index says which top-2 array a row belongs to): a single pair is forty blocks
    // per direction, and two launches in a row left most of the machine idle twice
    const size_t need = (size_t)((max_nA + 127) / 128) * all.size() * TC_HCAP * 128;      // candidate lists, slot-major per block
    if (need > g_scr.lists_cap) { MM_CUDA(g_scr.lists.alloc(need)); g_scr.lists_cap = need; }
    k_rerank<<<dim3((max_nA + 127) / 128, (unsigned)all.size()), 128, 0, st>>>(g_scr.jobs.p, desc, K, g_scr.masks.p, g_scr.lists.p, knn12, knn21, (int)j12.size(),
                                                                             g_scr.flagged.p, g_scr.n_flagged.p, flag_cap);
    MM_LAUNCH_CHECK();
    k_rescan<<<num_sms() * 4, 256, 0, st>>>(g_scr.jobs.p, desc, K, g_scr.flagged.p, g_scr.n_flagged.p, flag_cap, knn12, knn21, (int)j12.size());
    MM_LAUNCH_CHECK();
  }
  g_tc_rows.fetch_add((uint64_t)cand_rows);
  if (getenv("MM_MATCH_TC_STATS") != nullptr) {       // (a D2H copy into pageable memory stalls the launch queue: only on request)
    int nf = 0; MM_CUDA(cudaMemcpyAsync(&nf, g_scr.n_flagged.p, sizeof(int), cudaMemcpyDeviceToHost, st)); MM_CUDA(cudaStreamSynchronize(st));
    g_tc_flagged.fetch_add((uint64_t)nf);
    fprintf(stderr, "[match_tc] rows %lld flagged for exact rescan: %d\n", (long long)cand_rows, nf);
  }
  return MM_OK;
}

int match_tc_pairs(const float* desc, const float* xy, int K, const int64_t* offs, int64_t total_rows, const int32_t* ia, const int32_t* ib,
                   const PairJob* jobs_host, int n_pairs, double max_distance, Knn2* knn12, Knn2* knn21,
                   cudaStream_t st, bool required) {
  (void)xy;
  if (max_distance != -1.0) { if (required) set_error("the tcgen05 path does not take the keypoint-distance mask"); return MM_ERR_UNSUPPORTED; }
  if (K % 4 != 0 || K < 8 || (K + 4 + TC_KB - 1) / TC_KB > TC_MAX_KB) { if (required) set_error("tcgen05 path needs K %% 4 == 0 and K <= %d", TC_MAX_KB * TC_KB - 4); return MM_ERR_UNSUPPORTED; }
  if (getenv("MM_MATCH_NO_TC") && !required) return MM_ERR_UNSUPPORTED;

  // the whole descriptor array of the set: its row count is the offset past the last image referenced
  const int64_t rows = total_rows;
  // prepared copies are keyed by the base pointer; use the full extent the set was created with when known
  Prepared* P = nullptr;
  { std::lock_guard<std::mutex> lk(g_prep_mu);
    for (Prepared* q : g_prepared) if (q->desc == desc && q->K == K && q->valid && q->rows >= rows) { P = q; break; } }
  if (!P) { int rc = get_prepared(desc, rows, K, st, &P); if (rc) return required ? rc : MM_ERR_UNSUPPORTED; }

  // The hit masks and candidate lists are scratch per (pair, direction, row block): a call with many pairs runs as a
  // sequence of pair ranges whose scratch stays below a fixed budget (~9 MB per 5000 x 5000 pair, 3 GB by default:
  // ranges of ~330 pairs = 26 000 work items each, far more than the machine needs to be full).
  int64_t budget_words = (int64_t)768 << 20;
  if (const char* e = getenv("MM_MATCH_TC_SCRATCH_WORDS")) budget_words = std::max<int64_t>(1, atoll(e));      // (tests force several ranges)
  auto pair_words = [&](int p) {
    int64_t w = 0;
    for (int dir = 0; dir < 2; ++dir) {
      const int64_t nA = dir == 0 ? jobs_host[p].n1 : jobs_host[p].n2, nB = dir == 0 ? jobs_host[p].n2 : jobs_host[p].n1;
      w += (nA + TC_M - 1) / TC_M * ((nB + TC_N - 1) / TC_N * (TC_N / 32) * TC_M + (int64_t)TC_HCAP * 128);
    }
    return w;
  };
  for (int p0 = 0; p0 < n_pairs;) {
    int p1 = p0; int64_t words = 0;
    while (p1 < n_pairs && (p1 == p0 || words + pair_words(p1) <= budget_words)) { words += pair_words(p1); ++p1; }
    const int rc = tc_run_range(P, desc, K, offs, ia, ib, jobs_host, p0, p1, knn12, knn21, st); if (rc) return rc;
    p0 = p1;
  }
  return MM_OK;
}

}  // namespace mm
