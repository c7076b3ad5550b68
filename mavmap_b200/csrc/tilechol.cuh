// tilechol.cuh — numeric phase of the sparse tile Cholesky of the reduced camera system and the PCG that uses it as its
// preconditioner (K3).  The symbolic phase lives in tilechol_plan.h (host, once per session).
//
// Replaces in the reference: the CHOLMOD factorisation + solve that Ceres runs behind SPARSE_SCHUR
// (src/base3d/bundle_adjustment.cc:555).  north_star asks for PCG on the reduced camera system; with an exact factorisation
// as the preconditioner the PCG converges to the 1e-13 relative residual that reproduces the reference's direct solve in one
// or two iterations instead of the 150 - 750 the two-level aggregate preconditioner needed.
//
// k_tc_factor: persistent, warp-specialised, left-looking.  One task per tile L(i,j) (48 x 48, column-major, 18 KB):
//     C = A(i,j) - sum_k L(i,k) L(j,k)'          (the list of k comes from the plan)
//     i == j:  L(j,j) = chol(C), and its inverse (both orientations) for the substitutions
//     i >  j:  L(i,j) = C inv(L(j,j))'
//   A producer lane per CTA takes tasks from a global counter (tasks are in a topological order of the elimination tree, so
//   a task only waits for tasks that were handed out before it), waits for the ready flags of the operand tiles, and moves
//   them into a shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx).  Four consumer warps
//   multiply from shared memory (6 x 3 register fragment per thread, fp64 FMA pipe) and publish the finished tile with a
//   release store of its flag.  No grid-wide barrier anywhere; every tile is written once, in a fixed summation order
//   (deterministic).
// k_tc_apply: both substitutions as one task list over 48-vector slots (node blocks applied through their explicit inverses
//   W, long rows / columns split into chunks of four tiles), same flag scheme.
// k_dpcg_*: the PCG iteration around it (fixed-order reductions: bit-reproducible from run to run and from GPU to GPU).
#pragma once
#include "common.cuh"
#include "tilechol_plan.h"

namespace mm {

// Packed device copies of the plan (built by the session from TileCholPlan):
//   sched[slot]  8 ints  {task, kind (0 off-diagonal L, 1 diagonal L, 2 W), first update, number of updates,
//                         tile the task finishes with (0: L(j,j); 1: W index of its own inverse; 2: W index of inv(L(i,i))), flag that guards it,
//                         has_a | tile column j << 1, tile_nunk | W store index}
//   upd[u]       int4    {tile of the first operand (L), tile of the second (L, or WR with bit 30 of .w set), flag of the first, flag of the second}
//   sdesc[k]     8 ints  {kind, out slot, base, tile row, first item, number of items, 0, 0}
//   items[q]     int2    {matrix selector << 28 | tile, source slot}
struct TcDev {
  int nt, nt_pose, nt_border, n_img, n_cam_border; int n_l, n_a_tiles, n_tasks, n_stasks, n_slots;
  const int *sc_tile, *sc_off, *a_tiles, *img_tile, *img_slot, *unk_of;
  const int64_t* col_ptr;
  const int* sched; const int4* upd; const int* sdesc; const int2* items;
  double *Ls;                               // second copy of the off-diagonal tiles of L with swizzled rows: the operand form of the fp64 tensor-core products
  double *L, *WC, *WR, *slots, *invd;       // tiles of L; W (inverse of the node blocks) column-major and row-major; 48-vector slots; 1 / diag(L) per tile row
  int *ready, *sflag, *counters;            // ready: per factor task, then per tile row (its inverse W(j,j) is stored); sflag: per slot; counters: [0] factor, [1] substitution tasks
  unsigned long long* trace;                // optional (test hook): {start, end} in ns per factor task, then per substitution task
  unsigned long long* trace_diag;           // optional (test hook): 8 phase stamps per diagonal task (indexed by W index of its inverse)
};

// ---------------------------------------------------------------- small PTX helpers (own names: match_tc.cu has its own copies)
__device__ __forceinline__ uint32_t tc_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(100000u) : "memory");
  } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tc_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ int tc_ld_acquire(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void tc_st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int tc_ld_relaxed(const int* p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void tc_fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// spin on a ready flag (relaxed polls with back-off; the caller issues tc_fence_acq_rel() before it touches the data)
__device__ __forceinline__ void tc_spin_flag(const int* p, int epoch) { unsigned ns = 32; while (tc_ld_relaxed(p) != epoch) { __nanosleep(ns); if (ns < 256) ns *= 2; } }
__device__ __forceinline__ void tc_wait_flag(const int* p, int epoch) { tc_spin_flag(p, epoch); tc_fence_acq_rel(); }
// orders earlier generic-proxy accesses to global memory (other CTAs' tile stores, observed through the acquire above)
// before the async-proxy reads of the bulk copies that follow
__device__ __forceinline__ void tc_fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void tc_bar_consumers() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ unsigned long long tc_gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// ---------------------------------------------------------------- assembly of A into the tiles
__global__ void __launch_bounds__(256) k_tc_zero_tiles(int n, const int* __restrict__ a_tiles, double* __restrict__ L) {
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    double2* t = reinterpret_cast<double2*>(L + (size_t)TC_TT * a_tiles[b]);
    for (int i = threadIdx.x; i < TC_TT / 2; i += blockDim.x) t[i] = make_double2(0.0, 0.0);
  }
}
// one thread per element of the stored 6 x 6 blocks of S
__global__ void k_tc_scatter_blocks(int64_t nblk, const int* __restrict__ sc_tile, const int* __restrict__ sc_off, const double* __restrict__ S, double* __restrict__ L) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < 36 * nblk; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / 36; const int e = (int)(i - 36 * b), r = e / 6, c = e - 6 * r;
    const int off = sc_off[b]; const bool tr = (off >> 30) & 1; const int base = off & ((1 << 30) - 1);
    L[(size_t)TC_TT * sc_tile[b] + base + (tr ? tile_elem(c, r) : tile_elem(r, c))] = S[i];
  }
}
// intrinsics border: B [n_img][ncb][6][9] (d pose x d intr) into the border rows of every pose column, C [(9 ncb)^2] into the
// border-border tiles (lower tiles; diagonal tiles in full)
__global__ void k_tc_scatter_border(TcDev P, const double* __restrict__ Bm, const double* __restrict__ Cm) {
  const int ncb = P.n_cam_border;
  const int64_t nB = (int64_t)P.n_img * ncb * 54, nC = (int64_t)81 * ncb * ncb;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nB + nC; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < nB) {
      const int m = (int)(i % 9), q = (int)((i / 9) % 6), c = (int)((i / 54) % ncb), a = (int)(i / (54 * (int64_t)ncb));
      const int j = P.img_tile[a];
      const int64_t t = P.col_ptr[j + 1] - P.nt_border + c / TC_BORDER_PER_TILE;      // border rows are the last rows of every pose column
      P.L[(size_t)TC_TT * t + tile_elem(9 * (c % TC_BORDER_PER_TILE) + m, 6 * P.img_slot[a] + q)] = Bm[i];
    } else {
      const int64_t k = i - nB; const int n9 = 9 * ncb;
      const int r = (int)(k / n9), c = (int)(k % n9);
      const int br = r / (9 * TC_BORDER_PER_TILE), bc = c / (9 * TC_BORDER_PER_TILE);
      if (br < bc) continue;
      const int64_t t = P.col_ptr[P.nt_pose + bc] + (br - bc);
      P.L[(size_t)TC_TT * t + tile_elem(r - 9 * TC_BORDER_PER_TILE * br, c - 9 * TC_BORDER_PER_TILE * bc)] = Cm[k];
    }
  }
}

// ---------------------------------------------------------------- factorisation
constexpr int TC_STAGES = 3;
constexpr int TC_CONSUMERS = 128;
constexpr int TC_FACTOR_THREADS = TC_CONSUMERS + 32;
constexpr size_t TC_FACTOR_SMEM = sizeof(double) * 2 * TC_TT * TC_STAGES + 2048;
enum { TC_KIND_UPD = 0, TC_KIND_FIN_OFF = 1, TC_KIND_FIN_DIAG = 2, TC_KIND_EXIT = 3, TC_KIND_WUPD = 4, TC_KIND_FIN_W = 5 };

// acc[a][b] (rows r0 + a, columns c0 + b)  -= / +=  sum_q A[q * 48 + r0 + a] * B[q * 48 + c0 + b]       (A, B in shared memory)
template <bool SUB>
__device__ __forceinline__ void tc_frag_gemm(double (&acc)[6][3], const double* A, const double* B, int r0, int c0) {
#pragma unroll 4
  for (int q = 0; q < TC_T; ++q) {
    const double2* a2 = reinterpret_cast<const double2*>(A + q * TC_T + r0);
    const double2 a01 = a2[0], a23 = a2[1], a45 = a2[2];
    const double* bp = B + q * TC_T + c0;
    const double b0 = bp[0], b1 = bp[1], b2 = bp[2];
    const double av[6] = { a01.x, a01.y, a23.x, a23.y, a45.x, a45.y };
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      if (SUB) { acc[a][0] -= av[a] * b0; acc[a][1] -= av[a] * b1; acc[a][2] -= av[a] * b2; }
      else { acc[a][0] += av[a] * b0; acc[a][1] += av[a] * b1; acc[a][2] += av[a] * b2; }
    }
  }
}

// inverse of the lower-triangular 48 x 48 tile Ls (shared, column-major) into W (shared, ROW-major W[r * 48 + c]); 48 threads
// take one column each (uniform loops: zero entries above the diagonal are carried along)
__device__ __forceinline__ void tc_tri_inverse_column(const double* __restrict__ Ls, double* __restrict__ W, int c) {
  for (int r = 0; r < TC_T; ++r) {
    double s0 = (r == c) ? 1.0 : 0.0, s1 = 0.0;
    int q = 0;
    for (; q + 1 < r; q += 2) { s0 -= Ls[q * TC_T + r] * W[q * TC_T + c]; s1 -= Ls[(q + 1) * TC_T + r] * W[(q + 1) * TC_T + c]; }
    if (q < r) s0 -= Ls[q * TC_T + r] * W[q * TC_T + c];
    W[r * TC_T + c] = (r >= c) ? (s0 + s1) / Ls[r * TC_T + r] : 0.0;
  }
}

// publish a finished tile: every consumer thread has stored its part; the barrier orders those stores before thread 0's
// fence, the fence makes them visible device-wide before the flag
__device__ __forceinline__ void tc_publish(int* flag, int epoch, int tid) {
  tc_bar_consumers();
  if (tid == 0) { __threadfence(); tc_st_release(flag, epoch); }
}

// 6 x 6 register fragment (64 threads cover a tile): 36 FMAs per six 16-byte shared-memory loads
template <bool SUB>
__device__ __forceinline__ void tc_frag_gemm66(double (&acc)[6][6], const double* A, const double* B, int r0, int c0, int q0, int q1) {
#pragma unroll 2
  for (int q = q0; q < q1; ++q) {
    const double2* a2 = reinterpret_cast<const double2*>(A + q * TC_T + r0);
    const double2* b2 = reinterpret_cast<const double2*>(B + q * TC_T + c0);
    const double2 a01 = a2[0], a23 = a2[1], a45 = a2[2], b01 = b2[0], b23 = b2[1], b45 = b2[2];
    const double av[6] = { a01.x, a01.y, a23.x, a23.y, a45.x, a45.y }, bv[6] = { b01.x, b01.y, b23.x, b23.y, b45.x, b45.y };
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = 0; b < 6; ++b) { if (SUB) acc[a][b] -= av[a] * bv[b]; else acc[a][b] += av[a] * bv[b]; }
  }
}

// ---- tile products on the fp64 tensor path (mma.sync.m8n8k4.f64, SASS DMMA).  Measured per 48 x 48 x 48 product and CTA at two
// CTAs per SM (tools/bench_dmma.cu): 6 x 3 FMA fragments 2.86 us, DMMA from column-major tiles (leading dimension 48: 4-way bank
// conflicts on the fragment loads) 2.37 us, DMMA from tiles whose rows are XOR-swizzled per column 1.83 us (96 % of the nominal
// fp64 rate).  The operands arrive by contiguous TMA copies, so the swizzle has to be their layout in global memory: every
// finished off-diagonal tile of L is stored a second time in that form (TcDev::Ls), and only the products read it.
__host__ __device__ inline int tile_elem_swz(int r, int c) { return c * TC_T + (r ^ ((c & 3) << 2)); }
__device__ __forceinline__ void tc_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// acc[i][j][e] = element (r0 + 8 i + g, c0 + 8 j + 2 t + e) of the warp's 24 x 24 quadrant, g = lane >> 2, t = lane & 3:
//   C(r, c) -= sum_q A(r, q) B(c, q)     (A, B swizzled tiles in shared memory)
__device__ __forceinline__ void tc_dmma_product(double (&acc)[3][3][2], const double* A, const double* B, int r0, int c0, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll 3
  for (int q0 = 0; q0 < TC_T; q0 += 4) {
    double a[3], b[3];
    const int col = (q0 + t) * TC_T, sw = t << 2;                  // (q0 + t) & 3 == t
#pragma unroll
    for (int i = 0; i < 3; ++i) { a[i] = -A[col + ((r0 + 8 * i + g) ^ sw)]; b[i] = B[col + ((c0 + 8 * i + g) ^ sw)]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) tc_dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// Consumers: two groups of 64 threads.  Every tile product is split along its inner dimension (group g takes k in
// [24 g, 24 g + 24)), each group keeps its partial sum in a 6 x 6 register fragment per thread (36 FMAs per six 16-byte
// shared-memory loads), and the finishing stage of the task adds the two partial sums to the entries of A - which the
// producer has put into the free half of the stage by TMA - before all 128 threads finish the tile (6 x 3 fragments).
// Fixed split, fixed order: deterministic.
__global__ void __launch_bounds__(TC_FACTOR_THREADS, 2) k_tc_factor(TcDev P, int epoch, int* __restrict__ fail) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  double* bufs = reinterpret_cast<double*>(tc_smem);                                  // [stage][A | B][TC_TT]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(tc_smem + sizeof(double) * 2 * TC_TT * TC_STAGES);   // full[STAGES], empty[STAGES]
  int4* meta = reinterpret_cast<int4*>(bars + 2 * TC_STAGES);                         // {task, stage kind | same << 9 | has_a << 10, aux, aux2}
  double* colbuf = reinterpret_cast<double*>(meta + TC_STAGES);                       // [2][48] pivot columns, [48] reciprocal diagonal
  double* invd = colbuf + 2 * TC_T;
  const int tid = threadIdx.x;
  const uint32_t bar0 = tc_smem_addr(bars);
  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { tc_mbar_init(bar0 + 8 * s, 1); tc_mbar_init(bar0 + 8 * (TC_STAGES + s), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr uint32_t TILE_BYTES = (uint32_t)(sizeof(double) * TC_TT);
  if (tid >= TC_CONSUMERS) {
    // ------------------------------------------------ producer warp
    // Lane n looks after the n-th update of a batch of 32: it reads the update record and probes both ready flags (all
    // lanes at once: one round trip to L2 per batch instead of three per update), then the lanes issue their bulk copies
    // one after the other in list order (the order of the sum is part of the result).
    const int lane = tid - TC_CONSUMERS;
    int stage = 0; uint32_t phase = 0;
    for (;;) {
      int slot = 0;
      if (lane == 0) slot = atomicAdd(P.counters + 0, 1);
      slot = __shfl_sync(0xffffffffu, slot, 0);
      if (slot >= P.n_tasks) {
        if (lane == 0) { tc_mbar_wait(bar0 + 8 * (TC_STAGES + stage), phase ^ 1); meta[stage] = make_int4(0, TC_KIND_EXIT, 0, 0); tc_mbar_arrive(bar0 + 8 * stage); }
        break;
      }
      const int dv = lane < 8 ? P.sched[8 * (size_t)slot + lane] : 0;
      const int t = __shfl_sync(0xffffffffu, dv, 0), kind = __shfl_sync(0xffffffffu, dv, 1), ub = __shfl_sync(0xffffffffu, dv, 2), nu = __shfl_sync(0xffffffffu, dv, 3);
      const int fin_idx = __shfl_sync(0xffffffffu, dv, 4), fin_flag = __shfl_sync(0xffffffffu, dv, 5), hasa_j = __shfl_sync(0xffffffffu, dv, 6), aux = __shfl_sync(0xffffffffu, dv, 7);
      if (P.trace && lane == 0) P.trace[2 * (size_t)t] = tc_gtimer();
      const int upd_kind = kind == 2 ? TC_KIND_WUPD : TC_KIND_UPD;
      for (int base = 0; base < nu; base += 32) {
        const int cnt = min(32, nu - base);
        int4 rec = make_int4(0, 0, 0, 0); bool rdy = true;
        if (lane < cnt) {
          rec = P.upd[(size_t)ub + base + lane];
          const int fa = tc_ld_relaxed(P.ready + rec.z), fb = tc_ld_relaxed(P.ready + (rec.w & 0x3fffffff));
          rdy = fa == epoch && fb == epoch;
        }
        tc_fence_acq_rel();                            // (for the lanes whose flags were already set; the others fence after their spin)
        tc_fence_proxy_async_global();
        for (int n = 0; n < cnt; ++n) {
          if (lane == n) {
            if (!rdy) { tc_spin_flag(P.ready + rec.z, epoch); tc_spin_flag(P.ready + (rec.w & 0x3fffffff), epoch); tc_fence_acq_rel(); tc_fence_proxy_async_global(); }
            tc_mbar_wait(bar0 + 8 * (TC_STAGES + stage), phase ^ 1);
            const bool b_wr = (rec.w >> 30) & 1, same = !b_wr && rec.x == rec.y;
            meta[stage] = make_int4(t, upd_kind | ((int)same << 9), 0, 0);
            const uint32_t dstA = tc_smem_addr(bufs + (size_t)2 * TC_TT * stage);
            tc_mbar_expect_tx(bar0 + 8 * stage, TILE_BYTES * (same ? 1 : 2));
            // products of L tiles run on the tensor path and read the swizzled copies; the W products (L tile x row-major W) stay on the FMA path
            const double* srcL = upd_kind == TC_KIND_UPD ? P.Ls : P.L;
            tc_bulk_g2s(dstA, srcL + (size_t)TC_TT * rec.x, TILE_BYTES, bar0 + 8 * stage);
            if (!same) tc_bulk_g2s(dstA + TILE_BYTES, (b_wr ? P.WR : srcL) + (size_t)TC_TT * rec.y, TILE_BYTES, bar0 + 8 * stage);
          }
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          __syncwarp();
        }
      }
      if (lane == 0) {
        // finishing stage: the entries of A(i,j) (what the assembly left in the tile's own storage) into the first half; into the
        // second half L(j,j) for an off-diagonal tile of L (triangular solve), inv(L(i,i)) for a tile of W
        tc_mbar_wait(bar0 + 8 * (TC_STAGES + stage), phase ^ 1);
        const uint32_t dstA = tc_smem_addr(bufs + (size_t)2 * TC_TT * stage);
        const int has_a = hasa_j & 1, jcol = hasa_j >> 1;
        if (kind != 1) { tc_wait_flag(P.ready + fin_flag, epoch); tc_fence_proxy_async_global(); }
        meta[stage] = make_int4(t, (kind == 1 ? TC_KIND_FIN_DIAG : (kind == 2 ? TC_KIND_FIN_W : TC_KIND_FIN_OFF)) | (has_a << 10) | (jcol << 12), aux, fin_idx);
        const uint32_t bytes = TILE_BYTES * ((has_a ? 1 : 0) + (kind != 1 ? 1 : 0));
        if (bytes == 0) tc_mbar_arrive(bar0 + 8 * stage);
        else {
          tc_mbar_expect_tx(bar0 + 8 * stage, bytes);
          if (has_a) tc_bulk_g2s(dstA, P.L + (size_t)TC_TT * t, TILE_BYTES, bar0 + 8 * stage);
          if (kind != 1) tc_bulk_g2s(dstA + TILE_BYTES, (kind == 0 ? P.L : P.WC) + (size_t)TC_TT * fin_idx, TILE_BYTES, bar0 + 8 * stage);
        }
      }
      if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
      __syncwarp();
    }
    return;
  }
  // -------------------------------------------------- consumers (4 warps = 2 groups)
  const int grp = tid >> 6, gt = tid & 63, R0 = 6 * (gt >> 3), C0 = 6 * (gt & 7);       // 6 x 6 fragment of the group's partial sum
  const int tr = tid >> 4, tcn = tid & 15, r0 = 6 * tr, c0 = 3 * tcn, lane = tid & 31;  // 6 x 3 fragment of the finishing steps
  double acc[6][6];                                       // W products: split along k between the two groups (FMA path)
  double accd[3][3][2];                                   // L products: the warp's 24 x 24 quadrant in the DMMA fragment layout
  const int wq = tid >> 5, qr0 = 24 * (wq >> 1), qc0 = 24 * (wq & 1);
  int cur = -1;                                          // task whose partial sum `acc` / `accd` holds
  int stage = 0; uint32_t phase = 0;
  for (;;) {
    tc_mbar_wait(bar0 + 8 * stage, phase);
    const int4 m = meta[stage];
    const int skind = m.y & 0xff;
    if (skind == TC_KIND_EXIT) break;
    const int t = m.x;
    double* A = bufs + (size_t)2 * TC_TT * stage; double* B = A + TC_TT;
    if (skind == TC_KIND_UPD || skind == TC_KIND_WUPD) {
      if (cur != t) {
        cur = t;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = 0; b < 6; ++b) acc[a][b] = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) { accd[i][j][0] = 0.0; accd[i][j][1] = 0.0; }
      }
      const int q0 = grp * (TC_T / 2);
      if (skind == TC_KIND_UPD) tc_dmma_product(accd, A, ((m.y >> 9) & 1) ? A : B, qr0, qc0, lane);
      else tc_frag_gemm66<false>(acc, A, B, R0, C0, q0, q0 + TC_T / 2);
    } else {
      // ---- finishing stage: C = A(i,j) + partial sums, assembled in the first half of the stage.  Column-major for the tiles
      // of L, row-major for W (there it is the second operand of the final product; W tiles have no entries of A).
      const bool rowmajor = skind == TC_KIND_FIN_W, has_a = (m.y >> 10) & 1, mine = cur == t;
      if (P.trace_diag && tid == 0 && skind == TC_KIND_FIN_DIAG) P.trace_diag[8 * (size_t)m.w + 0] = tc_gtimer();
      if (!rowmajor) {
        // tiles of L: the partial sum sits in the DMMA fragments, every thread owns 18 entries of the tile (column-major in A)
        const int g = lane >> 2, tq = lane & 3;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              double* e = A + (qc0 + 8 * j + 2 * tq + e2) * TC_T + qr0 + 8 * i + g;
              double v = has_a ? *e : 0.0;
              if (mine) v += accd[i][j][e2];
              *e = v;
            }
      } else if (grp == 0) {
#pragma unroll
        for (int b = 0; b < 6; ++b)
#pragma unroll
          for (int a = 0; a < 6; a += 2) {
            if (!rowmajor) {
              double2* e = reinterpret_cast<double2*>(A + (C0 + b) * TC_T + R0 + a);
              double2 v = has_a ? *e : make_double2(0.0, 0.0);
              if (mine) { v.x += acc[a][b]; v.y += acc[a + 1][b]; }
              *e = v;
            } else { A[(R0 + a) * TC_T + C0 + b] = mine ? acc[a][b] : 0.0; A[(R0 + a + 1) * TC_T + C0 + b] = mine ? acc[a + 1][b] : 0.0; }
          }
      }
      tc_bar_consumers();
      if (rowmajor && grp == 1 && mine) {
#pragma unroll
        for (int b = 0; b < 6; ++b)
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            double* e = rowmajor ? A + (R0 + a) * TC_T + C0 + b : A + (C0 + b) * TC_T + R0 + a;
            *e += acc[a][b];
          }
      }
      cur = -1;
      tc_bar_consumers();
      if (skind == TC_KIND_FIN_OFF) {
        // L(i,j) L(j,j)' = C:  one thread per row of C, right-looking substitution with the row in registers; L(j,j) (second half
        // of the stage) is read as broadcasts, its reciprocal diagonal comes from the diagonal task.  No inverse on this path.
        const int jcol = m.y >> 12;
        if (tid < TC_T) invd[tid] = __ldcg(P.invd + (size_t)TC_T * jcol + tid);
        tc_bar_consumers();
        if (tid < TC_T) {
          // in place in the first half (column-major: thread r owns the words c * 48 + r, conflict-free), panels of four columns
          const int r = tid;
          for (int p = 0; p < TC_T; p += 4) {
            const double* L0 = B + p * TC_T; const double* L1 = L0 + TC_T; const double* L2 = L1 + TC_T;           // columns p .. p+2 of L(j,j)
            const double x0 = A[p * TC_T + r] * invd[p];
            const double x1 = (A[(p + 1) * TC_T + r] - x0 * L0[p + 1]) * invd[p + 1];
            const double x2 = (A[(p + 2) * TC_T + r] - x0 * L0[p + 2] - x1 * L1[p + 2]) * invd[p + 2];
            const double x3 = (A[(p + 3) * TC_T + r] - x0 * L0[p + 3] - x1 * L1[p + 3] - x2 * L2[p + 3]) * invd[p + 3];
            A[p * TC_T + r] = x0; A[(p + 1) * TC_T + r] = x1; A[(p + 2) * TC_T + r] = x2; A[(p + 3) * TC_T + r] = x3;
            const double* L3 = L2 + TC_T;
#pragma unroll 4
            for (int c2 = p + 4; c2 < TC_T; ++c2) A[c2 * TC_T + r] -= x0 * L0[c2] + x1 * L1[c2] + x2 * L2[c2] + x3 * L3[c2];
          }
        }
        tc_bar_consumers();
        {
          double2* dst = reinterpret_cast<double2*>(P.L + (size_t)TC_TT * t); const double2* src = reinterpret_cast<const double2*>(A);
          double2* dsw = reinterpret_cast<double2*>(P.Ls + (size_t)TC_TT * t);           // operand form for the products: rows r ^ 4 (c & 3); pairs of rows stay together
#pragma unroll 3
          for (int e = tid; e < TC_TT / 2; e += TC_CONSUMERS) {
            const double2 v = src[e];
            __stcg(dst + e, v);
            const int c = e / (TC_T / 2), r = 2 * (e - c * (TC_T / 2));
            __stcg(dsw + (tile_elem_swz(r, c) >> 1), v);
          }
        }
      } else if (skind == TC_KIND_FIN_W) {
        // W(i,j) = -inv(L(i,i)) C   (C row-major in A, inverse column-major in B)
        double out[6][3];
#pragma unroll
        for (int a = 0; a < 6; ++a) { out[a][0] = 0.0; out[a][1] = 0.0; out[a][2] = 0.0; }
        tc_frag_gemm<true>(out, B, A, r0, c0);
        const size_t wi = (size_t)TC_TT * m.z;
        double* dC = P.WC + wi; double* dR = P.WR + wi;
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int a = 0; a < 6; a += 2) __stcg(reinterpret_cast<double2*>(dC + (c0 + b) * TC_T + r0 + a), make_double2(out[a][b], out[a + 1][b]));
#pragma unroll
        for (int a = 0; a < 6; ++a) { __stcg(dR + (r0 + a) * TC_T + c0, out[a][0]); __stcg(dR + (r0 + a) * TC_T + c0 + 1, out[a][1]); __stcg(dR + (r0 + a) * TC_T + c0 + 2, out[a][2]); }
      } else {
        // diagonal tile: right-looking Cholesky in panels of six columns.  The tile lives in 6 x 3 register fragments; per
        // panel the owners put its columns into shared memory, 48 row threads factor the 6 x 6 diagonal block (each its own
        // copy: cheaper than a barrier) and solve their row of the panel, then all threads update their fragments.
        if (P.trace_diag && tid == 0) P.trace_diag[8 * (size_t)m.w + 1] = tc_gtimer();
        const int nunk = m.z, jcol = m.y >> 12;
        double f[6][3];
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int a = 0; a < 6; a += 2) { const double2 v = *reinterpret_cast<const double2*>(A + (c0 + b) * TC_T + r0 + a); f[a][b] = v.x; f[a + 1][b] = v.y; }
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) if (r0 + a >= nunk || c0 + b >= nunk) f[a][b] = (r0 + a == c0 + b) ? 1.0 : 0.0;      // padding: unit diagonal
        double* Pn = B;                                    // panel: 6 columns x 48 rows, column-major
        bool bad = false;
        for (int p = 0; p < TC_T; p += 6) {
          if (c0 >= p && c0 < p + 6) {
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
              for (int a = 0; a < 6; a += 2) *reinterpret_cast<double2*>(Pn + (c0 - p + b) * TC_T + r0 + a) = make_double2(f[a][b], f[a + 1][b]);
          }
          tc_bar_consumers();
          if (tid < TC_T) {
            double D[6][6], sinv[6];
#pragma unroll
            for (int j = 0; j < 6; ++j)
#pragma unroll
              for (int i = j; i < 6; ++i) D[i][j] = Pn[j * TC_T + p + i];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
              double d = D[k][k];
              if (!(d > 0.0) || !isfinite(d)) { bad = true; d = 1.0; }
              const double sk = rsqrt(d);
              sinv[k] = sk;
#pragma unroll
              for (int i = k + 1; i < 6; ++i) D[i][k] *= sk;
#pragma unroll
              for (int j = k + 1; j < 6; ++j)
#pragma unroll
                for (int i = j; i < 6; ++i) D[i][j] -= D[i][k] * D[j][k];
            }
            const int r = tid;
            if (r >= p) {        // row r of the panel: x L_dd' = c  (rows inside the diagonal block reproduce L_dd; their upper part is dropped on output)
              double x[6];
#pragma unroll
              for (int j = 0; j < 6; ++j) {
                double v = Pn[j * TC_T + r];
#pragma unroll
                for (int q = 0; q < j; ++q) v -= x[q] * D[j][q];
                x[j] = v * sinv[j];
              }
#pragma unroll
              for (int j = 0; j < 6; ++j) Pn[j * TC_T + r] = x[j];
            }
            if (tid == 0) {
#pragma unroll
              for (int k = 0; k < 6; ++k) invd[p + k] = sinv[k];
            }
          }
          tc_bar_consumers();
          if (c0 >= p + 6) {
            double lr[6][6], lc[3][6];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
#pragma unroll
              for (int a = 0; a < 6; a += 2) { const double2 v = *reinterpret_cast<const double2*>(Pn + q * TC_T + r0 + a); lr[a][q] = v.x; lr[a + 1][q] = v.y; }
#pragma unroll
              for (int b = 0; b < 3; ++b) lc[b][q] = Pn[q * TC_T + c0 + b];
            }
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
              for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int q = 0; q < 6; ++q) f[a][b] -= lr[a][q] * lc[b][q];
          } else if (c0 >= p) {
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
              for (int a = 0; a < 6; a += 2) { const double2 v = *reinterpret_cast<const double2*>(Pn + (c0 - p + b) * TC_T + r0 + a); f[a][b] = v.x; f[a + 1][b] = v.y; }
          }
          tc_bar_consumers();
        }
        if (P.trace_diag && tid == 0) P.trace_diag[8 * (size_t)m.w + 2] = tc_gtimer();
        if (bad && tid == 0) *fail = 1;
        // L (lower, zeros above) -> stage half A (column-major) and global memory
        double* dstL = P.L + (size_t)TC_TT * t;
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int a = 0; a < 6; a += 2) {
            const double v0 = (r0 + a >= c0 + b) ? f[a][b] : 0.0, v1 = (r0 + a + 1 >= c0 + b) ? f[a + 1][b] : 0.0;
            *reinterpret_cast<double2*>(A + (c0 + b) * TC_T + r0 + a) = make_double2(v0, v1);
            __stcg(reinterpret_cast<double2*>(dstL + (c0 + b) * TC_T + r0 + a), make_double2(v0, v1));
          }
        if (tid < TC_T) __stcg(P.invd + (size_t)TC_T * jcol + tid, invd[tid]);
        // L(j,j) is what the tiles below it wait for: publish it now; the inverse (needed by the node-block inverses and the
        // substitutions only) follows under its own flag
        tc_publish(P.ready + t, epoch, tid);
        if (P.trace && tid == 0) P.trace[2 * (size_t)t + 1] = tc_gtimer();
        // inverse W of L = [[L11, 0], [L21, L22]] (24 x 24 blocks), row-major in the free half B:
        //   W11, W22: one thread per column, right-looking substitution in panels of four rows (two warps side by side)
        //   W21 = -W22 (L21 W11): two 24^3 products on 96 threads
        if (P.trace_diag && tid == 0) P.trace_diag[8 * (size_t)m.w + 3] = tc_gtimer();
        constexpr int HB = TC_T / 2;
        if (tid < HB || (tid >= 32 && tid < 32 + HB)) {
          const int lo = tid < HB ? 0 : HB, c = tid < HB ? tid : HB + tid - 32;
#pragma unroll 8
          for (int r = 0; r < TC_T; ++r) B[r * TC_T + c] = (r == c) ? 1.0 : 0.0;
          for (int pp = lo; pp < lo + HB; pp += 4) {
            const double* L0 = A + pp * TC_T; const double* L1 = L0 + TC_T; const double* L2 = L1 + TC_T; const double* L3 = L2 + TC_T;   // columns pp .. pp+3 of L
            const double x0 = B[pp * TC_T + c] * invd[pp];
            const double x1 = (B[(pp + 1) * TC_T + c] - L0[pp + 1] * x0) * invd[pp + 1];
            const double x2 = (B[(pp + 2) * TC_T + c] - L0[pp + 2] * x0 - L1[pp + 2] * x1) * invd[pp + 2];
            const double x3 = (B[(pp + 3) * TC_T + c] - L0[pp + 3] * x0 - L1[pp + 3] * x1 - L2[pp + 3] * x2) * invd[pp + 3];
            B[pp * TC_T + c] = x0; B[(pp + 1) * TC_T + c] = x1; B[(pp + 2) * TC_T + c] = x2; B[(pp + 3) * TC_T + c] = x3;
#pragma unroll 4
            for (int r2 = pp + 4; r2 < lo + HB; ++r2) B[r2 * TC_T + c] -= L0[r2] * x0 + L1[r2] * x1 + L2[r2] * x2 + L3[r2] * x3;
          }
        }
        tc_bar_consumers();
        if (P.trace_diag && tid == 0) P.trace_diag[8 * (size_t)m.w + 4] = tc_gtimer();
        const int bi = tid >> 2, bg = 6 * (tid & 3);         // 96 threads: row bi of the lower-left block, columns bg .. bg + 5
        double w6[6] = {0, 0, 0, 0, 0, 0};
        if (tid < 4 * HB) {
#pragma unroll 4
          for (int q = 0; q < HB; ++q) {
            const double l = A[q * TC_T + HB + bi];
#pragma unroll
            for (int k2 = 0; k2 < 6; ++k2) w6[k2] += l * B[q * TC_T + bg + k2];
          }
#pragma unroll
          for (int k2 = 0; k2 < 6; ++k2) B[(HB + bi) * TC_T + bg + k2] = w6[k2];          // T = L21 W11 (rows 24.. are not read by this step)
        }
        tc_bar_consumers();
        if (tid < 4 * HB) {
#pragma unroll
          for (int k2 = 0; k2 < 6; ++k2) w6[k2] = 0.0;
#pragma unroll 4
          for (int q = 0; q < HB; ++q) {
            const double wv = B[(HB + bi) * TC_T + HB + q];
#pragma unroll
            for (int k2 = 0; k2 < 6; ++k2) w6[k2] -= wv * B[(HB + q) * TC_T + bg + k2];
          }
        }
        tc_bar_consumers();
        if (tid < 4 * HB) {
#pragma unroll
          for (int k2 = 0; k2 < 6; ++k2) B[(HB + bi) * TC_T + bg + k2] = w6[k2];
        }
        tc_bar_consumers();
        if (P.trace_diag && tid == 0) P.trace_diag[8 * (size_t)m.w + 5] = tc_gtimer();
        {
          const size_t wi = (size_t)TC_TT * (size_t)m.w;
          double* dC = P.WC + wi; double* dR = P.WR + wi;
          for (int e = tid; e < TC_TT; e += TC_CONSUMERS) {
            const int r = e / TC_T, c = e - TC_T * r;                 // e = r * 48 + c
            const double v = B[e];
            __stcg(dR + e, v);
            __stcg(dC + c * TC_T + r, v);
          }
        }
      }
      if (P.trace_diag && tid == 0 && skind == TC_KIND_FIN_DIAG) P.trace_diag[8 * (size_t)m.w + 6] = tc_gtimer();
      if (skind == TC_KIND_FIN_DIAG) tc_publish(P.ready + P.n_tasks + (m.y >> 12), epoch, tid);       // W(j,j) stored
      else { tc_publish(P.ready + t, epoch, tid); if (P.trace && tid == 0) P.trace[2 * (size_t)t + 1] = tc_gtimer(); }
      if (P.trace_diag && tid == 0 && skind == TC_KIND_FIN_DIAG) P.trace_diag[8 * (size_t)m.w + 7] = tc_gtimer();
      // (the partial sum is dead from here to the next task: saying so keeps its registers free for the finishing code)
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) acc[a][b] = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { accd[i][j][0] = 0.0; accd[i][j][1] = 0.0; }
    }
    __syncwarp();
    if (lane == 0) tc_mbar_arrive(bar0 + 8 * (TC_STAGES + stage));
    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
  }
}

// ---------------------------------------------------------------- substitutions (both sweeps, one task list, one launch)
// Every task writes one 48-vector slot and then releases the slot's flag; it reads slots written by earlier tasks of the list.
//   MV      out = sum_items M v      warps split the items; lane = row (coalesced column-major reads), v broadcast by shuffles
//   MVT     out = sum_items M' v     lane = column: a 48-long dot product over the lane's own contiguous column
//   SUM     out = base - sum_items v
//   MV_OUT  MV whose result also goes to the caller's vector (the solution)
constexpr int TC_APPLY_THREADS = 128;
constexpr int TC_APPLY_WARPS = TC_APPLY_THREADS / 32;
constexpr size_t TC_APPLY_SMEM = sizeof(double) * TC_TT * TC_APPLY_WARPS + 256;      // one tile buffer per warp + their mbarriers
// Task descriptor: 16 ints {kind, out slot, base, tile row, first item, number of items, -, -, items 0..3 inline as (matrix, source) pairs}.
// Every warp owns an 18 KB shared-memory buffer: lane 0 starts the 1-D TMA bulk copy of the item's tile BEFORE the warp waits
// for the flag of the item's source vector, so the tile is already on chip when the vector becomes available.
__global__ void __launch_bounds__(TC_APPLY_THREADS) k_tc_apply(TcDev P, int epoch, const double* __restrict__ rhs, double* __restrict__ z, const int* __restrict__ done) {
  if (done && *done) return;
  extern __shared__ __align__(128) unsigned char ap_smem[];
  double* tiles = reinterpret_cast<double*>(ap_smem);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ap_smem + sizeof(double) * TC_TT * TC_APPLY_WARPS);
  __shared__ double red[TC_APPLY_WARPS][TC_T];
  __shared__ double xs[TC_APPLY_WARPS][TC_T];
  __shared__ int s_desc[16];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int NW = TC_APPLY_WARPS;
  constexpr uint32_t TILE_BYTES = (uint32_t)(sizeof(double) * TC_TT);
  double* Tw = tiles + (size_t)TC_TT * w;
  const uint32_t bar = tc_smem_addr(bars + w), tdst = tc_smem_addr(Tw);
  uint32_t phase = 0;
  if (lane == 0) { tc_mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (;;) {
    __syncthreads();
    if (w == 0) {
      int k = 0;
      if (lane == 0) k = atomicAdd(P.counters + 1, 1);
      k = __shfl_sync(0xffffffffu, k, 0);
      if (lane < 16) s_desc[lane] = k < P.n_stasks ? P.sdesc[16 * (size_t)k + lane] : -1;
      __syncwarp();
      if (lane == 0) s_desc[6] = k;
    }
    __syncthreads();
    const int kind = s_desc[0];
    if (kind < 0) break;
    const int out = s_desc[1], base = s_desc[2], tile = s_desc[3], i0 = s_desc[4], ni = s_desc[5], k = s_desc[6];
    if (P.trace && tid == 0) P.trace[2 * ((size_t)P.n_tasks + k)] = tc_gtimer();
    double result = 0.0;                              // threads 0..47: entry tid of the output vector
    if (kind == TC_ST_SUM) {
      // all flags polled side by side (one round trip instead of one per item)
      for (int q = tid - 1; q < ni; q += TC_APPLY_THREADS) {
        if (q >= 0) tc_spin_flag(P.sflag + (q < 4 ? s_desc[9 + 2 * q] : P.items[i0 + q].y), epoch);
        else if (base >= 0) tc_spin_flag(P.sflag + base, epoch);
      }
      tc_fence_acq_rel();
      __syncthreads();
      if (tid < TC_T) {
        if (base >= 0) result = __ldcg(P.slots + (size_t)TC_T * base + tid);
        else { const int u = P.unk_of[(size_t)TC_T * tile + tid]; result = u >= 0 ? rhs[u] : 0.0; }
        for (int q = 0; q < ni; ++q) result -= __ldcg(P.slots + (size_t)TC_T * (q < 4 ? s_desc[9 + 2 * q] : P.items[i0 + q].y) + tid);
      }
    } else {
      double a0 = 0.0, a1 = 0.0;                      // MV: rows lane, 32 + lane;  MVT: columns lane, 32 + lane
      for (int q = w; q < ni; q += NW) {              // this warp's items
        int2 it;
        if (q < 4) it = make_int2(s_desc[8 + 2 * q], s_desc[9 + 2 * q]); else it = P.items[i0 + q];
        const int sel = it.x >> 28, idx = it.x & ((1 << 28) - 1), src = it.y;
        const double* T = (sel == TC_MAT_L ? P.L : (sel == TC_MAT_WC ? P.WC : P.WR)) + (size_t)TC_TT * idx;
        if (lane == 0) {
          tc_mbar_expect_tx(bar, TILE_BYTES);
          tc_bulk_g2s(tdst, T, TILE_BYTES, bar);
          tc_spin_flag(P.sflag + src, epoch);
          tc_fence_acq_rel();
        }
        __syncwarp();
        const double v0 = __ldcg(P.slots + (size_t)TC_T * src + lane), v1 = lane < 16 ? __ldcg(P.slots + (size_t)TC_T * src + 32 + lane) : 0.0;
        tc_mbar_wait(bar, phase); phase ^= 1;
        if (kind != TC_ST_MVT) {
#pragma unroll 8
          for (int c = 0; c < TC_T; ++c) {
            const double vc = __shfl_sync(0xffffffffu, c < 32 ? v0 : v1, c & 31);
            a0 += Tw[c * TC_T + lane] * vc;
            if (lane < 16) a1 += Tw[c * TC_T + 32 + lane] * vc;
          }
        } else {
          xs[w][lane] = v0; if (lane < 16) xs[w][32 + lane] = v1;
          __syncwarp();
          // column `lane` (and 32 + lane) of the tile, walked from a lane-dependent row so that the lanes hit different banks
          const double* c0p = Tw + (size_t)lane * TC_T; const double* c1p = Tw + (size_t)(32 + (lane & 15)) * TC_T;
          int r = lane;
#pragma unroll 8
          for (int n = 0; n < TC_T; ++n) {
            const double xv = xs[w][r];
            a0 += c0p[r] * xv;
            if (lane < 16) a1 += c1p[r] * xv;
            if (++r == TC_T) r = 0;
          }
        }
        __syncwarp();                                  // the buffer is free for the warp's next item
      }
      red[w][lane] = a0; if (lane < 16) red[w][32 + lane] = a1;
      __syncthreads();
      if (tid < TC_T) {
#pragma unroll
        for (int q = 0; q < NW; ++q) result += red[q][tid];
        if (kind == TC_ST_MV_OUT) { const int u = P.unk_of[(size_t)TC_T * tile + tid]; if (u >= 0) z[u] = result; }
      }
    }
    if (tid < TC_T) __stcg(P.slots + (size_t)TC_T * out + tid, result);
    __syncthreads();
    if (tid == 0) { __threadfence(); tc_st_release(P.sflag + out, epoch); if (P.trace) P.trace[2 * ((size_t)P.n_tasks + k) + 1] = tc_gtimer(); }
  }
}

// ---------------------------------------------------------------- PCG around the factorisation
// Unknowns: [6 n_img pose parameters | 9 ncb intrinsics].  All dot products are reduced in a fixed order by ONE CTA, so the
// solve is bit-reproducible (run to run, and between the ranks of a sharded session, which solve the same replicated system).
// sc: [0] b.b  [1] r.z  [2] r.z of the previous iteration  [3] p.Ap  [4] r.r  [5] r.r after the first iteration ;  ic: [0] converged flag, [1] iterations done
struct DpcgVec { int n; double *x, *r, *z, *p, *Ap; const double* b; double* sc; int* ic; double tol2; int max_iter; };

__device__ __forceinline__ double dpcg_block_sum_all(double v, double* red) {      // 1024 threads; the sum, in every thread
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0;
  t = warp_sum(t);
  return t;
}
__global__ void __launch_bounds__(1024) k_dpcg_init(DpcgVec V) {
  __shared__ double red[32];
  double bb = 0.0;
  for (int i = threadIdx.x; i < V.n; i += blockDim.x) { const double b = V.b[i]; V.x[i] = 0.0; V.r[i] = b; V.p[i] = 0.0; bb += b * b; }
  bb = dpcg_block_sum_all(bb, red);
  if (threadIdx.x == 0) { V.sc[0] = bb; V.sc[1] = 0.0; V.sc[2] = 0.0; V.ic[0] = bb > 0.0 ? 0 : 1; V.ic[1] = 0; }
}
// after z = M^-1 r:  beta = r.z / (r.z)_old,  p = z + beta p
__global__ void __launch_bounds__(1024) k_dpcg_direction(DpcgVec V) {
  if (V.ic[0]) return;
  __shared__ double red[32];
  double rz = 0.0;
  for (int i = threadIdx.x; i < V.n; i += blockDim.x) rz += V.r[i] * V.z[i];
  rz = dpcg_block_sum_all(rz, red);
  const double rz_old = V.sc[1];
  const double beta = V.ic[1] == 0 ? 0.0 : rz / rz_old;
  __syncthreads();
  for (int i = threadIdx.x; i < V.n; i += blockDim.x) V.p[i] = V.z[i] + beta * V.p[i];
  if (threadIdx.x == 0) { V.sc[2] = rz_old; V.sc[1] = rz; }
}
// after Ap = S p:  alpha = r.z / p.Ap,  x += alpha p,  r -= alpha Ap,  convergence test on r.r
__global__ void __launch_bounds__(1024) k_dpcg_update(DpcgVec V) {
  if (V.ic[0]) return;
  __shared__ double red[32];
  double pap = 0.0;
  for (int i = threadIdx.x; i < V.n; i += blockDim.x) pap += V.p[i] * V.Ap[i];
  pap = dpcg_block_sum_all(pap, red);
  const double alpha = pap > 0.0 ? V.sc[1] / pap : 0.0;
  double rr = 0.0;
  for (int i = threadIdx.x; i < V.n; i += blockDim.x) { V.x[i] += alpha * V.p[i]; const double r = V.r[i] - alpha * V.Ap[i]; V.r[i] = r; rr += r * r; }
  rr = dpcg_block_sum_all(rr, red);
  if (threadIdx.x == 0) {
    V.sc[3] = pap; V.sc[4] = rr; if (V.ic[1] == 0) V.sc[5] = rr;
    const int it = V.ic[1] + 1; V.ic[1] = it;
    if (rr <= V.tol2 * V.sc[0] || it >= V.max_iter || !(rr == rr) || !(pap > 0.0)) V.ic[0] = 1;
  }
}
// Ap = S p on the block-CSR of the pose part (one warp per block-row) + the dense intrinsics border
//   pose rows:  Ap_a = sum_b S_ab p_b + sum_c B_ac p_c          intrinsics rows: k_dpcg_spmv_border
__global__ void __launch_bounds__(128) k_dpcg_spmv(int n_img, const int* __restrict__ row_start, const int* __restrict__ row_col, const int* __restrict__ row_blk,
                                                   const double* __restrict__ S, const double* __restrict__ p, double* __restrict__ Ap,
                                                   int ncb, const double* __restrict__ Bm, const int* __restrict__ done) {
  if (*done) return;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_img) return;
  double y[6] = {0, 0, 0, 0, 0, 0};
  for (int e = row_start[w] + lane; e < row_start[w + 1]; e += 32) {
    const int col = row_col[e]; const int bid = row_blk[e];
    const bool tr = bid < 0;
    const double* B = S + 36 * (size_t)(tr ? -bid - 1 : bid);
    double pv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) pv[k] = p[6 * (size_t)col + k];
    double Bv[36];
    const double2* B2 = reinterpret_cast<const double2*>(B);
#pragma unroll
    for (int k = 0; k < 18; ++k) { const double2 t = B2[k]; Bv[2 * k] = t.x; Bv[2 * k + 1] = t.y; }
    if (!tr) {
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = 0; c < 6; ++c) y[a] += Bv[6 * a + c] * pv[c];
    } else {
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = 0; c < 6; ++c) y[a] += Bv[6 * c + a] * pv[c];
    }
  }
  // border: lanes split the 9 ncb intrinsics columns
  const double* pi = p + 6 * (size_t)n_img;
  for (int m = lane; m < 9 * ncb; m += 32) {
    const double pm = pi[m]; const int c = m / 9, mm_ = m - 9 * c;
    const double* Ba = Bm + ((size_t)w * ncb + c) * 54;
#pragma unroll
    for (int a = 0; a < 6; ++a) y[a] += Ba[9 * a + mm_] * pm;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) y[k] = warp_sum(y[k]);
  if (lane < 6) {
    double v = y[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) if (lane == k) v = y[k];
    Ap[6 * (size_t)w + lane] = v;
  }
}
// intrinsics rows:  Ap_c = sum_a B_ac' p_a + sum_c' C_cc' p_c' ;  one CTA per camera, fixed-order reduction
__global__ void __launch_bounds__(256) k_dpcg_spmv_border(int n_img, int ncb, const double* __restrict__ Bm, const double* __restrict__ Cm,
                                                          const double* __restrict__ p, double* __restrict__ Ap, const int* __restrict__ done) {
  if (*done) return;
  __shared__ double red[8][9];
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int a = tid; a < n_img; a += blockDim.x) {
    const double* Ba = Bm + ((size_t)a * ncb + c) * 54;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const double pq = p[6 * (size_t)a + q];
#pragma unroll
      for (int m = 0; m < 9; ++m) acc[m] += Ba[9 * q + m] * pq;
    }
  }
#pragma unroll
  for (int m = 0; m < 9; ++m) { const double s = warp_sum(acc[m]); if (lane == 0) red[w][m] = s; }
  __syncthreads();
  if (tid < 9) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += red[k][tid];
    const double* pi = p + 6 * (size_t)n_img;
    const int n9 = 9 * ncb;
    for (int k = 0; k < n9; ++k) s += Cm[(size_t)(9 * c + tid) * n9 + k] * pi[k];
    Ap[6 * (size_t)n_img + 9 * c + tid] = s;
  }
}

}  // namespace mm
