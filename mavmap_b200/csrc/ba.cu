// ba.cu — bundle-adjustment session: structure setup on the device (CUB sorts/scans), the
// Levenberg-Marquardt driver and the C ABI (mm_ba_*, mm_pose_refine).
//
// Replaces (mavmap/mavmap):  bundle_adjustment()  src/base3d/bundle_adjustment.cc:449-613 and
// pose_refinement() :139-225 — i.e. the ceres::Problem construction + ceres::Solve(SPARSE_SCHUR).
// The LM control flow restates Ceres 1.8 (trust_region_minimizer.cc,
// levenberg_marquardt_strategy.cc; rules listed in SURVEY.md §8a-a3' and oracle/orc_ba.c); the
// linear solve is PCG on the explicitly assembled reduced camera system, preconditioned by its exact sparse tile Cholesky
// (tilechol.cuh; two-level aggregate preconditioner as the fallback, dense Cholesky in one CTA up to 26 images).
// The host only sequences kernels and reads a handful of scalars per LM iteration; the session setup (structure, masks,
// renumbering, index check) runs on the device and the symbolic analysis of the factorisation on a host thread beside it.
#include <cub/cub.cuh>
#include <math.h>
#include <vector>
#include <string>
#include <algorithm>
#include <thread>
#include <time.h>
#include "ba_kernels.cuh"
#include "ba_pose.cuh"
#include "tilechol.cuh"

namespace mm {

// ------------------------------------------------------------------ multi-GPU shard ranges
// both lists are sorted by point inside an image / a block (stable sorts of point-major data), so a rank's part is a
// contiguous sub-range: found once at setup by bisection
__global__ void k_shard_cam_ranges(int n_img, const int* __restrict__ cam_start, const int* __restrict__ cam_perm, int o_lo, int o_hi,
                                   int* __restrict__ lo, int* __restrict__ hi) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_img) return;
  const int b = cam_start[w], e = cam_start[w + 1];
  int l = b, r = e;
  while (l < r) { const int m = (l + r) >> 1; if (cam_perm[m] < o_lo) l = m + 1; else r = m; }
  lo[w] = l; r = e;
  while (l < r) { const int m = (l + r) >> 1; if (cam_perm[m] < o_hi) l = m + 1; else r = m; }
  hi[w] = l;
}
__global__ void k_shard_blk_ranges(int64_t nblk, const int* __restrict__ bp_start, const int* __restrict__ bp_end, const int* __restrict__ sp_pt,
                                   int p_lo, int p_hi, int* __restrict__ lo, int* __restrict__ hi) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= nblk) return;
  int l = bp_start[b], r = bp_end[b]; const int e = r;
  while (l < r) { const int m = (l + r) >> 1; if (sp_pt[m] < p_lo) l = m + 1; else r = m; }
  lo[b] = l; r = e;
  while (l < r) { const int m = (l + r) >> 1; if (sp_pt[m] < p_hi) l = m + 1; else r = m; }
  hi[b] = l;
}
// scal[0] = local cost, scal[1] = local |x|^2, scal[2] = local failure flag (slots 3.. hold the per-rank gradient maxima)
__global__ void k_pack_scal(const double* __restrict__ loc, const int* __restrict__ fail, double* __restrict__ scal) {
  if (threadIdx.x == 0) { scal[0] = loc[0]; scal[1] = loc[5]; scal[2] = *fail ? 1.0 : 0.0; }
}
__global__ void k_pack_step(const double* __restrict__ loc, const int* __restrict__ fail, double* __restrict__ buf) {
  if (threadIdx.x == 0) { buf[0] = loc[1]; buf[1] = loc[3]; buf[2] = loc[4]; buf[3] = *fail ? 1.0 : 0.0; }
}
__global__ void k_unpack_step(const double* __restrict__ buf, double* __restrict__ red, int* __restrict__ fail, const double* __restrict__ prior_cost) {
  if (threadIdx.x == 0) { red[1] = buf[0] + (prior_cost ? prior_cost[0] : 0.0); red[3] = buf[1]; red[4] = buf[2]; if (buf[3] > 0.0) *fail = 1; }
}

// ------------------------------------------------------------------ setup kernels
__global__ void k_iota(int n, int* out) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = i; }
__global__ void k_hist(int64_t n, const int* __restrict__ keys, int* __restrict__ cnt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) atomicAdd(cnt + keys[i], 1);
}
__global__ void k_gather_img(int64_t n, const int* __restrict__ perm, const int* __restrict__ img_in, int* __restrict__ img) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) img[i] = img_in[perm[i]];
}
__global__ void k_gather_xy(int64_t n, const int* __restrict__ perm, const double2* __restrict__ xy_in, double2* __restrict__ xy) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) xy[i] = xy_in[perm[i]];
}
__global__ void k_pair_count(int n_pt, const int* __restrict__ pt_start, int64_t* __restrict__ cnt) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pt) { const int64_t k = pt_start[p + 1] - pt_start[p]; cnt[p] = k * (k - 1) / 2; }
}
__global__ void k_pair_keys(int n_pt, int n_img, const int* __restrict__ pt_start, const int* __restrict__ obs_img,
                            const int64_t* __restrict__ pair_off, unsigned long long* __restrict__ keys,
                            int* __restrict__ pair_lo, int* __restrict__ pair_hi) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pt) return;
  int64_t slot = pair_off[p];
  const int o0 = pt_start[p], o1 = pt_start[p + 1];
  for (int i = o0; i < o1 - 1; ++i) {
    const int a = obs_img[i];
    for (int j = i + 1; j < o1; ++j, ++slot) {
      const int b = obs_img[j];
      const unsigned long long lo = a < b ? a : b, hi = a < b ? b : a;
      keys[slot] = lo * (unsigned long long)n_img + hi;
      pair_lo[slot] = a <= b ? i : j; pair_hi[slot] = a <= b ? j : i;      // observation of the lower / higher image index
    }
  }
}
// head flag of every run of equal off-diagonal keys
__global__ void k_pair_heads(int64_t n, int n_img, const unsigned long long* __restrict__ keys, int* __restrict__ head) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    const bool diag = (k / n_img) == (k % n_img);
    head[i] = (!diag && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
  }
}
// block id of every (sorted) pair; per block the contiguous range of its pairs in sorted order and the pairs themselves
__global__ void k_pair_assign(int64_t n, int n_img, const unsigned long long* __restrict__ keys, const int* __restrict__ incl,
                              const int* __restrict__ head, const int* __restrict__ slot,
                              int* __restrict__ blk_a, int* __restrict__ blk_b,
                              const int* __restrict__ pair_lo, const int* __restrict__ pair_hi, int* __restrict__ sp_lo, int* __restrict__ sp_hi,
                              const int* __restrict__ obs_pt, int* __restrict__ sp_pt, int* __restrict__ bp_start, int* __restrict__ bp_end) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    const int a = (int)(k / n_img), b = (int)(k % n_img);
    const int sl = slot[i];
    const int id = (a == b) ? a : n_img + incl[i] - 1;
    if (a != b && head[i]) { blk_a[incl[i] - 1] = a; blk_b[incl[i] - 1] = b; }
    sp_lo[i] = pair_lo[sl]; sp_hi[i] = pair_hi[sl]; sp_pt[i] = obs_pt[pair_lo[sl]];
    if (i == 0 || keys[i - 1] != k) bp_start[id] = (int)i;
    if (i == n - 1 || keys[i + 1] != k) bp_end[id] = (int)(i + 1);
  }
}
// CSR entries: key = row * n_img + col, value = signed block reference (negative = transposed)
__global__ void k_csr_entries(int n_img, int n_off, const int* __restrict__ blk_a, const int* __restrict__ blk_b,
                              unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_img) { keys[i] = (unsigned long long)i * n_img + i; vals[i] = i; }
  if (i < n_off) {
    const int a = blk_a[i], b = blk_b[i], id = n_img + i;
    keys[n_img + 2 * (size_t)i] = (unsigned long long)a * n_img + b; vals[n_img + 2 * (size_t)i] = id;
    keys[n_img + 2 * (size_t)i + 1] = (unsigned long long)b * n_img + a; vals[n_img + 2 * (size_t)i + 1] = -id - 1;
  }
}
__global__ void k_csr_rows(int64_t n, int n_img, const unsigned long long* __restrict__ keys, int* __restrict__ col, int* __restrict__ cnt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    col[i] = (int)(k % n_img);
    atomicAdd(cnt + (int)(k / n_img), 1);
  }
}
// spatial proxy key of a point: (smallest, largest) index of the images that see it — images follow the flight path
__global__ void k_pt_minmax(int64_t n, const int* __restrict__ img, const int* __restrict__ pt, int* __restrict__ mn, int* __restrict__ mx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) { atomicMin(mn + pt[i], img[i]); atomicMax(mx + pt[i], img[i]); }
}
__global__ void k_pt_key(int n_pt, int n_img, const int* __restrict__ mn, const int* __restrict__ mx, unsigned long long* __restrict__ key) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pt) key[p] = (unsigned long long)mn[p] * (unsigned long long)(n_img + 1) + (unsigned long long)(mx[p] < 0 ? n_img : mx[p]);
}
__global__ void k_invert_perm(int n, const int* __restrict__ new2old, int* __restrict__ old2new) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) old2new[new2old[i]] = i;
}
__global__ void k_remap(int64_t n, const int* __restrict__ old2new, int* __restrict__ pt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) pt[i] = old2new[pt[i]];
}
__global__ void k_fill_int(int n, int* p, int v) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
__global__ void k_fill(int64_t n, double* p, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// range check of the observation indices on the device (the host does not walk the 10^7 observations): flag = 1 if any is out of range
__global__ void k_obs_check(int64_t n, const int* __restrict__ img, const int* __restrict__ pt, int n_img, int n_pt, int* __restrict__ flag) {
  bool bad = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bad |= (unsigned)img[i] >= (unsigned)n_img || (unsigned)pt[i] >= (unsigned)n_pt;
  if (bad) *flag = 1;
}
// masks: 1.0 = free and present in at least one residual block.  pose_const: 4 bytes per image (rvec, tx, ty, tz)
__global__ void k_pose_mask(int n_img, const unsigned char* __restrict__ pose_const, const int* __restrict__ cam_start, double* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img) return;
  const bool seen = cam_start[i + 1] > cam_start[i];
#pragma unroll
  for (int k = 0; k < 6; ++k) mask[6 * (size_t)i + k] = (seen && !pose_const[4 * (size_t)i + (k < 3 ? 0 : k - 2)]) ? 1.0 : 0.0;
}
__global__ void k_pt_mask(int n_pt, const unsigned char* __restrict__ pt_const /*caller order*/, const int* __restrict__ new2old, const int* __restrict__ pt_start, double* __restrict__ mask) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pt) mask[p] = (pt_start[p + 1] > pt_start[p] && !pt_const[new2old[p]]) ? 1.0 : 0.0;
}
// rows of 1 or 3 doubles between the caller's point order and the internal one
__global__ void k_gather_rows3(int n, const int* __restrict__ new2old, const double* __restrict__ src /*caller order*/, double* __restrict__ dst) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) { const size_t o = (size_t)new2old[p]; dst[3 * (size_t)p] = src[3 * o]; dst[3 * (size_t)p + 1] = src[3 * o + 1]; dst[3 * (size_t)p + 2] = src[3 * o + 2]; }
}
__global__ void k_scatter_rows3(int n, const int* __restrict__ new2old, const double* __restrict__ src /*internal order*/, double* __restrict__ dst) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) { const size_t o = (size_t)new2old[p]; dst[3 * o] = src[3 * (size_t)p]; dst[3 * o + 1] = src[3 * (size_t)p + 1]; dst[3 * o + 2] = src[3 * (size_t)p + 2]; }
}
__global__ void k_gather_rows1(int n, const int* __restrict__ new2old, const double* __restrict__ src, double* __restrict__ dst) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x; if (p < n) dst[p] = src[new2old[p]];
}
__global__ void k_scatter_rows1(int n, const int* __restrict__ new2old, const double* __restrict__ src, double* __restrict__ dst) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x; if (p < n) dst[new2old[p]] = src[p];
}

static inline int blocks_for(int64_t n, int block) { int64_t g = (n + block - 1) / block; return (int)(g < 1 ? 1 : g); }
static inline int grid_stride(int64_t n, int block) { int64_t g = (n + block - 1) / block; const int cap = num_sms() * 8; return (int)(g < 1 ? 1 : (g > cap ? cap : g)); }

}  // namespace mm

using namespace mm;

// ------------------------------------------------------------------ session
struct mm_ba_session {
  cudaStream_t stream = nullptr;
  mm_ba_options opt;
  int n_img = 0, n_cam = 0, n_pt = 0; int64_t n_obs = 0;
  int n_off = 0; int64_t nblk = 0, n_pairs = 0, n_ent = 0;
  std::vector<double> h_poses0, h_intr0, h_pts0;
  DevBuf<int> obs_perm;                                                 // setup only: point-sorted position -> caller's observation index (until the image coordinates are in)
  DevBuf<int> pt_new2old;                                               // internal point order (spatially clustered) -> caller's order
  bool keep_host_pts = true;                                            // mm_ba_solve: no session reset, so no host copy of the points
  DevBuf<double2> obs_xy; DevBuf<int> obs_img, obs_pt, pt_start, cam_perm, cam_start, img_cam, cam_model;
  DevBuf<int64_t> pair_off; DevBuf<int> row_start, row_col, row_blk, sp_lo, sp_hi, sp_pt, bp_start, bp_end; DevBuf<double> pinfo;
  DevBuf<double> S, Minv, poses, intr, pts, poses2, pts2, aux, aux2, rec;
  DevBuf<double> pose_mask, pt_mask, scale_c, scale_p, Vinv, gp, dp, gc, dc, rhs;
  DevBuf<double> vx, vr, vz, vp0, vp1, vAp, pcg_sc; DevBuf<int> pcg_ic;
  DevBuf<double> part_cost, part_pt, part_cam, part_x, red;   // red: [0]=cost [1]=new_cost [2]=gmax [3]=step_norm2 [4]=mcc [5]=xnorm2
  DevBuf<int> fail; DevBuf<unsigned long long> pcg_dbg;
  // multi-GPU: the points [p_lo, p_hi) (internal order) and with them the observations [o_lo, o_hi) belong to this rank;
  // poses, intrinsics and the reduced system are replicated; `ar` sums device doubles over the ranks (one exchange step
  // per Schur assembly, one per step evaluation).  world == 1: everything is local and `ar` is never called.
  int rank = 0, world = 1; mm_allreduce_fn ar = nullptr; void* ar_user = nullptr;
  int p_lo = 0, p_hi = 0; int64_t o_lo = 0, o_hi = 0;
  DevBuf<double> xch, loc, stepbuf, ud; size_t xch_count = 0, scal_off = 0;     // xch = [S | rhs | gc | ud | scal]: what the exchange step sums
  DevBuf<int> cam_lo, cam_hi, bp_lo, bp_hi;                                      // this rank's part of every image's / block's list
  int np_loc() const { return p_hi - p_lo; }
  int64_t no_loc() const { return o_hi - o_lo; }
  double* rec_base() const { return rec.p - (size_t)REC * (size_t)o_lo; }      // records are stored for the local observations only
  double* scal() const { return xch.p + scal_off; }
  // rotation constraints (constrain_rotation): rvec0 and weight per image, residual / Jacobian of the current iterate
  int n_prior = 0; DevBuf<double> pr_rot0, pr_w, pr_r, pr_J;
  // coarse level of the two-level preconditioner (ba_coarse.cuh)
  int cm = 0, n_agg = 0; std::vector<int> h_agg; std::vector<double> h_pose_mask; int coarse_iter = -1; double coarse_radius = 0.0;
  DevBuf<int> blk_a, blk_b, agg; DevBuf<double> Pc, Ac, gjC, gjR, crc, cqc, cyc; int gj_grid = 0;
  // refined intrinsics (single shared camera)
  bool refine = false;
  DevBuf<double> ji, intr2, intr_mask, scale_i, Apc, Bm, intr_acc, xch2, gi, di, xi, img_sq; DevBuf<unsigned> pt_cam_mask; size_t xch2_count = 0;
  double* ji_base() const { return ji.p - (size_t)18 * (size_t)o_lo; }      // like rec_base(): intrinsics Jacobians exist for the local observations only
  // sparse tile Cholesky preconditioner of the PCG solve (tilechol.cuh): plan (host), its device copy, tiles, PCG vectors
  bool tc_on = false; int tc_epoch = 0, tc_sepoch = 0, tc_grid_f = 0, tc_grid_s = 0, n_unk = 0, ncb = 0;
  TileCholPlan tc_plan; TcDev tc;
  // the symbolic analysis runs on a host thread beside the rest of the setup (tc_begin / tc_finish)
  std::thread tc_thread; int tc_plan_rc = 0; std::vector<int> tc_ha, tc_hb; std::vector<double> tc_pos;
  std::vector<int> tc_h_sched, tc_h_sdesc; std::vector<int4> tc_h_upd; std::vector<int2> tc_h_items; bool tc_force = false;
  DevBuf<int> tc_unk_of, tc_sc_tile, tc_sc_off, tc_a_tiles, tc_img_tile, tc_img_slot, tc_sched, tc_sdesc;
  DevBuf<int4> tc_upd; DevBuf<int2> tc_items;
  DevBuf<int> tc_ready, tc_sflag, tc_counters;
  DevBuf<int64_t> tc_col_ptr;
  DevBuf<double> tc_Ls;
  DevBuf<double> tc_L, tc_WC, tc_WR, tc_slots, tc_invd, dv_b, dv_r, dv_z, dv_p, dv_Ap;
  int grid_obs = 1, grid_pt = 1, grid_cam6 = 1, grid_x = 1, pcg_grid = 0, pcg_ecap = 0, pcg_threads = 0; bool pcg_cached = false; size_t pcg_smem = 0; const void* pcg_fn = nullptr;
  cudaEvent_t evs[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // LM state (host)
  double cost = 0, radius = 0, decrease_factor = 2, x_norm = 0, abs_gtol = 0, gmax = 0;
  int iter = 0, n_invalid = 0; bool started = false, finished = false, scaled = false;
  int last_pcg_iters = 0;
  mm_ba_summary sum;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

struct Timer {
  mm_ba_session* s; double* acc;
  Timer(mm_ba_session* s_, double* a) : s(s_), acc(a) { cudaEventRecord(s->ev0, s->stream); }
  ~Timer() { cudaEventRecord(s->ev1, s->stream); cudaEventSynchronize(s->ev1); float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); *acc += ms; }
};

// host wall-clock stamps of the phases of mm_ba_session_create (MM_SETUP_TIMING=1 prints them; the stream is drained at every stamp)
struct SetupClock {
  bool on; cudaStream_t st; double t0, last;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return 1e3 * ts.tv_sec + 1e-6 * ts.tv_nsec; }
  explicit SetupClock(cudaStream_t s) : on(getenv("MM_SETUP_TIMING") != nullptr), st(s) { t0 = last = now(); }
  void mark(const char* what) { if (!on) return; cudaStreamSynchronize(st); const double t = now(); fprintf(stderr, "[setup] %-28s %8.3f ms  (at %8.3f)\n", what, t - last, t - t0); last = t; }
};

// Large host -> device copies of pageable caller memory (mm_ba_solve hands over 10^7 observations = 240 MB): the driver's own
// staging moves them at ~10 GB/s because one thread feeds it.  Here several host threads copy slices of a chunk into one half of
// a persistent pinned double buffer while the DMA engine drains the other half.  Stream-ordered like cudaMemcpyAsync; the source
// has been read completely when the call returns.
struct Staging {
  static constexpr size_t CHUNK = (size_t)32 << 20;
  char* buf[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr}; bool ok = false; int dev = -1;
  static Staging& get() { static thread_local Staging st; return st; }
  bool ready() {
    int d = 0; cudaGetDevice(&d);
    if (ok && d == dev) return true;
    if (ok) return false;                       // created for another device: fall back to plain copies
    for (int b = 0; b < 2; ++b)
      if (cudaHostAlloc((void**)&buf[b], CHUNK, cudaHostAllocDefault) != cudaSuccess || cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
    ok = true; dev = d; return true;
  }
};
cudaError_t upload_async(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  Staging& S = Staging::get();
  if (bytes < ((size_t)4 << 20) || getenv("MM_NO_STAGING") || !S.ready()) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  static const unsigned n_thr = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
  size_t off = 0; int b = 0;
  while (off < bytes) {
    const size_t len = std::min(Staging::CHUNK, bytes - off);
    cudaError_t e = cudaEventSynchronize(S.ev[b]); if (e != cudaSuccess) return e;          // the previous copy out of this half has finished
    const char* from = static_cast<const char*>(src) + off; char* to = S.buf[b];
    const size_t slice = ((len + n_thr - 1) / n_thr + 4095) & ~(size_t)4095;
    std::thread workers[8]; unsigned nw = 0; size_t done_to = std::min(slice, len);
    for (unsigned t = 1; t < n_thr; ++t) {
      const size_t o = t * slice; if (o >= len) break;
      try { workers[nw] = std::thread([=] { memcpy(to + o, from + o, std::min(slice, len - o)); }); ++nw; done_to = std::min(o + slice, len); }
      catch (...) { break; }                      // no more threads to be had: this thread copies the rest
    }
    memcpy(to, from, std::min(slice, len));
    if (done_to < len) memcpy(to + done_to, from + done_to, len - done_to);
    for (unsigned t = 0; t < nw; ++t) workers[t].join();
    e = cudaMemcpyAsync(static_cast<char*>(dst) + off, to, len, cudaMemcpyHostToDevice, st); if (e != cudaSuccess) return e;
    e = cudaEventRecord(S.ev[b], st); if (e != cudaSuccess) return e;
    off += len; b ^= 1;
  }
  return cudaSuccess;
}

// One pool allocation for a group of temporaries (the structure setup used ~20 separate ones of 40 - 160 MB each; their
// alloc / free pattern fragmented the stream-ordered pool from call to call).  take() hands out 256-byte aligned pieces.
struct Arena {
  DevBuf<char> mem; size_t used = 0, cap = 0;
  static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
  cudaError_t reserve(size_t bytes) { used = 0; cap = bytes; return mem.alloc(std::max<size_t>(bytes, 256)); }
  template <typename T> T* take(size_t count) { T* q = reinterpret_cast<T*>(mem.p + used); used += pad(count * sizeof(T)); return used <= cap ? q : nullptr; }
};

thread_local SetupClock* g_setup_clock = nullptr;
#define SETUP_MARK(what) do { if (g_setup_clock) g_setup_clock->mark(what); } while (0)

int validate_problem(const mm_ba_problem* P) {
  if (!P || P->n_img < 0 || P->n_cam < 0 || P->n_pt < 0 || P->n_obs < 0) { set_error("invalid problem sizes"); return MM_ERR_INVALID_ARG; }
  if (P->n_obs >= (int64_t)1 << 31) { set_error("n_obs >= 2^31 not supported"); return MM_ERR_UNSUPPORTED; }
  // every array is checked against its own count: an empty-observation problem with images / cameras / points still dereferences them
  if ((P->n_img > 0 && (!P->poses || !P->pose_const || !P->img_cam)) || (P->n_cam > 0 && (!P->intr || !P->cam_model || !P->intr_const)) ||
      (P->n_pt > 0 && (!P->pts || !P->pt_const)) || (P->n_obs > 0 && (!P->obs_xy || !P->obs_img || !P->obs_pt))) { set_error("null array in problem"); return MM_ERR_INVALID_ARG; }
  for (int c = 0; c < P->n_cam; ++c) if (model_num_params(P->cam_model[c]) < 0) { set_error("unknown camera model code %d", P->cam_model[c]); return MM_ERR_INVALID_ARG; }
  for (int i = 0; i < P->n_img; ++i) if (P->img_cam[i] < 0 || P->img_cam[i] >= P->n_cam) { set_error("img_cam out of range"); return MM_ERR_INVALID_ARG; }
  if (P->n_obs > 0 && (P->n_img == 0 || P->n_pt == 0)) { set_error("observation index out of range"); return MM_ERR_INVALID_ARG; }
  return MM_OK;       // (the observation indices themselves are range-checked on the device, build_structure)
}

template <typename K, typename V>
size_t sort_bytes(int64_t n, int end_bit) {
  size_t b = 0;
  cub::DeviceRadixSort::SortPairs((void*)nullptr, b, (const K*)nullptr, (K*)nullptr, (const V*)nullptr, (V*)nullptr, (int)std::max<int64_t>(n, 1), 0, end_bit);
  return b;
}
template <typename K, typename V>
int sort_pairs(cudaStream_t st, void* tmp, size_t bytes, const K* kin, K* kout, const V* vin, V* vout, int64_t n, int end_bit) {
  MM_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, (int)n, 0, end_bit, st));
  count_launch(4);
  return MM_OK;
}
template <typename T>
size_t scan_bytes(int64_t n) { size_t b = 0; cub::DeviceScan::ExclusiveSum((void*)nullptr, b, (const T*)nullptr, (T*)nullptr, (int)std::max<int64_t>(n, 1)); size_t c = 0; cub::DeviceScan::InclusiveSum((void*)nullptr, c, (const T*)nullptr, (T*)nullptr, (int)std::max<int64_t>(n, 1)); return std::max(b, c); }
template <typename T>
int exclusive_sum(cudaStream_t st, void* tmp, size_t bytes, const T* in, T* out, int64_t n) {
  MM_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n, st));
  count_launch(2);
  return MM_OK;
}
template <typename T>
int inclusive_sum(cudaStream_t st, void* tmp, size_t bytes, const T* in, T* out, int64_t n) {
  MM_CUDA(cub::DeviceScan::InclusiveSum(tmp, bytes, in, out, (int)n, st));
  count_launch(2);
  return MM_OK;
}
static int bits_for(unsigned long long maxval) { int b = 1; while (b < 64 && (maxval >> b)) ++b; return b; }

// Structure of the problem on the device.  Everything that walks the observations runs here (range check of the indices, the
// masks of parameters without residuals, the point renumbering): the host only hands the arrays over.
int build_structure(mm_ba_session* s, const mm_ba_problem* P) {
  cudaStream_t st = s->stream;
  const int64_t n = s->n_obs; const int n_img = s->n_img, n_pt = s->n_pt;
  const int B = 256;
  const size_t n1 = (size_t)std::max<int64_t>(n, 1), np1 = (size_t)n_pt + 1, ni1 = (size_t)n_img + 1;
  const int bits_pt = bits_for((unsigned long long)std::max(n_pt, 1)), bits_img = bits_for((unsigned long long)std::max(n_img, 1));
  const int bits_key = bits_for((unsigned long long)(n_img + 1) * (unsigned long long)(n_img + 1));
  MM_CUDA(s->obs_xy.alloc(n1)); MM_CUDA(s->obs_img.alloc(n1)); MM_CUDA(s->obs_pt.alloc(n1));
  MM_CUDA(s->pt_start.alloc(np1)); MM_CUDA(s->cam_start.alloc(ni1)); MM_CUDA(s->cam_perm.alloc(n1)); MM_CUDA(s->pt_new2old.alloc((size_t)std::max(n_pt, 1)));
  MM_CUDA(s->pair_off.alloc(np1));
  // ---- arena 1: temporaries at observation / point granularity
  Arena A1;
  size_t cub1 = std::max(std::max(sort_bytes<int, int>(n, bits_pt), sort_bytes<int, int>(n, bits_img)), sort_bytes<unsigned long long, int>(n_pt, bits_key));
  cub1 = std::max(cub1, std::max(scan_bytes<int>((int64_t)np1), scan_bytes<int64_t>((int64_t)np1)));
  { size_t need = 0;
    const size_t sizes[] = { 4 * n1, 4 * n1, 4 * n1, 4 * n1, 4 * np1, 4 * ni1, 8 * np1, 4 * np1, 4 * np1, 4 * np1, 4 * np1, 8 * np1, 8 * np1, np1, 4 * ni1, 256, cub1 };
    for (size_t b : sizes) need += Arena::pad(b);
    MM_CUDA(A1.reserve(need)); }
  int* img_in = A1.take<int>(n1); int* pt_in = A1.take<int>(n1); int* iota = A1.take<int>(n1);
  MM_CUDA(s->obs_perm.alloc(n1)); int* perm = s->obs_perm.p;
  int* keys_out = A1.take<int>(n1); int* cnt_pt = A1.take<int>(np1); int* cnt_img = A1.take<int>(ni1); int64_t* cnt64 = A1.take<int64_t>(np1);
  int* mn = A1.take<int>(np1); int* mx = A1.take<int>(np1); int* ids = A1.take<int>(np1); int* old2new = A1.take<int>(np1);
  unsigned long long* key = A1.take<unsigned long long>(np1); unsigned long long* key_s = A1.take<unsigned long long>(np1);
  unsigned char* pt_const = A1.take<unsigned char>(np1); unsigned char* pose_const = A1.take<unsigned char>(4 * ni1); int* flag = A1.take<int>(64);
  void* cub_tmp = A1.take<char>(cub1);
  if (!cub_tmp) { set_error("internal: setup arena too small"); return MM_ERR_ALLOC; }
  // upload in caller order
  if (n > 0) {
    MM_CUDA(upload_async(img_in, P->obs_img, sizeof(int) * n, st));
    MM_CUDA(upload_async(pt_in, P->obs_pt, sizeof(int) * n, st));
  }
  if (n_pt > 0) MM_CUDA(upload_async(pt_const, P->pt_const, (size_t)n_pt, st));
  if (n_img > 0) MM_CUDA(cudaMemcpyAsync(pose_const, P->pose_const, 4 * (size_t)n_img, cudaMemcpyHostToDevice, st));
  MM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  if (n > 0) {
    k_obs_check<<<grid_stride(n, B), B, 0, st>>>(n, img_in, pt_in, n_img, n_pt, flag); MM_LAUNCH_CHECK();
    int bad = 0;
    MM_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, st)); MM_CUDA(cudaStreamSynchronize(st));
    if (bad) { set_error("observation index out of range"); return MM_ERR_INVALID_ARG; }
  }
  SETUP_MARK("structure: upload + check");
  // 0. renumber the points so that points seen by the same images are neighbours: gives the per-image gathers of
  //    K1/K2b/K4 cache locality.  The caller's numbering is restored on download (pt_new2old stays on the device).
  if (n > 0 && n_pt > 1 && !getenv("MM_BA_NO_REORDER")) {
    k_fill_int<<<blocks_for(n_pt, B), B, 0, st>>>(n_pt, mn, n_img); MM_LAUNCH_CHECK();
    k_fill_int<<<blocks_for(n_pt, B), B, 0, st>>>(n_pt, mx, -1); MM_LAUNCH_CHECK();
    k_pt_minmax<<<grid_stride(n, B), B, 0, st>>>(n, img_in, pt_in, mn, mx); MM_LAUNCH_CHECK();
    k_pt_key<<<blocks_for(n_pt, B), B, 0, st>>>(n_pt, n_img, mn, mx, key); MM_LAUNCH_CHECK();
    k_iota<<<blocks_for(n_pt, B), B, 0, st>>>(n_pt, ids); MM_LAUNCH_CHECK();
    int rc = sort_pairs<unsigned long long, int>(st, cub_tmp, cub1, key, key_s, ids, s->pt_new2old.p, n_pt, bits_key); if (rc) return rc;
    k_invert_perm<<<blocks_for(n_pt, B), B, 0, st>>>(n_pt, s->pt_new2old.p, old2new); MM_LAUNCH_CHECK();
    k_remap<<<grid_stride(n, B), B, 0, st>>>(n, old2new, pt_in); MM_LAUNCH_CHECK();
  } else if (n_pt > 0) {
    k_iota<<<blocks_for(n_pt, B), B, 0, st>>>(n_pt, s->pt_new2old.p); MM_LAUNCH_CHECK();
  }
  if (n > 0) {
    k_iota<<<blocks_for(n, B), B, 0, st>>>((int)n, iota); MM_LAUNCH_CHECK();
    // 1. stable sort by point (keeps the caller's order inside a track)
    int rc = sort_pairs<int, int>(st, cub_tmp, cub1, pt_in, s->obs_pt.p, iota, perm, n, bits_pt); if (rc) return rc;
    k_gather_img<<<grid_stride(n, B), B, 0, st>>>(n, perm, img_in, s->obs_img.p); MM_LAUNCH_CHECK();
  }
  // 2. point -> observation CSR
  MM_CUDA(cudaMemsetAsync(cnt_pt, 0, sizeof(int) * np1, st));
  if (n > 0) { k_hist<<<grid_stride(n, B), B, 0, st>>>(n, s->obs_pt.p, cnt_pt); MM_LAUNCH_CHECK(); }
  { int rc = exclusive_sum<int>(st, cub_tmp, cub1, cnt_pt, s->pt_start.p, (int64_t)np1); if (rc) return rc; }
  // 3. camera-sorted permutation of the (point-sorted) observation positions
  MM_CUDA(cudaMemsetAsync(cnt_img, 0, sizeof(int) * ni1, st));
  if (n > 0) {
    k_hist<<<grid_stride(n, B), B, 0, st>>>(n, s->obs_img.p, cnt_img); MM_LAUNCH_CHECK();
    int rc = sort_pairs<int, int>(st, cub_tmp, cub1, s->obs_img.p, keys_out, iota, s->cam_perm.p, n, bits_img); if (rc) return rc;
  }
  { int rc = exclusive_sum<int>(st, cub_tmp, cub1, cnt_img, s->cam_start.p, (int64_t)ni1); if (rc) return rc; }
  // masks: 1.0 = free and present in at least one residual block
  if (n_img > 0) { k_pose_mask<<<blocks_for(n_img, B), B, 0, st>>>(n_img, pose_const, s->cam_start.p, s->pose_mask.p); MM_LAUNCH_CHECK(); }
  if (n_pt > 0) { k_pt_mask<<<blocks_for(n_pt, B), B, 0, st>>>(n_pt, pt_const, s->pt_new2old.p, s->pt_start.p, s->pt_mask.p); MM_LAUNCH_CHECK(); }
  // 4. pairs of observations of one point -> blocks of the reduced camera system
  MM_CUDA(cudaMemsetAsync(cnt64, 0, sizeof(int64_t) * np1, st));
  if (n_pt > 0) { k_pair_count<<<blocks_for(n_pt, B), B, 0, st>>>(n_pt, s->pt_start.p, cnt64); MM_LAUNCH_CHECK(); }
  { int rc = exclusive_sum<int64_t>(st, cub_tmp, cub1, cnt64, s->pair_off.p, (int64_t)np1); if (rc) return rc; }
  int64_t n_pairs = 0;
  MM_CUDA(cudaMemcpyAsync(&n_pairs, s->pair_off.p + n_pt, sizeof(int64_t), cudaMemcpyDeviceToHost, st)); MM_CUDA(cudaStreamSynchronize(st));
  if (n_pairs >= ((int64_t)1 << 31)) { set_error("too many observation pairs (%lld)", (long long)n_pairs); return MM_ERR_UNSUPPORTED; }
  s->n_pairs = n_pairs;
  A1.mem.release();
  SETUP_MARK("structure: sorts + CSR");
  DevBuf<int>& blk_a = s->blk_a; DevBuf<int>& blk_b = s->blk_b;
  int n_off = 0;
  MM_CUDA(s->sp_lo.alloc((size_t)n_pairs)); MM_CUDA(s->sp_hi.alloc((size_t)n_pairs)); MM_CUDA(s->sp_pt.alloc((size_t)n_pairs));
  if (n_pairs > 0) {
    Arena A2;
    const int bits_pair = bits_for((unsigned long long)n_img * (unsigned long long)n_img);
    const size_t cub2 = std::max(sort_bytes<unsigned long long, int>(n_pairs, bits_pair), scan_bytes<int>(n_pairs));
    const size_t np = (size_t)n_pairs;
    MM_CUDA(A2.reserve(2 * Arena::pad(8 * np) + 6 * Arena::pad(4 * np) + Arena::pad(cub2)));
    unsigned long long* keys = A2.take<unsigned long long>(np); unsigned long long* keys_s = A2.take<unsigned long long>(np);
    int* slot = A2.take<int>(np); int* slot_s = A2.take<int>(np); int* head = A2.take<int>(np); int* incl = A2.take<int>(np); int* pair_lo = A2.take<int>(np); int* pair_hi = A2.take<int>(np);
    void* tmp2 = A2.take<char>(cub2);
    if (!tmp2) { set_error("internal: setup arena too small"); return MM_ERR_ALLOC; }
    k_pair_keys<<<blocks_for(n_pt, B), B, 0, st>>>(n_pt, n_img, s->pt_start.p, s->obs_img.p, s->pair_off.p, keys, pair_lo, pair_hi); MM_LAUNCH_CHECK();
    k_iota<<<blocks_for(n_pairs, B), B, 0, st>>>((int)n_pairs, slot); MM_LAUNCH_CHECK();
    int rc = sort_pairs<unsigned long long, int>(st, tmp2, cub2, keys, keys_s, slot, slot_s, n_pairs, bits_pair); if (rc) return rc;
    k_pair_heads<<<grid_stride(n_pairs, B), B, 0, st>>>(n_pairs, n_img, keys_s, head); MM_LAUNCH_CHECK();
    rc = inclusive_sum<int>(st, tmp2, cub2, head, incl, n_pairs); if (rc) return rc;
    MM_CUDA(cudaMemcpyAsync(&n_off, incl + (n_pairs - 1), sizeof(int), cudaMemcpyDeviceToHost, st)); MM_CUDA(cudaStreamSynchronize(st));
    MM_CUDA(blk_a.alloc((size_t)n_off)); MM_CUDA(blk_b.alloc((size_t)n_off));
    MM_CUDA(s->bp_start.alloc((size_t)n_img + n_off)); MM_CUDA(s->bp_end.alloc((size_t)n_img + n_off));
    MM_CUDA(cudaMemsetAsync(s->bp_start.p, 0, sizeof(int) * ((size_t)n_img + n_off), st)); MM_CUDA(cudaMemsetAsync(s->bp_end.p, 0, sizeof(int) * ((size_t)n_img + n_off), st));
    k_pair_assign<<<grid_stride(n_pairs, B), B, 0, st>>>(n_pairs, n_img, keys_s, incl, head, slot_s, blk_a.p, blk_b.p,
        pair_lo, pair_hi, s->sp_lo.p, s->sp_hi.p, s->obs_pt.p, s->sp_pt.p, s->bp_start.p, s->bp_end.p); MM_LAUNCH_CHECK();
    MM_CUDA(cudaStreamSynchronize(st));
  }
  if (!s->bp_start.p) { MM_CUDA(s->bp_start.alloc((size_t)std::max(n_img, 1))); MM_CUDA(s->bp_end.alloc((size_t)std::max(n_img, 1)));
    MM_CUDA(cudaMemsetAsync(s->bp_start.p, 0, sizeof(int) * (size_t)std::max(n_img, 1), st)); MM_CUDA(cudaMemsetAsync(s->bp_end.p, 0, sizeof(int) * (size_t)std::max(n_img, 1), st)); }
  if (!blk_a.p) { MM_CUDA(blk_a.alloc(1)); MM_CUDA(blk_b.alloc(1)); }
  s->n_off = n_off; s->nblk = (int64_t)n_img + n_off;
  SETUP_MARK("structure: pair lists");
  return MM_OK;
}

// the image coordinates of the observations are not needed for the structure: they follow once the symbolic analysis of the
// tile Cholesky is running on its host thread (160 MB at cfg4, the largest part of the upload)
int upload_obs_xy(mm_ba_session* s, const mm_ba_problem* P) {
  cudaStream_t st = s->stream; const int64_t n = s->n_obs;
  if (n > 0) {
    DevBuf<double2> xy_in; MM_CUDA(xy_in.alloc((size_t)n));
    MM_CUDA(upload_async(xy_in.p, P->obs_xy, sizeof(double2) * n, st));
    k_gather_xy<<<grid_stride(n, 256), 256, 0, st>>>(n, s->obs_perm.p, xy_in.p, s->obs_xy.p); MM_LAUNCH_CHECK();
    MM_CUDA(cudaStreamSynchronize(st));
  }
  s->obs_perm.release();
  return MM_OK;
}

// 5. block-CSR rows for the SpMV (each off-diagonal block referenced from both rows)
int build_block_csr(mm_ba_session* s) {
  cudaStream_t st = s->stream; const int n_img = s->n_img, n_off = s->n_off; const int B = 256;
  const int64_t n_ent = (int64_t)n_img + 2 * (int64_t)n_off; s->n_ent = n_ent;
  MM_CUDA(s->row_start.alloc((size_t)n_img + 1)); MM_CUDA(s->row_col.alloc((size_t)n_ent)); MM_CUDA(s->row_blk.alloc((size_t)n_ent));
  if (n_ent > 0) {
    Arena A3; const size_t ne = (size_t)n_ent, ni1 = (size_t)n_img + 1;
    const int bits = bits_for((unsigned long long)n_img * (unsigned long long)n_img);
    const size_t cub3 = std::max(sort_bytes<unsigned long long, int>(n_ent, bits), scan_bytes<int>((int64_t)ni1));
    MM_CUDA(A3.reserve(2 * Arena::pad(8 * ne) + Arena::pad(4 * ne) + Arena::pad(4 * ni1) + Arena::pad(cub3)));
    unsigned long long* keys = A3.take<unsigned long long>(ne); unsigned long long* keys_s = A3.take<unsigned long long>(ne); int* vals = A3.take<int>(ne); int* cnt = A3.take<int>(ni1);
    void* tmp3 = A3.take<char>(cub3);
    if (!tmp3) { set_error("internal: setup arena too small"); return MM_ERR_ALLOC; }
    MM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ni1, st));
    k_csr_entries<<<blocks_for(std::max(n_img, n_off), B), B, 0, st>>>(n_img, n_off, s->blk_a.p, s->blk_b.p, keys, vals); MM_LAUNCH_CHECK();
    int rc = sort_pairs<unsigned long long, int>(st, tmp3, cub3, keys, keys_s, vals, s->row_blk.p, n_ent, bits); if (rc) return rc;
    k_csr_rows<<<grid_stride(n_ent, B), B, 0, st>>>(n_ent, n_img, keys_s, s->row_col.p, cnt); MM_LAUNCH_CHECK();
    rc = exclusive_sum<int>(st, tmp3, cub3, cnt, s->row_start.p, (int64_t)ni1); if (rc) return rc;
    MM_CUDA(cudaStreamSynchronize(st));
  } else {
    MM_CUDA(cudaMemsetAsync(s->row_start.p, 0, sizeof(int) * ((size_t)n_img + 1), st));
  }
  MM_CUDA(cudaStreamSynchronize(st));
  return MM_OK;
}

// parameters of the first iterate: poses / intrinsics from the session's host copies, points from `pts_host` (caller order)
// through the renumbering on the device
int upload_params(mm_ba_session* s, const double* pts_host) {
  cudaStream_t st = s->stream;
  MM_CUDA(cudaMemcpyAsync(s->poses.p, s->h_poses0.data(), sizeof(double) * 6 * (size_t)s->n_img, cudaMemcpyHostToDevice, st));
  MM_CUDA(cudaMemcpyAsync(s->intr.p, s->h_intr0.data(), sizeof(double) * MM_INTR_STRIDE * (size_t)s->n_cam, cudaMemcpyHostToDevice, st));
  MM_CUDA(cudaMemcpyAsync(s->intr2.p, s->h_intr0.data(), sizeof(double) * MM_INTR_STRIDE * (size_t)s->n_cam, cudaMemcpyHostToDevice, st));
  if (s->n_pt > 0) {      // pts2 (the candidate buffer) is free before the first iteration: staging in caller order
    MM_CUDA(upload_async(s->pts2.p, pts_host, sizeof(double) * 3 * (size_t)s->n_pt, st));
    k_gather_rows3<<<blocks_for(s->n_pt, 256), 256, 0, st>>>(s->n_pt, s->pt_new2old.p, s->pts2.p, s->pts.p); MM_LAUNCH_CHECK();
  }
  return MM_OK;
}

// ---- coarse level: aggregates (host, greedy over the block graph of S) ------------------------------------------
int build_aggregates(mm_ba_session* s) {
  s->cm = 0; s->n_agg = 0;
  const int n = s->n_img;
  int target = getenv("MM_PCG_AGG") ? atoi(getenv("MM_PCG_AGG")) : 8;
  if (getenv("MM_PCG_NO_COARSE") || n < 64 || target < 2 || s->n_ent <= n) return MM_OK;
  while ((int64_t)CM * ((n + target - 1) / target) > 2048) ++target;          // keeps the dense coarse inverse <= 32 MB
  std::vector<int> rs((size_t)n + 1), col((size_t)s->n_ent);
  MM_CUDA(cudaMemcpy(rs.data(), s->row_start.p, sizeof(int) * rs.size(), cudaMemcpyDeviceToHost));
  MM_CUDA(cudaMemcpy(col.data(), s->row_col.p, sizeof(int) * col.size(), cudaMemcpyDeviceToHost));
  std::vector<int> agg((size_t)n, -1);
  std::vector<std::vector<int>> mem;
  std::vector<int> front, nxt;
  for (int seed = 0; seed < n; ++seed) {
    if (agg[seed] >= 0) continue;
    const int id = (int)mem.size(); mem.emplace_back();
    std::vector<int>& g = mem.back();
    g.push_back(seed); agg[seed] = id; front.assign(1, seed);
    while ((int)g.size() < target && !front.empty()) {
      nxt.clear();
      for (int v : front)
        for (int e = rs[v]; e < rs[v + 1] && (int)g.size() < target; ++e) {
          const int u = col[e];
          if (agg[u] < 0) { agg[u] = id; g.push_back(u); nxt.push_back(u); }
        }
      front.swap(nxt);
    }
  }
  // left-over fragments join a neighbouring aggregate (a single image cannot carry 7 modes)
  const int min_size = std::max(2, target / 2);
  for (size_t a = 0; a < mem.size(); ++a) {
    if (mem[a].empty() || (int)mem[a].size() >= min_size) continue;
    int dst = -1;
    for (int v : mem[a]) { for (int e = rs[v]; e < rs[v + 1]; ++e) if (agg[col[e]] != (int)a) { dst = agg[col[e]]; break; } if (dst >= 0) break; }
    if (dst < 0) continue;
    for (int v : mem[a]) { agg[v] = dst; mem[dst].push_back(v); }
    mem[a].clear();
  }
  std::vector<int> remap(mem.size(), -1); int na = 0;
  for (size_t a = 0; a < mem.size(); ++a) if (!mem[a].empty()) remap[a] = na++;
  for (int i = 0; i < n; ++i) agg[i] = remap[agg[i]];
  s->h_agg = agg; s->n_agg = na; s->cm = CM * na;
  const size_t m = (size_t)s->cm;
  MM_CUDA(s->agg.alloc((size_t)n)); MM_CUDA(s->Pc.alloc((size_t)PCS * n)); MM_CUDA(s->Ac.alloc(m * m));
  MM_CUDA(s->gjC.alloc(m * GJ_NB)); MM_CUDA(s->gjR.alloc(m * GJ_NB)); MM_CUDA(s->crc.alloc(2 * m)); MM_CUDA(s->cqc.alloc(m)); MM_CUDA(s->cyc.alloc(m));
  MM_CUDA(cudaMemcpy(s->agg.p, agg.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
  int per_sm = 0;
  MM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_spd_inverse, 256, 0));
  const int tiles = (s->cm + 63) / 64;
  s->gj_grid = std::max(1, std::min(std::max(1, per_sm) * num_sms(), std::max(tiles * tiles, (s->cm + 255) / 256)));
  return MM_OK;
}

// ---- multi-GPU: this rank's points / observations and its part of every per-image and per-block list -----------
int build_shard(mm_ba_session* s) {
  cudaStream_t st = s->stream;
  const int n_pt = s->n_pt, n_img = s->n_img;
  s->p_lo = (int)((int64_t)n_pt * s->rank / s->world); s->p_hi = (int)((int64_t)n_pt * (s->rank + 1) / s->world);
  int o2[2] = {0, 0};
  MM_CUDA(cudaMemcpy(&o2[0], s->pt_start.p + s->p_lo, sizeof(int), cudaMemcpyDeviceToHost));
  MM_CUDA(cudaMemcpy(&o2[1], s->pt_start.p + s->p_hi, sizeof(int), cudaMemcpyDeviceToHost));
  s->o_lo = o2[0]; s->o_hi = o2[1];
  if (s->world == 1) {       // windows onto the full lists
    s->cam_lo.view(s->cam_start.p, (size_t)n_img); s->cam_hi.view(s->cam_start.p + 1, (size_t)n_img);
    s->bp_lo.view(s->bp_start.p, (size_t)s->nblk); s->bp_hi.view(s->bp_end.p, (size_t)s->nblk);
    return MM_OK;
  }
  MM_CUDA(s->cam_lo.alloc((size_t)std::max(n_img, 1))); MM_CUDA(s->cam_hi.alloc((size_t)std::max(n_img, 1)));
  MM_CUDA(s->bp_lo.alloc((size_t)std::max<int64_t>(s->nblk, 1))); MM_CUDA(s->bp_hi.alloc((size_t)std::max<int64_t>(s->nblk, 1)));
  if (n_img > 0) { k_shard_cam_ranges<<<blocks_for(n_img, 128), 128, 0, st>>>(n_img, s->cam_start.p, s->cam_perm.p, (int)s->o_lo, (int)s->o_hi, s->cam_lo.p, s->cam_hi.p); MM_LAUNCH_CHECK(); }
  if (s->nblk > 0) { k_shard_blk_ranges<<<blocks_for(s->nblk, 128), 128, 0, st>>>(s->nblk, s->bp_start.p, s->bp_end.p, s->sp_pt.p, s->p_lo, s->p_hi, s->bp_lo.p, s->bp_hi.p); MM_LAUNCH_CHECK(); }
  MM_CUDA(cudaStreamSynchronize(st));
  return MM_OK;
}

// prolongation blocks: the 7 similarity modes of each aggregate in the (scaled) pose parameters of its images.
//   translation d:  d t = -R d          rotation a (about the world origin):  d w = -Jl^-1 R a, d t = 0
//   scale about the aggregate centre c:  d t = -R (C - c),   C = -R' t the camera centre
// (rotation about c differs from rotation about the origin by a translation, so the span is the same.)
int build_coarse_basis(mm_ba_session* s) {
  if (!s->cm) return MM_OK;
  const int n = s->n_img;
  if (s->h_pose_mask.size() != 6 * (size_t)n) {
    s->h_pose_mask.resize(6 * (size_t)n);
    MM_CUDA(cudaMemcpyAsync(s->h_pose_mask.data(), s->pose_mask.p, sizeof(double) * 6 * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
  }
  std::vector<double> sc(6 * (size_t)n), P((size_t)PCS * n, 0.0), R(9 * (size_t)n), M(9 * (size_t)n), C(3 * (size_t)n), cen(3 * (size_t)s->n_agg, 0.0);
  std::vector<int> cnt((size_t)s->n_agg, 0);
  MM_CUDA(cudaMemcpyAsync(sc.data(), s->scale_c.p, sizeof(double) * sc.size(), cudaMemcpyDeviceToHost, s->stream));
  MM_CUDA(cudaStreamSynchronize(s->stream));
  for (int i = 0; i < n; ++i) {
    const double* p = s->h_poses0.data() + 6 * (size_t)i;
    double Jl[9]; double* Ri = R.data() + 9 * (size_t)i;
    rotation_and_left_jacobian(p, Ri, Jl);
    // M = Jl^-1 R  (adjugate inverse of the 3 x 3 left Jacobian)
    const double a = Jl[0], b = Jl[1], c = Jl[2], d = Jl[3], e = Jl[4], f = Jl[5], g = Jl[6], h = Jl[7], k = Jl[8];
    const double det = a * (e * k - f * h) - b * (d * k - f * g) + c * (d * h - e * g);
    const double id = (fabs(det) > 1e-300) ? 1.0 / det : 0.0;
    const double Ji[9] = { (e * k - f * h) * id, (c * h - b * k) * id, (b * f - c * e) * id,
                           (f * g - d * k) * id, (a * k - c * g) * id, (c * d - a * f) * id,
                           (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id };
    double* Mi = M.data() + 9 * (size_t)i;
    for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) Mi[3 * r + q] = Ji[3 * r] * Ri[q] + Ji[3 * r + 1] * Ri[3 + q] + Ji[3 * r + 2] * Ri[6 + q];
    for (int q = 0; q < 3; ++q) C[3 * (size_t)i + q] = -(Ri[q] * p[3] + Ri[3 + q] * p[4] + Ri[6 + q] * p[5]);
    const int gidx = s->h_agg[i];
    for (int q = 0; q < 3; ++q) cen[3 * (size_t)gidx + q] += C[3 * (size_t)i + q];
    cnt[gidx]++;
  }
  for (int gi_ = 0; gi_ < s->n_agg; ++gi_) for (int q = 0; q < 3; ++q) cen[3 * (size_t)gi_ + q] /= std::max(cnt[gi_], 1);
  for (int i = 0; i < n; ++i) {
    const double* Ri = R.data() + 9 * (size_t)i; const double* Mi = M.data() + 9 * (size_t)i;
    const double* c0 = cen.data() + 3 * (size_t)s->h_agg[i];
    double* Pi = P.data() + (size_t)PCS * i;
    double dc[3]; for (int q = 0; q < 3; ++q) dc[q] = C[3 * (size_t)i + q] - c0[q];
    for (int r = 0; r < 3; ++r) {
      for (int k = 0; k < 3; ++k) { Pi[CM * (3 + r) + k] = -Ri[3 * r + k]; Pi[CM * r + 3 + k] = -Mi[3 * r + k]; }
      Pi[CM * (3 + r) + 6] = -(Ri[3 * r] * dc[0] + Ri[3 * r + 1] * dc[1] + Ri[3 * r + 2] * dc[2]);
    }
    for (int r = 0; r < 6; ++r) {
      const double w = s->h_pose_mask[6 * (size_t)i + r] / sc[6 * (size_t)i + r];       // scaled coordinates, free parameters only
      for (int k = 0; k < CM; ++k) Pi[CM * r + k] *= w;
    }
  }
  MM_CUDA(cudaMemcpy(s->Pc.p, P.data(), sizeof(double) * P.size(), cudaMemcpyHostToDevice));
  return MM_OK;
}

// Ac = P' S P and its explicit inverse (in place)
int launch_coarse_setup(mm_ba_session* s) {
  if (!s->cm) return MM_OK;
  cudaStream_t st = s->stream;
  int m = s->cm;
  MM_CUDA(cudaMemsetAsync(s->Ac.p, 0, sizeof(double) * (size_t)m * m, st));
  k_coarse_assemble<<<blocks_for(s->nblk * 32, 128), 128, 0, st>>>(s->n_img, s->nblk, s->blk_a.p, s->blk_b.p, s->S.p, s->agg.p, s->Pc.p, m, s->Ac.p); MM_LAUNCH_CHECK();
  k_coarse_ridge<<<blocks_for(m, 128), 128, 0, st>>>(m, s->Ac.p); MM_LAUNCH_CHECK();
  double* Ap = s->Ac.p; double* Ck = s->gjC.p; double* Rk = s->gjR.p; unsigned long long* dbg = nullptr;
  void* args[] = { &m, &Ap, &Ck, &Rk, &dbg };
  MM_CUDA(cudaLaunchCooperativeKernel((const void*)k_spd_inverse, dim3(s->gj_grid), dim3(256), args, 0, st));
  count_launch();
  return MM_OK;
}

int all_reduce(mm_ba_session* s, double* buf, size_t count) {
  if (s->world <= 1) return MM_OK;
  const int rc = s->ar(s->ar_user, buf, (int64_t)count, (void*)s->stream);
  if (rc) { set_error("the all-reduce callback failed (%d)", rc); return MM_ERR_CUDA; }
  return MM_OK;
}

LossParams loss_of(const mm_ba_options& o) {
  LossParams L; L.type = o.loss_type; L.b = o.loss_scale * o.loss_scale; L.c = 1.0 / L.b; return L;
}
LMDiag lm_of(const mm_ba_session* s) { LMDiag d; d.radius = s->radius; d.min_diag = s->opt.min_lm_diagonal; d.max_diag = s->opt.max_lm_diagonal; return d; }

// K1 at the current iterate: records + cost -> red[0].  reuse_candidate_cost: the iterate is the candidate that the cost-only
// pass has just evaluated (accepted step), so its cost (and that of the rotation constraints) is taken over instead of recomputed.
template <bool WITH_COST>
int launch_k1(mm_ba_session* s) {
  cudaStream_t st = s->stream;
  if (s->refine)
    k_residual_jacobian<true, true, WITH_COST><<<s->grid_obs, 256, K1_SMEM, st>>>(s->no_loc(), s->obs_xy.p + s->o_lo, s->obs_img.p + s->o_lo, s->obs_pt.p + s->o_lo, s->aux.p, s->pts.p, s->intr.p,
        s->img_cam.p, s->cam_model.p, s->pose_mask.p, s->pt_mask.p, loss_of(s->opt), s->rec.p, s->part_cost.p, s->ji.p, s->intr_mask.p);
  else
    k_residual_jacobian<true, false, WITH_COST><<<s->grid_obs, 256, K1_SMEM, st>>>(s->no_loc(), s->obs_xy.p + s->o_lo, s->obs_img.p + s->o_lo, s->obs_pt.p + s->o_lo, s->aux.p, s->pts.p, s->intr.p,
        s->img_cam.p, s->cam_model.p, s->pose_mask.p, s->pt_mask.p, loss_of(s->opt), s->rec.p, s->part_cost.p);
  MM_LAUNCH_CHECK();
  return MM_OK;
}
int launch_linearize(mm_ba_session* s, bool reuse_candidate_cost = false) {
  cudaStream_t st = s->stream;
  k_pose_aux<<<blocks_for(s->n_img, 128), 128, 0, st>>>(s->n_img, s->poses.p, s->pose_mask.p, s->aux.p); MM_LAUNCH_CHECK();
  if (reuse_candidate_cost) {
    int rc = launch_k1<false>(s); if (rc) return rc;
    MM_CUDA(cudaMemcpyAsync(s->loc.p + 0, s->loc.p + 1, sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (s->n_prior) {      // the constraints' Jacobian at the new iterate; their cost was evaluated with the candidate
      k_rot_prior<<<1, 256, 0, st>>>(s->n_img, s->poses.p, s->pose_mask.p, s->pr_rot0.p, s->pr_w.p, s->pr_r.p, s->pr_J.p, s->loc.p + 6); MM_LAUNCH_CHECK();
    }
    return MM_OK;
  }
  int rc = launch_k1<true>(s); if (rc) return rc;
  k_reduce_sum<<<1, 256, 0, st>>>(s->part_cost.p, s->grid_obs, s->loc.p + 0); MM_LAUNCH_CHECK();
  if (s->n_prior) { k_rot_prior<<<1, 256, 0, st>>>(s->n_img, s->poses.p, s->pose_mask.p, s->pr_rot0.p, s->pr_w.p, s->pr_r.p, s->pr_J.p, s->loc.p + 6); MM_LAUNCH_CHECK(); }
  return MM_OK;
}
// K4: cost at the candidate (poses2/pts2) -> red[1]
int launch_cost_candidate(mm_ba_session* s) {
  cudaStream_t st = s->stream;
  k_pose_aux<<<blocks_for(s->n_img, 128), 128, 0, st>>>(s->n_img, s->poses2.p, s->pose_mask.p, s->aux2.p); MM_LAUNCH_CHECK();
  k_residual_jacobian<false><<<s->grid_obs, 256, K1_SMEM_COST, st>>>(s->no_loc(), s->obs_xy.p + s->o_lo, s->obs_img.p + s->o_lo, s->obs_pt.p + s->o_lo, s->aux2.p, s->pts2.p, s->refine ? s->intr2.p : s->intr.p,
      s->img_cam.p, s->cam_model.p, s->pose_mask.p, s->pt_mask.p, loss_of(s->opt), nullptr, s->part_cost.p); MM_LAUNCH_CHECK();
  k_reduce_sum<<<1, 256, 0, st>>>(s->part_cost.p, s->grid_obs, s->loc.p + 1); MM_LAUNCH_CHECK();
  // the step scalars (candidate cost, |step|^2, model cost change) summed over the ranks -> red[1], red[3], red[4]
  if (s->n_prior) { k_rot_prior<<<1, 256, 0, st>>>(s->n_img, s->poses2.p, s->pose_mask.p, s->pr_rot0.p, s->pr_w.p, nullptr, nullptr, s->loc.p + 7); MM_LAUNCH_CHECK(); }
  k_pack_step<<<1, 32, 0, st>>>(s->loc.p, s->fail.p, s->stepbuf.p); MM_LAUNCH_CHECK();
  { const int rc = all_reduce(s, s->stepbuf.p, 4); if (rc) return rc; }
  k_unpack_step<<<1, 32, 0, st>>>(s->stepbuf.p, s->red.p, s->fail.p, s->n_prior ? s->loc.p + 7 : nullptr); MM_LAUNCH_CHECK();
  return MM_OK;
}
int launch_scale(mm_ba_session* s) {
  cudaStream_t st = s->stream;
  if (s->opt.jacobi_scaling) {
    k_colnorm_point<<<blocks_for(s->np_loc(), 128), 128, 0, st>>>(s->np_loc(), s->pt_start.p + s->p_lo, s->rec_base(), s->scale_p.p + 3 * (size_t)s->p_lo); MM_LAUNCH_CHECK();
    k_colnorm_cam<<<blocks_for((int64_t)s->n_img * 32, 128), 128, 0, st>>>(s->n_img, s->cam_lo.p, s->cam_hi.p, s->cam_perm.p, s->rec_base(), s->ud.p); MM_LAUNCH_CHECK();
    { const int rc = all_reduce(s, s->ud.p, 6 * (size_t)s->n_img); if (rc) return rc; }
    k_scale_finish<<<blocks_for(6 * (int64_t)s->n_img, 256), 256, 0, st>>>(6 * s->n_img, s->ud.p, s->scale_c.p, s->n_prior ? s->pr_J.p : nullptr); MM_LAUNCH_CHECK();
    if (s->refine) {
      k_colnorm_intr_img<<<blocks_for((int64_t)s->n_img * 32, 128), 128, 0, st>>>(s->n_img, s->cam_lo.p, s->cam_hi.p, s->cam_perm.p, s->ji_base(), s->img_sq.p); MM_LAUNCH_CHECK();
      { const int rc = all_reduce(s, s->img_sq.p, 9 * (size_t)s->n_img); if (rc) return rc; }
      k_intr_scale<<<s->n_cam, 256, 0, st>>>(s->n_img, s->img_cam.p, s->img_sq.p, s->scale_i.p); MM_LAUNCH_CHECK();
    }
  }
  return MM_OK;
}
// K2: reduced camera system at the current radius (+ gradient max-norm -> red[2])
int launch_schur(mm_ba_session* s, bool with_coarse = true) {
  cudaStream_t st = s->stream;
  MM_CUDA(cudaMemsetAsync(s->red.p + 2, 0, sizeof(double), st));
  MM_CUDA(cudaMemsetAsync(s->scal(), 0, sizeof(double) * (3 + (size_t)s->world), st));
  const LMDiag lm = lm_of(s);
  const int P0 = s->p_lo, NP = s->np_loc();
  k_schur_point<<<blocks_for(NP, 128), 128, 0, st>>>(NP, s->pt_start.p + P0, s->rec_base(), s->scale_p.p + 3 * (size_t)P0, lm,
      s->Vinv.p + 6 * (size_t)P0, s->gp.p + 3 * (size_t)P0, s->dp.p + 3 * (size_t)P0, s->scal() + 3 + s->rank, s->fail.p, s->pinfo.p + PINFO * (size_t)P0); MM_LAUNCH_CHECK();
  k_schur_blocks<<<blocks_for(s->nblk * 32, 128), 128, 0, st>>>(s->n_img, s->nblk, s->blk_a.p, s->blk_b.p, s->bp_lo.p, s->bp_hi.p, s->sp_lo.p, s->sp_hi.p,
      s->sp_pt.p, s->rec_base(), s->scale_c.p, s->pinfo.p, s->S.p); MM_LAUNCH_CHECK();
  if (s->n_img > 0) { k_schur_cam<<<s->n_img, 128, 0, st>>>(s->n_img, s->cam_lo.p, s->cam_hi.p, s->cam_perm.p, s->obs_pt.p, s->rec_base(),
      s->scale_c.p, s->pinfo.p, s->S.p, s->rhs.p, s->gc.p, s->ud.p); MM_LAUNCH_CHECK(); }
  // exchange step: S | rhs | gc | ud | scalars summed over the ranks (nothing to do on one GPU), then the LM diagonal
  k_pack_scal<<<1, 32, 0, st>>>(s->loc.p, s->fail.p, s->scal()); MM_LAUNCH_CHECK();
  { const int rc = all_reduce(s, s->xch.p, s->xch_count); if (rc) return rc; }
  k_cam_finish<<<blocks_for(6 * (int64_t)s->n_img, 128), 128, 0, st>>>(6 * s->n_img, s->ud.p, s->gc.p, s->rhs.p, s->scale_c.p, lm, s->S.p, s->dc.p, s->scal(), s->world,
      s->red.p, s->fail.p, s->n_prior ? s->pr_r.p : nullptr, s->n_prior ? s->pr_J.p : nullptr, s->n_prior ? s->loc.p + 6 : nullptr); MM_LAUNCH_CHECK();
  if (!s->tc_on) { k_precond<<<blocks_for(s->n_img, 64), 64, 0, st>>>(s->n_img, s->S.p, s->Minv.p, s->fail.p); MM_LAUNCH_CHECK(); }   // (the tile factorisation checks definiteness itself)
  if (with_coarse && s->cm) {
    // The coarse inverse is only a preconditioner: once the iteration has settled (S changes little from one LM step to the
    // next) it is refreshed every third step, or when the trust-region radius - the damping inside S - moved by more than 30x.
    const double rr = s->coarse_radius > 0.0 ? s->radius / s->coarse_radius : 0.0;
    const bool refresh = s->iter <= 3 || s->coarse_iter < 0 || s->iter - s->coarse_iter >= 3 || rr > 30.0 || rr < 1.0 / 30.0 || getenv("MM_PCG_COARSE_EVERY");
    if (refresh) { const int rc = launch_coarse_setup(s); if (rc) return rc; s->coarse_iter = s->iter; s->coarse_radius = s->radius; }
  }
  if (s->refine) {
    // border of the reduced system: B (6 x 9 per image and camera) | C, b, diagonal, gradient of the intrinsics; summed over the
    // ranks like S (one more exchange step), then the LM diagonal of the intrinsics
    const int P0 = s->p_lo, NP = s->np_loc(); const int n9 = 9 * s->n_cam;
    MM_CUDA(cudaMemsetAsync(s->intr_acc.p, 0, sizeof(double) * ((size_t)n9 * n9 + 3 * (size_t)n9), st));
    k_schur_intr_point<<<blocks_for(NP, 128), 128, 0, st>>>(NP, s->n_cam, s->pt_start.p + P0, s->obs_img.p, s->img_cam.p, s->rec_base(), s->ji_base(),
        s->scale_p.p + 3 * (size_t)P0, s->scale_i.p, s->Vinv.p + 6 * (size_t)P0, s->gp.p + 3 * (size_t)P0, s->Apc.p + 27 * (size_t)P0 * s->n_cam, s->pt_cam_mask.p + P0, s->intr_acc.p); MM_LAUNCH_CHECK();
    k_schur_cam_intr<<<blocks_for((int64_t)s->n_img * 32, 128), 128, 0, st>>>(s->n_img, s->n_cam, s->cam_lo.p, s->cam_hi.p, s->cam_perm.p, s->obs_pt.p, s->img_cam.p, s->rec_base(), s->ji_base(),
        s->scale_c.p, s->scale_p.p, s->scale_i.p, s->Vinv.p, s->Apc.p, s->pt_cam_mask.p, s->Bm.p); MM_LAUNCH_CHECK();
    { const int rc = all_reduce(s, s->xch2.p, s->xch2_count); if (rc) return rc; }
    k_intr_finalize<<<blocks_for(n9, 64), 64, 0, st>>>(n9, lm, s->scale_i.p, s->intr_acc.p, s->gi.p, s->di.p, s->red.p + 2); MM_LAUNCH_CHECK();
  }
  return MM_OK;
}

// ---- sparse tile Cholesky (tilechol_plan.h / tilechol.cuh): plan at session creation, factorisation per Schur assembly ------
template <typename T, typename U>
int tc_upload(DevBuf<T>& d, const std::vector<U>& h) {
  static_assert(sizeof(T) == sizeof(U), "element size");
  MM_CUDA(d.alloc(std::max<size_t>(h.size(), 1)));
  if (!h.empty()) MM_CUDA(cudaMemcpy(d.p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return MM_OK;
}
// host part (no CUDA calls: runs on a worker thread): symbolic analysis + the packed copies of the plan
void tc_plan_host(mm_ba_session* s) {
  const int n = s->n_img;
  TileCholPlan& P = s->tc_plan;
  s->tc_plan_rc = build_tilechol_plan(n, s->n_off, s->tc_ha.data(), s->tc_hb.data(), getenv("MM_TC_NO_GEOMETRY") ? nullptr : s->tc_pos.data(), s->ncb, P);
  if (s->tc_plan_rc != 0) return;
  const size_t n_tasks = (size_t)(P.n_l + P.n_wtask);
  if (P.n_upd + P.n_wupd >= ((int64_t)1 << 31) || (int64_t)P.it_mat.size() >= ((int64_t)1 << 31)) { s->tc_plan_rc = -100; return; }
  std::vector<int>& sched = s->tc_h_sched; std::vector<int>& sdesc = s->tc_h_sdesc; std::vector<int4>& upd = s->tc_h_upd; std::vector<int2>& items = s->tc_h_items;
  sched.assign(8 * n_tasks, 0); sdesc.assign(16 * (size_t)P.n_stasks, 0); upd.resize((size_t)(P.n_upd + P.n_wupd)); items.resize(P.it_mat.size());
  for (int64_t u = 0; u < P.n_upd; ++u) upd[u] = make_int4(P.upd_a[u], P.upd_b[u], P.upd_a[u], P.upd_b[u]);
  // flags: [0, n_tasks) one per factor task (a diagonal task sets its flag when L(j,j) is stored), then one per tile row, set when
  // the inverse W(j,j) is stored as well
  for (int64_t u = 0; u < P.n_wupd; ++u) {
    int fl = P.wupd_flag[u];
    if (fl < P.n_l) fl = (int)n_tasks + P.col_idx[fl];            // operand W(k,k): produced by the diagonal task of column k, second flag
    upd[P.n_upd + u] = make_int4(P.wupd_l[u], P.wupd_w[u], P.wupd_l[u], fl | (1 << 30));
  }
  for (size_t slot = 0; slot < n_tasks; ++slot) {
    const int t = P.task_order[slot]; int* d = sched.data() + 8 * slot;
    d[0] = t;
    if (t < P.n_l) {
      const int j = P.col_idx[t]; const bool diag = P.row_idx[t] == j;
      d[1] = diag ? 1 : 0; d[2] = (int)P.upd_ptr[t]; d[3] = (int)(P.upd_ptr[t + 1] - P.upd_ptr[t]);
      d[4] = diag ? (int)(P.w_row_ptr[j + 1] - 1) : (int)P.col_ptr[j]; d[5] = (int)P.col_ptr[j]; d[6] = P.has_a[t] | (j << 1); d[7] = diag ? P.tile_nunk[j] : 0;
    } else {
      const int w = t - (int)P.n_l; const int i = P.wt_row[w];
      d[1] = 2; d[2] = (int)(P.n_upd + P.wupd_ptr[w]); d[3] = (int)(P.wupd_ptr[w + 1] - P.wupd_ptr[w]);
      d[4] = (int)(P.w_row_ptr[i + 1] - 1); d[5] = (int)n_tasks + i; d[6] = i << 1; d[7] = P.wt_store[w];
    }
  }
  for (int k = 0; k < P.n_stasks; ++k) {
    int* d = sdesc.data() + 16 * (size_t)k;
    const int64_t b = P.st_item_ptr[k], e = P.st_item_ptr[k + 1];
    d[0] = P.st_kind[k]; d[1] = P.st_out[k]; d[2] = P.st_base[k]; d[3] = P.st_tile[k]; d[4] = (int)b; d[5] = (int)(e - b);
    for (int q = 0; q < 4 && b + q < e; ++q) { d[8 + 2 * q] = P.it_mat[b + q]; d[9 + 2 * q] = P.it_src[b + q]; }
  }
  for (size_t q = 0; q < items.size(); ++q) items[q] = make_int2(P.it_mat[q], P.it_src[q]);
}
// decides whether the session solves through the tile factorisation and starts the symbolic analysis on a host thread
int tc_begin(mm_ba_session* s) {
  s->tc_on = false; s->tc_plan_rc = 1;           // 1 = not attempted
  const int pre = s->opt.pcg_preconditioner;
  const char* env = getenv("MM_PCG_PRECOND");               // "twolevel" | "tilechol" overrides the option (experiments)
  const bool force_two = (env && !strcmp(env, "twolevel")) || (!env && pre == MM_PRECOND_TWO_LEVEL);
  const bool force_tc = (env && !strcmp(env, "tilechol")) || (!env && pre == MM_PRECOND_TILE_CHOLESKY);
  s->tc_force = force_tc;
  if (s->n_img <= 0 || s->n_obs == 0) return MM_OK;
  if (force_two && !s->refine) return MM_OK;
  if (!s->refine && !force_tc && s->n_img <= DENSE_MAX_IMG) return MM_OK;          // dense Cholesky in one CTA
  const int n = s->n_img;
  s->tc_ha.assign((size_t)std::max(s->n_off, 1), 0); s->tc_hb.assign((size_t)std::max(s->n_off, 1), 0);
  if (s->n_off > 0) {
    MM_CUDA(cudaMemcpyAsync(s->tc_ha.data(), s->blk_a.p, sizeof(int) * (size_t)s->n_off, cudaMemcpyDeviceToHost, s->stream));
    MM_CUDA(cudaMemcpyAsync(s->tc_hb.data(), s->blk_b.p, sizeof(int) * (size_t)s->n_off, cudaMemcpyDeviceToHost, s->stream));
    MM_CUDA(cudaStreamSynchronize(s->stream));
  }
  // camera centres C = -R' t guide the dissection
  s->tc_pos.resize(3 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    const double* p = s->h_poses0.data() + 6 * (size_t)i; double R[9], Jl[9];
    rotation_and_left_jacobian(p, R, Jl);
    for (int q = 0; q < 3; ++q) s->tc_pos[3 * (size_t)i + q] = -(R[q] * p[3] + R[3 + q] * p[4] + R[6 + q] * p[5]);
  }
  s->ncb = s->refine ? s->n_cam : 0;
  s->tc_plan_rc = 2;                              // 2 = running
  bool started = false;
  if (!getenv("MM_TC_PLAN_INLINE")) { try { s->tc_thread = std::thread(tc_plan_host, s); started = true; } catch (...) {} }
  if (!started) tc_plan_host(s);
  return MM_OK;
}
// joins the analysis and puts the plan on the device.  *fallback = true: the session goes on with the two-level preconditioner
int tc_finish(mm_ba_session* s, bool* fallback) {
  *fallback = false;
  if (s->tc_thread.joinable()) s->tc_thread.join();
  if (s->tc_plan_rc == 1) { *fallback = true; return MM_OK; }
  s->tc_ha = std::vector<int>(); s->tc_hb = std::vector<int>(); s->tc_pos = std::vector<double>();
  TileCholPlan& P = s->tc_plan;
  if (s->tc_plan_rc != 0) { set_error("tile Cholesky plan failed (%d)", s->tc_plan_rc); return MM_ERR_UNSUPPORTED; }
  SETUP_MARK("tilechol: wait for the host plan");
  // memory: the tiles of L must fit comfortably beside the Jacobian records
  size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
  const size_t need = sizeof(double) * TC_TT * (2 * (size_t)P.n_l + 2 * (size_t)P.n_w);      // L, its operand copy, W in both orientations
  if (need > free_b / 2 && !s->tc_force && !s->refine) { P = TileCholPlan(); *fallback = true; return MM_OK; }          // falls back to the two-level PCG
  const int n = s->n_img;
  const size_t n_tasks = (size_t)(P.n_l + P.n_wtask);
  int rc;
  if ((rc = tc_upload(s->tc_unk_of, P.unk_of)) || (rc = tc_upload(s->tc_sc_tile, P.sc_tile)) || (rc = tc_upload(s->tc_sc_off, P.sc_off)) || (rc = tc_upload(s->tc_a_tiles, P.a_tiles)) ||
      (rc = tc_upload(s->tc_img_tile, P.img_tile)) || (rc = tc_upload(s->tc_img_slot, P.img_slot)) || (rc = tc_upload(s->tc_col_ptr, P.col_ptr)) ||
      (rc = tc_upload(s->tc_sched, s->tc_h_sched)) || (rc = tc_upload(s->tc_sdesc, s->tc_h_sdesc)) || (rc = tc_upload(s->tc_upd, s->tc_h_upd)) || (rc = tc_upload(s->tc_items, s->tc_h_items))) return rc;
  s->tc_h_sched = std::vector<int>(); s->tc_h_sdesc = std::vector<int>(); s->tc_h_upd = std::vector<int4>(); s->tc_h_items = std::vector<int2>();
  SETUP_MARK("tilechol: plan upload");
  MM_CUDA(s->tc_Ls.alloc((size_t)TC_TT * (size_t)P.n_l));
  MM_CUDA(s->tc_L.alloc((size_t)TC_TT * (size_t)P.n_l)); MM_CUDA(s->tc_WC.alloc((size_t)TC_TT * (size_t)P.n_w)); MM_CUDA(s->tc_WR.alloc((size_t)TC_TT * (size_t)P.n_w));
  MM_CUDA(s->tc_slots.alloc((size_t)TC_T * (size_t)P.n_slots));
  MM_CUDA(s->tc_ready.alloc(n_tasks + (size_t)P.nt)); MM_CUDA(s->tc_sflag.alloc((size_t)P.n_slots)); MM_CUDA(s->tc_counters.alloc(4)); MM_CUDA(s->tc_invd.alloc((size_t)TC_T * P.nt));
  MM_CUDA(cudaMemset(s->tc_ready.p, 0, sizeof(int) * (n_tasks + (size_t)P.nt))); MM_CUDA(cudaMemset(s->tc_sflag.p, 0, sizeof(int) * (size_t)P.n_slots));
  SETUP_MARK("tilechol: tile allocation");
  s->n_unk = 6 * n + 9 * s->ncb;
  MM_CUDA(s->dv_b.alloc((size_t)s->n_unk)); MM_CUDA(s->dv_r.alloc((size_t)s->n_unk)); MM_CUDA(s->dv_z.alloc((size_t)s->n_unk)); MM_CUDA(s->dv_p.alloc((size_t)s->n_unk)); MM_CUDA(s->dv_Ap.alloc((size_t)s->n_unk));
  TcDev& D = s->tc;
  D.nt = P.nt; D.nt_pose = P.nt_pose; D.nt_border = P.nt - P.nt_pose; D.n_img = n; D.n_cam_border = s->ncb; D.n_l = (int)P.n_l; D.n_a_tiles = (int)P.a_tiles.size();
  D.n_tasks = (int)n_tasks; D.n_stasks = P.n_stasks; D.n_slots = P.n_slots;
  D.unk_of = s->tc_unk_of.p; D.sc_tile = s->tc_sc_tile.p; D.sc_off = s->tc_sc_off.p; D.a_tiles = s->tc_a_tiles.p; D.img_tile = s->tc_img_tile.p; D.img_slot = s->tc_img_slot.p;
  D.col_ptr = s->tc_col_ptr.p; D.sched = s->tc_sched.p; D.upd = s->tc_upd.p; D.sdesc = s->tc_sdesc.p; D.items = s->tc_items.p;
  D.Ls = s->tc_Ls.p;
  D.L = s->tc_L.p; D.WC = s->tc_WC.p; D.WR = s->tc_WR.p; D.slots = s->tc_slots.p; D.invd = s->tc_invd.p;
  D.ready = s->tc_ready.p; D.sflag = s->tc_sflag.p; D.counters = s->tc_counters.p; D.trace = nullptr; D.trace_diag = nullptr;
  MM_CUDA(cudaFuncSetAttribute((const void*)k_tc_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_FACTOR_SMEM));
  int per_sm = 0;
  MM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tc_factor, TC_FACTOR_THREADS, TC_FACTOR_SMEM));
  s->tc_grid_f = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)std::max(1, per_sm) * num_sms(), (int64_t)n_tasks));
  int per_sm_s = 0;
  MM_CUDA(cudaFuncSetAttribute((const void*)k_tc_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_APPLY_SMEM));
  MM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_s, k_tc_apply, TC_APPLY_THREADS, TC_APPLY_SMEM));
  // few CTAs per SM: a CTA that has taken a task polls the flags of its operands, and thousands of pollers saturate the L2
  const int apply_per_sm = getenv("MM_TC_APPLY_CTAS") ? std::max(1, atoi(getenv("MM_TC_APPLY_CTAS"))) : 4;
  s->tc_grid_s = std::max(1, std::min(std::min(std::max(1, per_sm_s), apply_per_sm) * num_sms(), P.n_stasks));
  s->tc_on = true; s->tc_epoch = 0; s->tc_sepoch = 0;
  SETUP_MARK("tilechol: upload + tiles");
  return MM_OK;
}
// assembly of S (and the intrinsics border) into the tiles + numeric factorisation
int launch_tc_factor(mm_ba_session* s) {
  cudaStream_t st = s->stream; TcDev D = s->tc;
  MM_CUDA(cudaMemsetAsync(D.counters, 0, sizeof(int) * 4, st));
  k_tc_zero_tiles<<<std::min(D.n_a_tiles, 8 * num_sms()), 256, 0, st>>>(D.n_a_tiles, D.a_tiles, D.L); MM_LAUNCH_CHECK();
  k_tc_scatter_blocks<<<grid_stride(36 * s->nblk, 256), 256, 0, st>>>(s->nblk, D.sc_tile, D.sc_off, s->S.p, D.L); MM_LAUNCH_CHECK();
  if (s->ncb) { k_tc_scatter_border<<<grid_stride((int64_t)s->n_img * s->ncb * 54 + 81 * s->ncb * s->ncb, 256), 256, 0, st>>>(D, s->Bm.p, s->intr_acc.p); MM_LAUNCH_CHECK(); }
  int epoch = ++s->tc_epoch; int* fail = s->fail.p;
  void* args[] = { &D, &epoch, &fail };
  MM_CUDA(cudaLaunchCooperativeKernel((const void*)k_tc_factor, dim3(s->tc_grid_f), dim3(TC_FACTOR_THREADS), args, TC_FACTOR_SMEM, st));
  count_launch();
  return MM_OK;
}
// z = M^-1 r with the factorisation (both substitutions); skipped on the device once the solve has converged
int launch_tc_apply(mm_ba_session* s, const double* r, double* z, const int* done) {
  cudaStream_t st = s->stream; TcDev D = s->tc;
  MM_CUDA(cudaMemsetAsync(D.counters + 1, 0, sizeof(int), st));
  int epoch = ++s->tc_sepoch;
  void* args[] = { &D, &epoch, &r, &z, &done };
  MM_CUDA(cudaLaunchCooperativeKernel((const void*)k_tc_apply, dim3(s->tc_grid_s), dim3(TC_APPLY_THREADS), args, TC_APPLY_SMEM, st)); count_launch();
  return MM_OK;
}
// PCG on the reduced system [poses | intrinsics] preconditioned by the exact tile factorisation
int launch_dpcg(mm_ba_session* s) {
  cudaStream_t st = s->stream;
  int rc = launch_tc_factor(s); if (rc) return rc;
  const size_t n6 = 6 * (size_t)s->n_img;
  MM_CUDA(cudaMemcpyAsync(s->dv_b.p, s->rhs.p, sizeof(double) * n6, cudaMemcpyDeviceToDevice, st));
  if (s->ncb) MM_CUDA(cudaMemcpyAsync(s->dv_b.p + n6, s->intr_acc.p + 81 * (size_t)s->ncb * s->ncb, sizeof(double) * 9 * (size_t)s->ncb, cudaMemcpyDeviceToDevice, st));
  DpcgVec V; V.n = s->n_unk; V.x = s->vx.p; V.r = s->dv_r.p; V.z = s->dv_z.p; V.p = s->dv_p.p; V.Ap = s->dv_Ap.p; V.b = s->dv_b.p; V.sc = s->pcg_sc.p; V.ic = s->pcg_ic.p;
  const double tol = s->opt.tile_cholesky_tolerance > 0.0 ? s->opt.tile_cholesky_tolerance : s->opt.pcg_tolerance;      // (0: refine down to pcg_tolerance)
  V.tol2 = tol * tol; V.max_iter = s->opt.pcg_max_iterations;
  k_dpcg_init<<<1, 1024, 0, st>>>(V); MM_LAUNCH_CHECK();
  int done_it = 0;
  while (done_it < V.max_iter) {
    for (int k = 0; k < 2; ++k, ++done_it) {
      if ((rc = launch_tc_apply(s, V.r, V.z, V.ic))) return rc;
      k_dpcg_direction<<<1, 1024, 0, st>>>(V); MM_LAUNCH_CHECK();
      k_dpcg_spmv<<<blocks_for((int64_t)s->n_img * 32, 128), 128, 0, st>>>(s->n_img, s->row_start.p, s->row_col.p, s->row_blk.p, s->S.p, V.p, V.Ap, s->ncb, s->ncb ? s->Bm.p : nullptr, V.ic); MM_LAUNCH_CHECK();
      if (s->ncb) { k_dpcg_spmv_border<<<s->ncb, 256, 0, st>>>(s->n_img, s->ncb, s->Bm.p, s->intr_acc.p, V.p, V.Ap, V.ic); MM_LAUNCH_CHECK(); }
      k_dpcg_update<<<1, 1024, 0, st>>>(V); MM_LAUNCH_CHECK();
    }
    int hic[2] = {0, 0};
    MM_CUDA(cudaMemcpyAsync(hic, s->pcg_ic.p, sizeof hic, cudaMemcpyDeviceToHost, st));
    MM_CUDA(cudaStreamSynchronize(st));
    if (getenv("MM_PCG_DEBUG")) { double sc[8]; cudaMemcpy(sc, s->pcg_sc.p, sizeof sc, cudaMemcpyDeviceToHost); printf("dpcg: %d iterations, |r|/|b| = %.3e (first iteration %.3e)\n", hic[1], sqrt(sc[4] / sc[0]), sqrt(sc[5] / sc[0])); }
    if (hic[0]) break;
  }
  return MM_OK;
}

// every rank solved the same replicated system, but the dot products of the solve are accumulated with atomics whose
// order differs from GPU to GPU: rank 0's solution is the one everybody continues with
int pcg_broadcast(mm_ba_session* s) {
  if (s->world <= 1) return MM_OK;
  if (s->rank != 0) MM_CUDA(cudaMemsetAsync(s->vx.p, 0, sizeof(double) * 6 * (size_t)s->n_img, s->stream));
  return all_reduce(s, s->vx.p, 6 * (size_t)s->n_img);
}
// K3: solve S y = rhs with the persistent cooperative kernel (no host sync; iteration count -> pcg_ic[1])
int launch_pcg(mm_ba_session* s) {
  cudaStream_t st = s->stream;
  MM_CUDA(cudaMemsetAsync(s->pcg_sc.p, 0, sizeof(double) * 16, st));
  MM_CUDA(cudaMemsetAsync(s->pcg_ic.p, 0, sizeof(int) * 4, st));
  if (s->tc_on) return launch_dpcg(s);          // deterministic: every rank of a sharded session computes the same bits
  if (!s->refine && s->n_img <= DENSE_MAX_IMG && s->n_img > 0 && !getenv("MM_PCG_ONLY")) {
    // local-BA sized system: dense Cholesky in one CTA (ba_coarse.cuh); reported as 0 linear iterations
    const int n = 6 * s->n_img;
    const size_t smem = sizeof(double) * ((size_t)n * (n + 1) + n);
    // (the attribute belongs to the current device / context: set on every call, it is cheap)
    MM_CUDA(cudaFuncSetAttribute(k_dense_chol_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * ((size_t)6 * DENSE_MAX_IMG * (6 * DENSE_MAX_IMG + 1) + 6 * DENSE_MAX_IMG))));
    k_dense_chol_solve<<<1, 512, smem, st>>>(s->n_img, s->nblk, s->blk_a.p, s->blk_b.p, s->S.p, s->rhs.p, s->vx.p, s->fail.p); MM_LAUNCH_CHECK();
    return pcg_broadcast(s);
  }
  if (s->pcg_grid == 0) {
    // choose between the cached kernel (rows resident in shared memory) and the streaming kernel
    std::vector<int> h_rs((size_t)s->n_img + 1);
    MM_CUDA(cudaMemcpyAsync(h_rs.data(), s->row_start.p, sizeof(int) * h_rs.size(), cudaMemcpyDeviceToHost, st));
    MM_CUDA(cudaStreamSynchronize(st));
    int dev = 0, max_smem = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    s->pcg_cached = false;
    const int tries[3] = { 128, 256, 512 };       // smallest CTA whose grid is still co-resident: more SMs share the rows (cfg2: 406 / 392 / 341 LM it/s)
    for (int ti = 0; ti < 3 && !s->pcg_cached && !getenv("MM_PCG_STREAMING") && !s->refine; ++ti) {
      const int threads = tries[ti], NW = threads / 32;
      if (getenv("MM_PCG_THREADS") && atoi(getenv("MM_PCG_THREADS")) != threads) continue;
      const int blocks = (s->n_img + NW - 1) / NW;
      int e_cap = 1;
      for (int b = 0; b < blocks; ++b) e_cap = std::max(e_cap, h_rs[std::min((b + 1) * NW, s->n_img)] - h_rs[b * NW]);
      const size_t smem = sizeof(double) * ((size_t)e_cap * 36 + NW * 36) + sizeof(int) * (size_t)e_cap + 16;
      if (smem + 2048 > (size_t)max_smem) continue;
      const void* fn = threads == 512 ? (const void*)k_pcg_cached<512> : (threads == 256 ? (const void*)k_pcg_cached<256> : (const void*)k_pcg_cached<128>);
      int per_sm = 0;
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&per_sm, fn, threads, smem, 0) == cudaSuccess && per_sm * num_sms() >= blocks) {
        s->pcg_cached = true; s->pcg_grid = std::max(1, blocks); s->pcg_smem = smem; s->pcg_ecap = e_cap; s->pcg_threads = threads; s->pcg_fn = fn;
      }
      cudaGetLastError();
    }
    if (!s->pcg_cached) {
      int per_sm = 0;
      MM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_persistent, 256, 0));
      const int cap = std::max(1, per_sm * num_sms());
      s->pcg_grid = std::max(1, std::min(cap, (s->n_img + 7) / 8));
    }
  }
  PcgArgs a; a.n_img = s->n_img; a.row_start = s->row_start.p; a.row_col = s->row_col.p; a.row_blk = s->row_blk.p; a.S = s->S.p; a.Minv = s->Minv.p;
  a.b = s->rhs.p; a.x = s->vx.p; a.r = s->vr.p; a.z = s->vz.p; a.p0 = s->vp0.p; a.p1 = s->vp1.p; a.Ap = s->vAp.p; a.sc = s->pcg_sc.p; a.ic = s->pcg_ic.p;
  a.tol2 = s->opt.pcg_tolerance * s->opt.pcg_tolerance; a.max_iter = s->opt.pcg_max_iterations;
  a.dbg = nullptr;
  a.n_intr = 0; a.Bm = nullptr; a.Cm = nullptr; a.Cinv = nullptr; a.bi = nullptr; a.xi = a.zi = a.pi0 = a.pi1 = a.bt = nullptr;      // (refined intrinsics always go through the tile factorisation)
  a.cm = s->cm; a.agg = s->agg.p; a.Pc = s->Pc.p; a.Ainv = s->Ac.p; a.rc = s->crc.p; a.qc = s->cqc.p; a.yc = s->cyc.p;
  if (s->cm) { MM_CUDA(cudaMemsetAsync(s->crc.p, 0, sizeof(double) * 2 * (size_t)s->cm, st)); MM_CUDA(cudaMemsetAsync(s->cqc.p, 0, sizeof(double) * (size_t)s->cm, st)); }
  if (getenv("MM_PCG_DEBUG")) { if (!s->pcg_dbg.p) MM_CUDA(s->pcg_dbg.alloc(6 * 32)); a.dbg = s->pcg_dbg.p; }
  if (s->pcg_cached) {
    int e_cap = s->pcg_ecap;
    void* cargs[] = { &a, &e_cap };
    MM_CUDA(cudaLaunchCooperativeKernel(s->pcg_fn, dim3(s->pcg_grid), dim3(s->pcg_threads), cargs, s->pcg_smem, st));
    count_launch();
    return pcg_broadcast(s);
  }
  void* args[] = { &a };
  MM_CUDA(cudaLaunchCooperativeKernel((void*)k_pcg_persistent, dim3(s->pcg_grid), dim3(256), args, 0, st));
  count_launch();
  return pcg_broadcast(s);
}
// K4: back-substitution, candidate parameters, step norm and model cost change -> red[3], red[4]
int launch_update(mm_ba_session* s) {
  cudaStream_t st = s->stream;
  const int P0 = s->p_lo, NP = s->np_loc();
  const int gp_ = blocks_for(NP, 128), gc_ = blocks_for((int64_t)6 * s->n_img, 128);
  k_backsub<<<gp_, 128, 0, st>>>(NP, s->pt_start.p + P0, s->obs_img.p, s->rec_base(), s->scale_c.p, s->scale_p.p + 3 * (size_t)P0, s->Vinv.p + 6 * (size_t)P0,
      s->gp.p + 3 * (size_t)P0, s->dp.p + 3 * (size_t)P0, s->vx.p, s->pts.p + 3 * (size_t)P0, s->pts2.p + 3 * (size_t)P0, s->part_pt.p,
      s->refine ? s->Apc.p + 27 * (size_t)P0 * s->n_cam : nullptr, s->refine ? s->xi.p : nullptr, s->n_cam, s->refine ? s->pt_cam_mask.p + P0 : nullptr); MM_LAUNCH_CHECK();
  k_update_cam<<<gc_, 128, 0, st>>>(6 * s->n_img, s->vx.p, s->scale_c.p, s->gc.p, s->dc.p, s->poses.p, s->poses2.p, s->part_cam.p); MM_LAUNCH_CHECK();
  if (s->refine) { k_update_intr<<<1, 32, 0, st>>>(s->n_cam, s->xi.p, s->scale_i.p, s->gi.p, s->di.p, s->intr.p, s->intr2.p, s->part_cam.p + 2 * (size_t)gc_); MM_LAUNCH_CHECK(); }
  // the (replicated) camera part of |step|^2 and of the model cost change is counted on rank 0 only
  k_reduce_pairs<<<1, 256, 0, st>>>(s->part_pt.p, gp_, s->part_cam.p, s->rank == 0 ? gc_ + (s->refine ? 1 : 0) : 0, s->loc.p + 3); MM_LAUNCH_CHECK();
  return MM_OK;
}
int launch_xnorm(mm_ba_session* s) {
  cudaStream_t st = s->stream;
  k_xnorm<<<s->grid_x, 256, 0, st>>>(s->rank == 0 ? 6 * s->n_img : 0, s->poses.p, s->pose_mask.p, s->np_loc(), s->pts.p + 3 * (size_t)s->p_lo, s->pt_mask.p + s->p_lo, s->part_x.p,
      s->refine && s->rank == 0 ? s->intr.p : nullptr, s->intr_mask.p, s->n_cam); MM_LAUNCH_CHECK();
  k_reduce_sum<<<1, 256, 0, st>>>(s->part_x.p, s->grid_x, s->loc.p + 5); MM_LAUNCH_CHECK();
  return MM_OK;
}
int read_red(mm_ba_session* s, double* out8) {
  MM_CUDA(cudaMemcpyAsync(out8, s->red.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, s->stream));
  MM_CUDA(cudaStreamSynchronize(s->stream));
  return MM_OK;
}
void trace_push(mm_ba_session* s, int accepted, int lin) {
  mm_ba_summary& S = s->sum; const int i = S.num_iterations;
  if (i < MM_BA_TRACE_MAX) { S.trace_cost[i] = s->cost; S.trace_radius[i] = s->radius; S.trace_gradient_max_norm[i] = s->gmax; S.trace_accepted[i] = accepted; S.trace_linear_iterations[i] = lin; }
  S.num_iterations = i + 1;
  if (s->cost < S.final_cost) S.final_cost = s->cost;     // SetSummaryFinalCost: min over the recorded iterations
}

// initial evaluation (iteration 0 of Ceres' minimizer)
int lm_start(mm_ba_session* s) {
  mm_ba_summary& S = s->sum;
  memset(&S, 0, sizeof S);
  S.num_residuals = 2 * s->n_obs + s->n_prior; S.final_cost = INFINITY; S.termination = MM_TERM_NO_CONVERGENCE;
  s->radius = s->opt.initial_trust_region_radius; s->decrease_factor = 2.0; s->iter = 0; s->n_invalid = 0;
  s->finished = false; s->started = true; s->scaled = false; s->coarse_iter = -1; s->coarse_radius = 0.0;
  if (s->n_obs == 0) { S.termination = MM_TERM_EMPTY; S.initial_cost = S.final_cost = 0.0; S.return_value = NAN; s->finished = true; return MM_OK; }
  int rc;
  { Timer t(s, &S.ms_linearize);
    if ((rc = launch_linearize(s))) return rc;
    k_fill<<<grid_stride(6 * (int64_t)s->n_img, 256), 256, 0, s->stream>>>(6 * (int64_t)s->n_img, s->scale_c.p, 1.0); MM_LAUNCH_CHECK();
    k_fill<<<grid_stride(3 * (int64_t)s->n_pt, 256), 256, 0, s->stream>>>(3 * (int64_t)s->n_pt, s->scale_p.p, 1.0); MM_LAUNCH_CHECK();
    if ((rc = launch_scale(s))) return rc; }
  if ((rc = build_coarse_basis(s))) return rc;
  if ((rc = launch_xnorm(s))) return rc;                     // before the Schur pass: its exchange step carries cost and |x|^2
  { Timer t(s, &S.ms_schur); if ((rc = launch_schur(s))) return rc; }
  double red[8]; if ((rc = read_red(s, red))) return rc;
  s->cost = red[0]; s->gmax = red[2]; s->x_norm = sqrt(red[5]);
  S.initial_cost = s->cost;
  trace_push(s, 1, 0);
  s->abs_gtol = s->opt.gradient_tolerance * fmax(s->gmax, 1e-12);
  if (!isfinite(s->cost)) { S.termination = MM_TERM_NUMERICAL_FAILURE; s->finished = true; set_error("non-finite initial cost"); return MM_ERR_NUMERICAL; }
  if (s->gmax <= s->abs_gtol) { S.termination = MM_TERM_GRADIENT_TOLERANCE; s->finished = true; }
  // the iterate at which we stand is the best one so far: poses/pts already hold it (x_min == x)
  return MM_OK;
}

// one LM iteration (one linear solve + candidate evaluation + accept/reject)
int lm_iterate(mm_ba_session* s) {
  mm_ba_summary& S = s->sum; const mm_ba_options& O = s->opt;
  if (s->iter >= O.max_num_iterations) { S.termination = MM_TERM_NO_CONVERGENCE; s->finished = true; return MM_OK; }
  s->iter++;
  int rc, pcg_it = 0, fail = 0;
  cudaEventRecord(s->evs[0], s->stream);
  if ((rc = launch_pcg(s))) return rc;
  cudaEventRecord(s->evs[1], s->stream);
  if ((rc = launch_update(s))) return rc;
  if ((rc = launch_cost_candidate(s))) return rc;
  cudaEventRecord(s->evs[2], s->stream);
  double red[8];
  { int hic[4];
    MM_CUDA(cudaMemcpyAsync(hic, s->pcg_ic.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, s->stream));
    MM_CUDA(cudaMemcpyAsync(&fail, s->fail.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    if ((rc = read_red(s, red))) return rc;           // synchronises the stream
    pcg_it = hic[1];
    if (s->pcg_dbg.p) { unsigned long long h[192]; cudaMemcpy(h, s->pcg_dbg.p, sizeof h, cudaMemcpyDeviceToHost); for (int i = 2; i < 6; ++i) printf("pcg(%s,%d thr,%d ctas) it %d: spmv %llu red %llu sync %llu upd %llu sync %llu  next %llu ns\n", s->pcg_cached ? "cached" : "streaming", s->pcg_threads, s->pcg_grid, i, h[6*i+1]-h[6*i], h[6*i+2]-h[6*i+1], h[6*i+3]-h[6*i+2], h[6*i+4]-h[6*i+3], h[6*i+5]-h[6*i+4], h[6*(i+1)]-h[6*i+5]); }
    if (fail) MM_CUDA(cudaMemsetAsync(s->fail.p, 0, sizeof(int), s->stream));
    float ms = 0; cudaEventElapsedTime(&ms, s->evs[0], s->evs[1]); S.ms_pcg += ms; cudaEventElapsedTime(&ms, s->evs[1], s->evs[2]); S.ms_update += ms; }
  s->last_pcg_iters = pcg_it;
  const double new_cost = red[1], step_norm = sqrt(red[3]), mcc = red[4];
  bool valid = !fail && isfinite(step_norm) && isfinite(mcc) && isfinite(new_cost) && !(mcc < 0.0);
  bool successful = false; double rel_dec = 0.0;
  if (getenv("MM_BA_DIAG") && (!valid || !(s->cost - new_cost > O.min_relative_decrease * mcc)))      // a step that is about to be refused: say why
    fprintf(stderr, "[mm_ba] iteration %d refused: fail flag %d, cost %.9e -> %.9e, model change %.6e, |step| %.6e, radius %.3e, %d PCG iterations\n",
            s->iter, fail, s->cost, new_cost, mcc, step_norm, s->radius, pcg_it);
  if (!valid) {
    if (++s->n_invalid >= O.max_num_consecutive_invalid_steps) { S.termination = MM_TERM_NUMERICAL_FAILURE; s->finished = true; return MM_OK; }
  } else {
    s->n_invalid = 0;
    if (step_norm <= O.parameter_tolerance * (s->x_norm + O.parameter_tolerance)) { S.termination = MM_TERM_PARAMETER_TOLERANCE; s->finished = true; return MM_OK; }
    const double cost_change = s->cost - new_cost;
    if (fabs(cost_change) < O.function_tolerance * s->cost) { S.termination = MM_TERM_FUNCTION_TOLERANCE; s->finished = true; return MM_OK; }
    rel_dec = cost_change / mcc;
    successful = rel_dec > O.min_relative_decrease;
  }
  if (successful) {
    S.num_successful_steps++;
    s->radius = s->radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rel_dec - 1.0, 3));
    s->radius = fmin(O.max_trust_region_radius, s->radius);
    s->decrease_factor = 2.0;
    // keep the previous iterate in poses2/pts2 until the gradient test has passed (Ceres 1.8 commits
    // x_min only after it)
    std::swap(s->poses.p, s->poses2.p); std::swap(s->pts.p, s->pts2.p); std::swap(s->aux.p, s->aux2.p); if (s->refine) std::swap(s->intr.p, s->intr2.p);
    cudaEventRecord(s->evs[3], s->stream);
    if ((rc = launch_linearize(s, !getenv("MM_K1_RECOMPUTE_COST")))) return rc;
    if ((rc = launch_xnorm(s))) return rc;
    cudaEventRecord(s->evs[4], s->stream);
    if ((rc = launch_schur(s))) return rc;
    cudaEventRecord(s->evs[5], s->stream);
    if ((rc = read_red(s, red))) return rc;
    { float ms = 0; cudaEventElapsedTime(&ms, s->evs[3], s->evs[4]); S.ms_linearize += ms; cudaEventElapsedTime(&ms, s->evs[4], s->evs[5]); S.ms_schur += ms; }
    s->cost = red[0]; s->gmax = red[2]; s->x_norm = sqrt(red[5]);
    if (s->gmax <= s->abs_gtol) {
      std::swap(s->poses.p, s->poses2.p); std::swap(s->pts.p, s->pts2.p); std::swap(s->aux.p, s->aux2.p); if (s->refine) std::swap(s->intr.p, s->intr2.p);   // discard, as Ceres 1.8 does
      S.termination = MM_TERM_GRADIENT_TOLERANCE; s->finished = true; return MM_OK;
    }
  } else {
    S.num_unsuccessful_steps++;
    s->radius = s->radius / s->decrease_factor; s->decrease_factor *= 2.0;
    if (s->radius >= O.min_trust_region_radius) {
      cudaEventRecord(s->evs[4], s->stream);
      if ((rc = launch_schur(s))) return rc;
      cudaEventRecord(s->evs[5], s->stream); cudaEventSynchronize(s->evs[5]);
      float ms = 0; cudaEventElapsedTime(&ms, s->evs[4], s->evs[5]); S.ms_schur += ms;
    }
  }
  if (s->radius < O.min_trust_region_radius) { S.termination = MM_TERM_PARAMETER_TOLERANCE; s->finished = true; return MM_OK; }
  trace_push(s, successful ? 1 : 0, pcg_it);
  if (O.print_progress)
    printf("% 4d: f:% 8e new:% 8e g:% 3.2e rho:% 3.2e mu:% 3.2e li:% 3d ok:%d\n", s->iter, s->cost, new_cost, s->gmax, rel_dec, s->radius, pcg_it, (int)successful);
  return MM_OK;
}

}  // namespace

extern "C" {

void mm_ba_options_default(mm_ba_options* o) {
  if (!o) return;
  memset(o, 0, sizeof *o);
  o->max_num_iterations = 100; o->function_tolerance = 1e-4; o->gradient_tolerance = 1e-8;   // bundle_adjustment.h:40-42
  o->loss_type = MM_LOSS_CAUCHY; o->loss_scale = 1.0;                                          // :45
  o->parameter_tolerance = 1e-8; o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32; o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->jacobi_scaling = 1; o->max_num_consecutive_invalid_steps = 10;
  o->linear_solver = MM_SOLVER_PCG; o->pcg_tolerance = 1e-13; o->pcg_max_iterations = 2000; o->print_progress = 0;
  o->pcg_preconditioner = MM_PRECOND_AUTO; o->tile_cholesky_tolerance = 1e-8;
}

void mm_ba_session_destroy(mm_ba_session* s) {
  if (!s) return;
  if (s->tc_thread.joinable()) s->tc_thread.join();
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  for (int i = 0; i < 8; ++i) if (s->evs[i]) cudaEventDestroy(s->evs[i]);
  delete s;
}

namespace {
int session_create(const mm_ba_problem* P, const mm_ba_options* opt, void* stream, int32_t rank, int32_t world,
                   mm_allreduce_fn allreduce, void* user, bool keep_host_pts, mm_ba_session** out) {
  if (!out || !opt) { set_error("null argument"); return MM_ERR_INVALID_ARG; }
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world || (world > 1 && !allreduce)) { set_error("invalid shard (rank %d of %d) or missing all-reduce callback", rank, world); return MM_ERR_INVALID_ARG; }
  int rc = validate_problem(P); if (rc) return rc;
  if (opt->linear_solver != MM_SOLVER_PCG) { set_error("the device engine solves the reduced system with PCG only"); return MM_ERR_UNSUPPORTED; }
  std::vector<char> cam_used((size_t)std::max(P->n_cam, 1), 0);
  for (int i = 0; i < P->n_img; ++i) cam_used[P->img_cam[i]] = 1;
  bool refine = false;
  for (int c = 0; c < P->n_cam; ++c) if (!P->intr_const[c] && cam_used[c]) refine = true;
  if (refine && P->n_cam > 32) { set_error("refine_camera_params handles up to 32 cameras (got %d)", P->n_cam); return MM_ERR_UNSUPPORTED; }
  rc = ensure_device(); if (rc) return rc;
  const double t_enter = SetupClock::now();
  mm_ba_session* s = new mm_ba_session();
  s->stream = (cudaStream_t)stream; s->opt = *opt; s->keep_host_pts = keep_host_pts;
  s->n_img = P->n_img; s->n_cam = P->n_cam; s->n_pt = P->n_pt; s->n_obs = P->n_obs; s->refine = refine;
  s->rank = rank; s->world = world; s->ar = allreduce; s->ar_user = user;
  auto fail_out = [&](int code) { mm_ba_session_destroy(s); return code; };
  if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess) { set_error("cudaEventCreate failed"); return fail_out(MM_ERR_CUDA); }
  for (int i = 0; i < 8; ++i) if (cudaEventCreate(&s->evs[i]) != cudaSuccess) { set_error("cudaEventCreate failed"); return fail_out(MM_ERR_CUDA); }
  SetupClock clk(s->stream);
  struct ClockScope { ClockScope(SetupClock* c) { g_setup_clock = c->on ? c : nullptr; } ~ClockScope() { g_setup_clock = nullptr; } } clock_scope(&clk);
  s->h_poses0.assign(P->poses, P->poses + 6 * (size_t)P->n_img);
  s->h_intr0.assign(P->intr, P->intr + MM_INTR_STRIDE * (size_t)P->n_cam);
  if (keep_host_pts) s->h_pts0.assign(P->pts, P->pts + 3 * (size_t)P->n_pt);      // (mm_ba_session_reset starts again from them)
  const size_t n_img = (size_t)std::max(P->n_img, 1), n_pt = (size_t)std::max(P->n_pt, 1), n_cam = (size_t)std::max(P->n_cam, 1);
#define A(buf, count) do { if ((buf).alloc(count) != cudaSuccess) { set_error("cudaMalloc failed for " #buf); cudaGetLastError(); return fail_out(MM_ERR_ALLOC); } } while (0)
  A(s->poses, 6 * n_img); A(s->poses2, 6 * n_img); A(s->intr, MM_INTR_STRIDE * n_cam); A(s->pts, 3 * n_pt); A(s->pts2, 3 * n_pt);
  A(s->aux, AUX * n_img); A(s->aux2, AUX * n_img);
  A(s->pose_mask, 6 * n_img); A(s->pt_mask, n_pt); A(s->scale_c, 6 * n_img); A(s->scale_p, 3 * n_pt);
  A(s->Vinv, 6 * n_pt); A(s->pinfo, PINFO * n_pt); A(s->gp, 3 * n_pt); A(s->dp, 3 * n_pt); A(s->dc, 6 * n_img);
  A(s->loc, 8); A(s->stepbuf, 8);
  A(s->vx, 6 * n_img + 9 * n_cam); A(s->vr, 6 * n_img); A(s->vz, 6 * n_img); A(s->vp0, 6 * n_img); A(s->vp1, 6 * n_img); A(s->vAp, 6 * n_img);
  A(s->pcg_sc, 16); A(s->pcg_ic, 4); A(s->red, 8); A(s->fail, 1); A(s->img_cam, n_img); A(s->cam_model, n_cam); A(s->Minv, 36 * n_img);
  MM_CUDA(cudaFuncSetAttribute((const void*)k_residual_jacobian<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM));
  MM_CUDA(cudaFuncSetAttribute((const void*)k_residual_jacobian<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM));
  MM_CUDA(cudaFuncSetAttribute((const void*)k_residual_jacobian<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM));
  MM_CUDA(cudaFuncSetAttribute((const void*)k_residual_jacobian<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM));
  MM_CUDA(cudaFuncSetAttribute((const void*)k_residual_jacobian<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM_COST));
  s->grid_x = grid_stride(std::max<int64_t>(3 * (int64_t)P->n_pt, 1), 256);
  A(s->part_cost, (size_t)grid_stride(std::max<int64_t>(P->n_obs, 1), 256)); A(s->part_pt, 2 * (size_t)blocks_for(P->n_pt, 128)); A(s->part_cam, 2 * (size_t)blocks_for(6 * (int64_t)P->n_img, 128) + 2);
  A(s->part_x, (size_t)s->grid_x);
  A(s->intr2, MM_INTR_STRIDE * n_cam); A(s->intr_mask, 9 * n_cam); A(s->scale_i, 9 * n_cam);
  if (refine) {
    // exchange buffer of the border: B | C, b, diag, gradient  (contiguous: one all-reduce)
    const size_t nB = 54 * n_img * n_cam, nA = 81 * n_cam * n_cam + 27 * n_cam;
    s->xch2_count = nB + nA;
    A(s->xch2, s->xch2_count); s->Bm.view(s->xch2.p, nB); s->intr_acc.view(s->xch2.p + nB, nA);
    A(s->gi, 9 * n_cam); A(s->di, 9 * n_cam); s->xi.view(s->vx.p + 6 * n_img, 9 * n_cam); A(s->img_sq, 9 * n_img);
  }
  { // free intrinsics: the parameters of the camera's model, for cameras that are used and not held constant
    std::vector<double> im(9 * n_cam, 0.0), ones(9 * n_cam, 1.0);
    for (int c = 0; c < P->n_cam; ++c)
      if (refine && cam_used[c] && !P->intr_const[c]) for (int k = 0; k < model_num_params(P->cam_model[c]); ++k) im[9 * (size_t)c + k] = 1.0;
    if (cudaMemcpy(s->intr_mask.p, im.data(), sizeof(double) * im.size(), cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(s->scale_i.p, ones.data(), sizeof(double) * ones.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(s->img_cam.p, P->img_cam, sizeof(int) * (size_t)P->n_img, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(s->cam_model.p, P->cam_model, sizeof(int) * (size_t)P->n_cam, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("mask upload failed"); return fail_out(MM_ERR_CUDA); } }
  if (P->rot_prior && P->rot_prior_w) {
    for (int i = 0; i < P->n_img; ++i) if (P->rot_prior_w[i] != 0.0) s->n_prior++;
    if (s->n_prior) {
      A(s->pr_rot0, 3 * n_img); A(s->pr_w, n_img); A(s->pr_r, n_img); A(s->pr_J, 3 * n_img);
      if (cudaMemcpy(s->pr_rot0.p, P->rot_prior, sizeof(double) * 3 * (size_t)P->n_img, cudaMemcpyHostToDevice) != cudaSuccess ||
          cudaMemcpy(s->pr_w.p, P->rot_prior_w, sizeof(double) * (size_t)P->n_img, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("constraint upload failed"); return fail_out(MM_ERR_CUDA); }
    }
  }
  cudaMemsetAsync(s->fail.p, 0, sizeof(int), s->stream);
  cudaMemsetAsync(s->red.p, 0, sizeof(double) * 8, s->stream);
  cudaMemsetAsync(s->loc.p, 0, sizeof(double) * 8, s->stream);
  SETUP_MARK("allocations");
  // structure on the device (also: index check, masks, point renumbering); the symbolic analysis of the tile Cholesky starts
  // on a host thread as soon as the block list of the reduced system exists and runs beside everything up to the first solve
  rc = build_structure(s, P); if (rc) return fail_out(rc);
  rc = tc_begin(s); if (rc) return fail_out(rc);
  SETUP_MARK("tilechol: block list -> host thread");
  rc = upload_obs_xy(s, P); if (rc) return fail_out(rc);
  SETUP_MARK("image coordinates");
  rc = build_block_csr(s); if (rc) return fail_out(rc);
  if (s->tc_plan_rc == 1) { rc = build_aggregates(s); if (rc) return fail_out(rc); }      // no tile factorisation for this session
  rc = build_shard(s); if (rc) return fail_out(rc);
  A(s->rec, (size_t)REC * (size_t)std::max<int64_t>(s->no_loc(), 1));
  if (refine) { A(s->ji, 18 * (size_t)std::max<int64_t>(s->no_loc(), 1)); A(s->Apc, 27 * n_pt * n_cam); A(s->pt_cam_mask, n_pt); }
  s->grid_obs = grid_stride(std::max<int64_t>(s->no_loc(), 1), 256);
  { // the exchange buffer: S | rhs | gc | ud | scalars, contiguous so that one all-reduce moves it
    const size_t nS = 36 * (size_t)std::max<int64_t>(s->nblk, 1), nv = 6 * n_img;
    s->scal_off = nS + 3 * nv; s->xch_count = s->scal_off + 3 + (size_t)world;
    A(s->xch, s->xch_count);
    s->S.view(s->xch.p, nS); s->rhs.view(s->xch.p + nS, nv); s->gc.view(s->xch.p + nS + nv, nv); s->ud.view(s->xch.p + nS + 2 * nv, nv);
    cudaMemsetAsync(s->xch.p, 0, sizeof(double) * s->xch_count, s->stream); }
#undef A
  rc = upload_params(s, P->pts); if (rc) return fail_out(rc);
  SETUP_MARK("block CSR + records + parameters");
  memset(&s->sum, 0, sizeof s->sum);
  rc = lm_start(s); if (rc) return fail_out(rc);
  SETUP_MARK("lm_start (first linearisation)");
  { bool fallback = false;
    rc = tc_finish(s, &fallback); if (rc) return fail_out(rc);
    if (fallback && s->tc_plan_rc != 1 && !s->finished) {
      // (rare) the tile factorisation was dropped after the first linearisation: two-level preconditioner on the assembled system
      if ((rc = build_aggregates(s)) || (rc = build_coarse_basis(s))) return fail_out(rc);
      k_precond<<<blocks_for(s->n_img, 64), 64, 0, s->stream>>>(s->n_img, s->S.p, s->Minv.p, s->fail.p); count_launch();
      if ((rc = launch_coarse_setup(s))) return fail_out(rc);
      s->coarse_iter = s->iter; s->coarse_radius = s->radius;
    } }
  MM_CUDA(cudaStreamSynchronize(s->stream));
  s->sum.ms_setup = std::max(0.0, SetupClock::now() - t_enter - s->sum.ms_linearize - s->sum.ms_schur);
  *out = s;
  return MM_OK;
}
}  // namespace

int mm_ba_session_create(const mm_ba_problem* P, const mm_ba_options* opt, void* stream, mm_ba_session** out) {
  return session_create(P, opt, stream, 0, 1, nullptr, nullptr, true, out);
}

int mm_ba_session_create_sharded(const mm_ba_problem* P, const mm_ba_options* opt, void* stream, int32_t rank, int32_t world,
                                 mm_allreduce_fn allreduce, void* user, mm_ba_session** out) {
  return session_create(P, opt, stream, rank, world, allreduce, user, true, out);
}

int mm_ba_session_reset(mm_ba_session* s) {
  if (!s) return MM_ERR_INVALID_ARG;
  if (!s->keep_host_pts) { set_error("this session keeps no host copy of the initial points"); return MM_ERR_UNSUPPORTED; }
  int rc = upload_params(s, s->h_pts0.data()); if (rc) return rc;
  const double setup = s->sum.ms_setup;
  rc = lm_start(s);
  s->sum.ms_setup = setup;
  return rc;
}

int mm_ba_session_iterate(mm_ba_session* s, int32_t n, int32_t* n_done) {
  if (!s) return MM_ERR_INVALID_ARG;
  int done = 0;
  for (int i = 0; i < n && !s->finished; ++i) {
    const int before = s->iter;
    int rc = lm_iterate(s); if (rc) { if (n_done) *n_done = done; return rc; }
    if (s->iter > before) ++done;
  }
  s->sum.return_value = sqrt(s->sum.final_cost / (double)std::max<int64_t>(s->sum.num_residuals, 1));
  if (s->sum.num_residuals == 0) s->sum.return_value = NAN;
  s->sum.ms_total = s->sum.ms_setup + s->sum.ms_linearize + s->sum.ms_schur + s->sum.ms_pcg + s->sum.ms_update;
  if (n_done) *n_done = done;
  return MM_OK;
}

int mm_ba_session_summary(mm_ba_session* s, mm_ba_summary* out) {
  if (!s || !out) return MM_ERR_INVALID_ARG;
  *out = s->sum;
  return MM_OK;
}

int64_t mm_ba_session_num_blocks(mm_ba_session* s) { return s ? s->nblk : -1; }
int mm_ba_session_solver_info(mm_ba_session* s, double* out8) {
  if (!s || !out8) return MM_ERR_INVALID_ARG;
  for (int k = 0; k < 8; ++k) out8[k] = 0.0;
  const TileCholPlan& P = s->tc_plan;
  if (s->tc_on) { out8[0] = 2; out8[1] = (double)P.n_l; out8[2] = (double)(P.n_upd + P.n_wupd); out8[3] = P.flops; out8[4] = P.nt; out8[5] = (double)P.n_w; out8[6] = P.n_stasks;
                  out8[7] = sizeof(double) * (double)TC_TT * (2.0 * (double)P.n_l + 2.0 * (double)P.n_w) / 1e6; }
  else if (s->cm) { out8[0] = 1; out8[4] = s->cm; }
  return MM_OK;
}
int32_t mm_ba_session_coarse_dim(mm_ba_session* s) { return s ? s->cm : -1; }

int mm_debug_spd_inverse(double* a, int32_t m) {
  if (!a || m <= 0) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  int rc = ensure_device(); if (rc) return rc;
  DevBuf<double> A, Ck, Rk;
  MM_CUDA(A.alloc((size_t)m * m)); MM_CUDA(Ck.alloc((size_t)m * GJ_NB)); MM_CUDA(Rk.alloc((size_t)m * GJ_NB));
  MM_CUDA(cudaMemcpy(A.p, a, sizeof(double) * (size_t)m * m, cudaMemcpyHostToDevice));
  int per_sm = 0;
  MM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_spd_inverse, 256, 0));
  const int tiles = (m + 63) / 64;
  const int grid = std::max(1, std::min(std::max(1, per_sm) * num_sms(), std::max(tiles * tiles, (m + 255) / 256)));
  DevBuf<unsigned long long> dbgb; unsigned long long* dbg = nullptr;
  if (getenv("MM_GJ_DEBUG")) { MM_CUDA(dbgb.alloc(64)); MM_CUDA(cudaMemset(dbgb.p, 0, 64 * sizeof(unsigned long long))); dbg = dbgb.p; }
  int mm_ = m; double* Ap = A.p; double* Cp = Ck.p; double* Rp = Rk.p;
  void* args[] = { &mm_, &Ap, &Cp, &Rp, &dbg };
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, nullptr);
  MM_CUDA(cudaLaunchCooperativeKernel((const void*)k_spd_inverse, dim3(grid), dim3(256), args, 0, nullptr));
  cudaEventRecord(e1, nullptr);
  count_launch();
  MM_CUDA(cudaDeviceSynchronize());
  if (dbg) {
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[64]; cudaMemcpy(h, dbg, sizeof h, cudaMemcpyDeviceToHost);
    printf("spd_inverse m=%d grid=%d: %.3f ms\n", m, grid, ms);
    for (int k = 1; k < 4 && 32 * k < m; ++k) printf("  step %d: pivot %llu panel %llu sync %llu tiles %llu sync %llu ns\n", k, h[6*k+1]-h[6*k], h[6*k+2]-h[6*k+1], h[6*k+3]-h[6*k+2], h[6*k+4]-h[6*k+3], h[6*k+5]-h[6*k+4]);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  MM_CUDA(cudaMemcpy(a, A.p, sizeof(double) * (size_t)m * m, cudaMemcpyDeviceToHost));
  return MM_OK;
}


/* ---- test hooks of the sparse tile Cholesky ---------------------------------------------------------------- */
struct mm_tilechol_plan { mm::TileCholPlan P; };

int mm_debug_tilechol_plan_create(int32_t n_img, int32_t n_off, const int32_t* blk_a, const int32_t* blk_b, const double* pos,
                                  int32_t n_cam_border, mm_tilechol_plan** out) {
  if (!out || n_img <= 0 || n_off < 0 || (n_off > 0 && (!blk_a || !blk_b))) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  for (int e = 0; e < n_off; ++e) if (blk_a[e] < 0 || blk_a[e] >= n_img || blk_b[e] < 0 || blk_b[e] >= n_img) { set_error("block index out of range"); return MM_ERR_INVALID_ARG; }
  mm_tilechol_plan* h = new mm_tilechol_plan();
  const int rc = build_tilechol_plan(n_img, n_off, blk_a, blk_b, pos, n_cam_border, h->P);
  if (rc) { delete h; set_error("tile Cholesky plan failed (%d)", rc); return MM_ERR_UNSUPPORTED; }
  *out = h;
  return MM_OK;
}
void mm_debug_tilechol_plan_destroy(mm_tilechol_plan* h) { delete h; }
int64_t mm_debug_tilechol_plan_array(mm_tilechol_plan* h, int32_t which, int64_t* out, int64_t cap) {
  if (!h) return -1;
  const TileCholPlan& P = h->P;
  auto copy = [&](const auto& v) -> int64_t { const int64_t n = (int64_t)v.size(); if (out) for (int64_t i = 0; i < n && i < cap; ++i) out[i] = (int64_t)v[i]; return n; };
  switch (which) {
    case 0: { const std::vector<int64_t> sc = { P.nt_pose, P.nt, P.n_l, P.n_upd, P.n_nodes, P.max_height, (int64_t)P.flops, TC_T, TC_TI, P.n_w, P.n_wtask, P.n_slots, P.n_stasks, P.n_tnodes }; return copy(sc); }
    case 1: return copy(P.img_tile);   case 2: return copy(P.img_slot);  case 3: return copy(P.tile_nunk); case 4: return copy(P.col_ptr);
    case 5: return copy(P.row_idx);    case 6: return copy(P.col_idx);   case 7: return copy(P.has_a);     case 8: return copy(P.upd_ptr);
    case 9: return copy(P.upd_a);      case 10: return copy(P.upd_b);    case 11: return copy(P.rowp_ptr); case 12: return copy(P.rowp_tile);
    case 13: return copy(P.rowp_col);  case 14: return copy(P.unk_of);   case 15: return copy(P.sc_tile);  case 16: return copy(P.sc_off);
    case 17: return copy(P.tile_height);
    case 18: return copy(P.tile_node);  case 19: return copy(P.node_first); case 20: return copy(P.node_nt);   case 21: return copy(P.w_row_ptr);
    case 22: return copy(P.wt_row);     case 23: return copy(P.wt_col);     case 24: return copy(P.wt_store);  case 25: return copy(P.wupd_ptr);
    case 26: return copy(P.wupd_l);     case 27: return copy(P.wupd_w);     case 28: return copy(P.wupd_flag); case 29: return copy(P.task_order);
    case 30: return copy(P.st_kind);    case 31: return copy(P.st_out);     case 32: return copy(P.st_base);   case 33: return copy(P.st_tile);
    case 34: return copy(P.st_item_ptr); case 35: return copy(P.it_mat);    case 36: return copy(P.it_src);
    default: return -1;
  }
}
/* factorisation + one application z = M^-1 rhs on the device for a block-sparse SPD matrix given like the reduced system:
 * S [(n_img + n_off) * 36] (diagonal blocks first, then the blocks (blk_a < blk_b) row-major), optional border Bm [n_img][ncb][6][9],
 * Cm [(9 ncb)^2]; rhs and z have 6 n_img + 9 ncb entries.  Host buffers. */
int mm_debug_tilechol_solve(int32_t n_img, int32_t n_off, const int32_t* blk_a, const int32_t* blk_b, const double* pos, const double* S,
                            int32_t ncb, const double* Bm, const double* Cm, const double* rhs, double* z, int32_t reps, double* ms_factor, double* ms_apply) {
  if (n_img <= 0 || n_off < 0 || !S || !rhs || !z || ncb < 0 || (ncb > 0 && (!Bm || !Cm))) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  int rc = ensure_device(); if (rc) return rc;
  mm_ba_session* s = new mm_ba_session();
  struct Guard { mm_ba_session* s; ~Guard() { delete s; } } guard{ s };
  mm_ba_options_default(&s->opt); s->opt.pcg_preconditioner = MM_PRECOND_TILE_CHOLESKY;
  s->n_img = n_img; s->n_off = n_off; s->nblk = (int64_t)n_img + n_off; s->n_obs = 1; s->refine = ncb > 0; s->n_cam = ncb;
  s->h_poses0.assign(6 * (size_t)n_img, 0.0);
  if (pos) for (int i = 0; i < n_img; ++i) for (int q = 0; q < 3; ++q) s->h_poses0[6 * (size_t)i + 3 + q] = -pos[3 * (size_t)i + q];
  MM_CUDA(s->blk_a.alloc((size_t)std::max(n_off, 1))); MM_CUDA(s->blk_b.alloc((size_t)std::max(n_off, 1))); MM_CUDA(s->S.alloc(36 * (size_t)s->nblk)); MM_CUDA(s->fail.alloc(1));
  if (n_off) { MM_CUDA(cudaMemcpy(s->blk_a.p, blk_a, sizeof(int) * (size_t)n_off, cudaMemcpyHostToDevice)); MM_CUDA(cudaMemcpy(s->blk_b.p, blk_b, sizeof(int) * (size_t)n_off, cudaMemcpyHostToDevice)); }
  MM_CUDA(cudaMemcpy(s->S.p, S, sizeof(double) * 36 * (size_t)s->nblk, cudaMemcpyHostToDevice));
  MM_CUDA(cudaMemset(s->fail.p, 0, sizeof(int)));
  if (ncb) {
    MM_CUDA(s->Bm.alloc((size_t)54 * n_img * ncb)); MM_CUDA(s->intr_acc.alloc((size_t)81 * ncb * ncb + 27 * (size_t)ncb));
    MM_CUDA(cudaMemcpy(s->Bm.p, Bm, sizeof(double) * 54 * (size_t)n_img * ncb, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(s->intr_acc.p, Cm, sizeof(double) * 81 * (size_t)ncb * ncb, cudaMemcpyHostToDevice));
  }
  rc = tc_begin(s); if (rc) return rc;
  { bool fallback = false; rc = tc_finish(s, &fallback); if (rc) return rc; }
  if (!s->tc_on) { set_error("tile Cholesky not enabled"); return MM_ERR_UNSUPPORTED; }
  DevBuf<double> d_r, d_z; MM_CUDA(d_r.alloc((size_t)s->n_unk)); MM_CUDA(d_z.alloc((size_t)s->n_unk));
  MM_CUDA(cudaMemcpy(d_r.p, rhs, sizeof(double) * (size_t)s->n_unk, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  if (reps < 1) reps = 1;
  DevBuf<unsigned long long> trace; const size_t n_trace = 2 * ((size_t)s->tc.n_tasks + (size_t)s->tc.n_stasks);
  DevBuf<unsigned long long> trace_d; const size_t n_trace_d = 8 * (size_t)s->tc_plan.n_w;
  if (getenv("MM_TC_TRACE")) { MM_CUDA(trace.alloc(n_trace)); MM_CUDA(cudaMemset(trace.p, 0, sizeof(unsigned long long) * n_trace)); s->tc.trace = trace.p;
    MM_CUDA(trace_d.alloc(n_trace_d)); MM_CUDA(cudaMemset(trace_d.p, 0, sizeof(unsigned long long) * n_trace_d)); s->tc.trace_diag = trace_d.p; }
  for (int r = -1; r < reps && rc == MM_OK; ++r) { if (r == 0) cudaEventRecord(e0, nullptr); rc = launch_tc_factor(s); }
  cudaEventRecord(e1, nullptr);
  for (int r = 0; r < reps && rc == MM_OK; ++r) rc = launch_tc_apply(s, d_r.p, d_z.p, nullptr);
  cudaEventRecord(e2, nullptr);
  if (rc == MM_OK && cudaDeviceSynchronize() != cudaSuccess) { set_error("tile Cholesky kernels failed: %s", cudaGetErrorString(cudaGetLastError())); rc = MM_ERR_CUDA; }
  float f0 = 0, f1 = 0; cudaEventElapsedTime(&f0, e0, e1); cudaEventElapsedTime(&f1, e1, e2);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  if (rc) return rc;
  if (ms_factor) *ms_factor = f0 / reps; if (ms_apply) *ms_apply = f1 / reps;
  if (trace.p) {        // time stamps of the last factorisation / application: {start, end} per task, raw uint64 ns
    std::vector<unsigned long long> h(n_trace);
    MM_CUDA(cudaMemcpy(h.data(), trace.p, sizeof(unsigned long long) * n_trace, cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(getenv("MM_TC_TRACE"), "wb")) { fwrite(h.data(), sizeof(unsigned long long), n_trace, f); fclose(f); }
    std::vector<unsigned long long> hd(n_trace_d);
    MM_CUDA(cudaMemcpy(hd.data(), trace_d.p, sizeof(unsigned long long) * n_trace_d, cudaMemcpyDeviceToHost));
    const std::string pd = std::string(getenv("MM_TC_TRACE")) + ".diag";
    if (FILE* f = fopen(pd.c_str(), "wb")) { fwrite(hd.data(), sizeof(unsigned long long), n_trace_d, f); fclose(f); }
  }
  int fail = 0; MM_CUDA(cudaMemcpy(&fail, s->fail.p, sizeof(int), cudaMemcpyDeviceToHost));
  MM_CUDA(cudaMemcpy(z, d_z.p, sizeof(double) * (size_t)s->n_unk, cudaMemcpyDeviceToHost));
  if (fail) { set_error("the matrix is not positive definite"); return MM_ERR_NUMERICAL; }
  return MM_OK;
}

int mm_ba_session_download(mm_ba_session* s, double* poses, double* intr, double* pts, double* pt_err) {
  if (!s) return MM_ERR_INVALID_ARG;
  cudaStream_t st = s->stream;
  if (poses) MM_CUDA(cudaMemcpyAsync(poses, s->poses.p, sizeof(double) * 6 * (size_t)s->n_img, cudaMemcpyDeviceToHost, st));
  if (intr) MM_CUDA(cudaMemcpyAsync(intr, s->intr.p, sizeof(double) * MM_INTR_STRIDE * (size_t)s->n_cam, cudaMemcpyDeviceToHost, st));
  if (s->world > 1 && s->n_pt > 0 && (pts || pt_err)) {      // every rank holds current values of its own points only
    if (s->p_lo > 0) MM_CUDA(cudaMemsetAsync(s->pts.p, 0, sizeof(double) * 3 * (size_t)s->p_lo, st));
    if (s->p_hi < s->n_pt) MM_CUDA(cudaMemsetAsync(s->pts.p + 3 * (size_t)s->p_hi, 0, sizeof(double) * 3 * (size_t)(s->n_pt - s->p_hi), st));
    int rc = all_reduce(s, s->pts.p, 3 * (size_t)s->n_pt); if (rc) return rc;
  }
  // pts2 (candidate buffer) and Vinv are free between LM iterations: staging for the caller's point order
  if (pts && s->n_pt > 0) {
    k_scatter_rows3<<<blocks_for(s->n_pt, 256), 256, 0, st>>>(s->n_pt, s->pt_new2old.p, s->pts.p, s->pts2.p); MM_LAUNCH_CHECK();
    MM_CUDA(cudaMemcpyAsync(pts, s->pts2.p, sizeof(double) * 3 * (size_t)s->n_pt, cudaMemcpyDeviceToHost, st));
  }
  if (pt_err && s->n_pt > 0 && s->n_obs > 0) {
    double* e_int = s->Vinv.p; double* e_ext = s->Vinv.p + (size_t)s->n_pt;       // internal order | caller order
    MM_CUDA(cudaMemcpyAsync(e_ext, pt_err, sizeof(double) * (size_t)s->n_pt, cudaMemcpyHostToDevice, st));
    k_gather_rows1<<<blocks_for(s->n_pt, 256), 256, 0, st>>>(s->n_pt, s->pt_new2old.p, e_ext, e_int); MM_LAUNCH_CHECK();
    k_pose_aux<<<blocks_for(s->n_img, 128), 128, 0, st>>>(s->n_img, s->poses.p, s->pose_mask.p, s->aux.p); MM_LAUNCH_CHECK();
    k_point_errors<<<blocks_for(s->n_pt, 128), 128, 0, st>>>(s->n_pt, s->pt_start.p, s->obs_xy.p, s->obs_img.p, s->aux.p, s->pts.p, s->intr.p,
        s->img_cam.p, s->cam_model.p, e_int); MM_LAUNCH_CHECK();     // (all points are current on every rank after the gather above)
    k_scatter_rows1<<<blocks_for(s->n_pt, 256), 256, 0, st>>>(s->n_pt, s->pt_new2old.p, e_int, e_ext); MM_LAUNCH_CHECK();
    MM_CUDA(cudaMemcpyAsync(pt_err, e_ext, sizeof(double) * (size_t)s->n_pt, cudaMemcpyDeviceToHost, st));
    MM_CUDA(cudaStreamSynchronize(st));
    // the Schur data must be rebuilt if the session continues
    if (!s->finished) { int rc = launch_schur(s); if (rc) return rc; }
  }
  MM_CUDA(cudaStreamSynchronize(st));
  return MM_OK;
}

int mm_ba_session_time_kernel(mm_ba_session* s, int32_t which, int32_t reps, double* ms_out) {
  if (!s || !ms_out || reps <= 0 || s->n_obs == 0) return MM_ERR_INVALID_ARG;
  cudaStream_t st = s->stream; int rc = MM_OK;
  // warm-up
  for (int r = -1; r < reps; ++r) {
    if (r == 0) MM_CUDA(cudaEventRecord(s->ev0, st));
    switch (which) {
      case 0: k_residual_jacobian<true, false, false><<<s->grid_obs, 256, K1_SMEM, st>>>(s->no_loc(), s->obs_xy.p + s->o_lo, s->obs_img.p + s->o_lo, s->obs_pt.p + s->o_lo, s->aux.p, s->pts.p, s->intr.p,
                  s->img_cam.p, s->cam_model.p, s->pose_mask.p, s->pt_mask.p, loss_of(s->opt), s->rec.p, s->part_cost.p); count_launch(); break;
      case 1: rc = launch_schur(s, false); break;
      case 4: rc = launch_coarse_setup(s); break;
      case 5: if (!s->tc_on) return MM_ERR_UNSUPPORTED; rc = launch_tc_factor(s); break;
      case 6: if (!s->tc_on) return MM_ERR_UNSUPPORTED; rc = launch_tc_apply(s, s->dv_b.p, s->dv_z.p, nullptr); break;
      case 2: k_residual_jacobian<false><<<s->grid_obs, 256, K1_SMEM_COST, st>>>(s->no_loc(), s->obs_xy.p + s->o_lo, s->obs_img.p + s->o_lo, s->obs_pt.p + s->o_lo, s->aux.p, s->pts.p, s->intr.p,
                  s->img_cam.p, s->cam_model.p, s->pose_mask.p, s->pt_mask.p, loss_of(s->opt), nullptr, s->part_cost.p); count_launch(); break;
      case 3: {
        MM_CUDA(cudaMemsetAsync(s->pcg_ic.p, 0, sizeof(int) * 4, st));
        k_pcg_spmv<<<blocks_for((int64_t)s->n_img * 32, 128), 128, 0, st>>>(s->n_img, s->row_start.p, s->row_col.p, s->row_blk.p, s->S.p,
            s->vz.p, s->vp0.p, s->vp1.p, s->vAp.p, s->pcg_sc.p + 8, s->pcg_ic.p, 1); count_launch(); break; }
      default: return MM_ERR_INVALID_ARG;
    }
    if (rc) return rc;
  }
  MM_CUDA(cudaEventRecord(s->ev1, st)); MM_CUDA(cudaEventSynchronize(s->ev1));
  float ms = 0; MM_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
  *ms_out = ms / reps;
  if (which == 3) { MM_CUDA(cudaMemsetAsync(s->pcg_sc.p, 0, sizeof(double) * 8, st)); }
  if (which == 1) { rc = launch_coarse_setup(s); if (rc) return rc; }      // leave the session consistent
  return MM_OK;
}

int mm_ba_solve(mm_ba_problem* P, const mm_ba_options* opt, mm_ba_summary* summary) {
  if (!P || !opt) { set_error("null argument"); return MM_ERR_INVALID_ARG; }
  mm_ba_session* s = nullptr;
  int rc = session_create(P, opt, nullptr, 0, 1, nullptr, nullptr, false, &s); if (rc) return rc;
  int32_t done = 0;
  rc = mm_ba_session_iterate(s, opt->max_num_iterations + 1, &done);
  if (rc == MM_OK) rc = mm_ba_session_download(s, P->poses, P->intr, P->pts, P->pt_err);
  if (summary) *summary = s->sum;
  mm_ba_session_destroy(s);
  return rc;
}

int mm_pose_refine(double* rvec, double* tvec, int model_code, const double* params, int64_t n, const double* points2D,
                   const double* points3D, const uint8_t* inlier_mask, const mm_ba_options* opt, mm_ba_summary* summary, double* ret) {
  // bundle_adjustment.cc:139-225: one free pose; points (:187) and intrinsics (:193) constant
  if (!rvec || !tvec || !params || n < 0 || model_num_params(model_code) < 0 || !opt || (n > 0 && (!points2D || !points3D))) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  std::vector<double> pts, obs; std::vector<int32_t> oi, op; std::vector<uint8_t> pc;
  for (int64_t i = 0; i < n; ++i) if (!inlier_mask || inlier_mask[i]) {
    op.push_back((int32_t)pc.size()); oi.push_back(0); pc.push_back(1);
    pts.insert(pts.end(), points3D + 3 * i, points3D + 3 * i + 3); obs.insert(obs.end(), points2D + 2 * i, points2D + 2 * i + 2);
  }
  double poses[6] = { rvec[0], rvec[1], rvec[2], tvec[0], tvec[1], tvec[2] };
  if (!pc.empty() && !getenv("MM_POSE_REFINE_GENERAL")) {
    // latency path: the whole LM loop in one single-CTA kernel (ba_pose.cuh); one upload, one launch, one download
    int rc = ensure_device(); if (rc) return rc;
    const size_t m = pc.size();
    DevBuf<double> d_in, d_pose; DevBuf<mm_ba_summary> d_sum;       // d_in = [uv (2m) | X (3m) | intr (9)]
    MM_CUDA(d_in.alloc(5 * m + MM_INTR_STRIDE)); MM_CUDA(d_pose.alloc(6)); MM_CUDA(d_sum.alloc(1));
    double intr9[MM_INTR_STRIDE] = {0};
    memcpy(intr9, params, sizeof(double) * (size_t)model_num_params(model_code));
    MM_CUDA(cudaMemsetAsync(d_sum.p, 0, sizeof(mm_ba_summary), nullptr));             // (the kernel fills the trace only up to num_iterations)
    MM_CUDA(cudaMemcpyAsync(d_in.p, obs.data(), sizeof(double) * 2 * m, cudaMemcpyHostToDevice, nullptr));
    MM_CUDA(cudaMemcpyAsync(d_in.p + 2 * m, pts.data(), sizeof(double) * 3 * m, cudaMemcpyHostToDevice, nullptr));
    MM_CUDA(cudaMemcpyAsync(d_in.p + 5 * m, intr9, sizeof intr9, cudaMemcpyHostToDevice, nullptr));
    MM_CUDA(cudaMemcpyAsync(d_pose.p, poses, sizeof poses, cudaMemcpyHostToDevice, nullptr));
    k_pose_refine<<<1, 256, 0, nullptr>>>((int)m, reinterpret_cast<const double2*>(d_in.p), d_in.p + 2 * m, model_code, d_in.p + 5 * m, *opt, d_pose.p, d_sum.p);
    MM_LAUNCH_CHECK();
    mm_ba_summary S;
    MM_CUDA(cudaMemcpyAsync(&S, d_sum.p, sizeof S, cudaMemcpyDeviceToHost, nullptr));
    MM_CUDA(cudaMemcpyAsync(poses, d_pose.p, sizeof poses, cudaMemcpyDeviceToHost, nullptr));
    MM_CUDA(cudaStreamSynchronize(nullptr));
    S.return_value = sqrt(S.final_cost / (double)std::max<int64_t>(S.num_residuals, 1));
    S.ms_setup = S.ms_linearize = S.ms_schur = S.ms_pcg = S.ms_update = S.ms_total = 0.0;
    if (S.termination == MM_TERM_NUMERICAL_FAILURE && !isfinite(S.initial_cost)) { set_error("non-finite initial cost"); return MM_ERR_NUMERICAL; }
    for (int k = 0; k < 3; ++k) { rvec[k] = poses[k]; tvec[k] = poses[3 + k]; }
    if (ret) *ret = S.return_value; if (summary) *summary = S;
    return MM_OK;
  }
  uint8_t pose_const[4] = {0, 0, 0, 0}; int32_t img_cam[1] = {0}; double intr[MM_INTR_STRIDE] = {0};
  memcpy(intr, params, sizeof(double) * (size_t)model_num_params(model_code));
  int32_t cam_model[1] = { model_code }; uint8_t intr_const[1] = {1};
  mm_ba_problem P; memset(&P, 0, sizeof P);
  P.n_img = 1; P.n_cam = 1; P.n_pt = (int32_t)pc.size(); P.n_obs = (int64_t)pc.size();
  double dummy3[3] = {0, 0, 0}, dummy2[2] = {0, 0}; int32_t dummyi[1] = {0}; uint8_t dummyc[1] = {1};
  P.poses = poses; P.pose_const = pose_const; P.img_cam = img_cam; P.intr = intr; P.cam_model = cam_model; P.intr_const = intr_const;
  P.pts = pts.empty() ? dummy3 : pts.data(); P.pt_const = pc.empty() ? dummyc : pc.data();
  P.obs_xy = obs.empty() ? dummy2 : obs.data(); P.obs_img = oi.empty() ? dummyi : oi.data(); P.obs_pt = op.empty() ? dummyi : op.data();
  mm_ba_summary S; int rc = mm_ba_solve(&P, opt, &S);
  if (rc == MM_OK) { for (int k = 0; k < 3; ++k) { rvec[k] = poses[k]; tvec[k] = poses[3 + k]; } if (ret) *ret = S.return_value; if (summary) *summary = S; }
  return rc;
}

/* Batch of independent pose refinements in one launch ("batched across candidate pairs").  Problem b refines
 * rvecs/tvecs[b] against its own 2D-3D pairs offsets[b] .. offsets[b+1] of the concatenated arrays (already inlier-filtered), with
 * its own camera (model_codes[b], params[b][9]).  Each problem is solved exactly as mm_pose_refine solves it (same kernel body). */
int mm_pose_refine_batch(int32_t n_problems, double* rvecs, double* tvecs, const int32_t* model_codes, const double* params,
                         const int64_t* offsets, const double* points2D, const double* points3D, const mm_ba_options* opt,
                         mm_ba_summary* summaries, double* rets) {
  if (n_problems < 0 || !opt || (n_problems > 0 && (!rvecs || !tvecs || !model_codes || !params || !offsets))) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  if (n_problems == 0) return MM_OK;
  const size_t B = (size_t)n_problems;
  if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return MM_ERR_INVALID_ARG; }
  for (size_t b = 0; b < B; ++b) {
    if (offsets[b + 1] <= offsets[b]) { set_error("problem %zu has no 2D-3D pairs", b); return MM_ERR_INVALID_ARG; }
    if (model_num_params(model_codes[b]) < 0) { set_error("unknown camera model code %d", model_codes[b]); return MM_ERR_INVALID_ARG; }
  }
  const size_t m = (size_t)offsets[B];
  if (!points2D || !points3D) { set_error("null buffer"); return MM_ERR_INVALID_ARG; }
  int rc = ensure_device(); if (rc) return rc;
  DevBuf<double> d_in, d_pose, d_intr; DevBuf<int64_t> d_off; DevBuf<int> d_model; DevBuf<mm_ba_summary> d_sum;      // d_in = [uv (2m) | X (3m)]
  MM_CUDA(d_in.alloc(5 * m)); MM_CUDA(d_pose.alloc(6 * B)); MM_CUDA(d_intr.alloc(MM_INTR_STRIDE * B)); MM_CUDA(d_off.alloc(B + 1)); MM_CUDA(d_model.alloc(B)); MM_CUDA(d_sum.alloc(B));
  std::vector<double> poses(6 * B), intr(MM_INTR_STRIDE * B, 0.0);
  for (size_t b = 0; b < B; ++b) {
    for (int k = 0; k < 3; ++k) { poses[6 * b + k] = rvecs[3 * b + k]; poses[6 * b + 3 + k] = tvecs[3 * b + k]; }
    memcpy(intr.data() + MM_INTR_STRIDE * b, params + MM_INTR_STRIDE * b, sizeof(double) * (size_t)model_num_params(model_codes[b]));
  }
  MM_CUDA(cudaMemsetAsync(d_sum.p, 0, sizeof(mm_ba_summary) * B, nullptr));          // (the kernel fills the trace only up to num_iterations)
  MM_CUDA(cudaMemcpyAsync(d_in.p, points2D, sizeof(double) * 2 * m, cudaMemcpyHostToDevice, nullptr));
  MM_CUDA(cudaMemcpyAsync(d_in.p + 2 * m, points3D, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, nullptr));
  MM_CUDA(cudaMemcpyAsync(d_intr.p, intr.data(), sizeof(double) * intr.size(), cudaMemcpyHostToDevice, nullptr));
  MM_CUDA(cudaMemcpyAsync(d_pose.p, poses.data(), sizeof(double) * poses.size(), cudaMemcpyHostToDevice, nullptr));
  MM_CUDA(cudaMemcpyAsync(d_off.p, offsets, sizeof(int64_t) * (B + 1), cudaMemcpyHostToDevice, nullptr));
  MM_CUDA(cudaMemcpyAsync(d_model.p, model_codes, sizeof(int) * B, cudaMemcpyHostToDevice, nullptr));
  k_pose_refine_batch<<<(unsigned)B, 256, 0, nullptr>>>(d_off.p, reinterpret_cast<const double2*>(d_in.p), d_in.p + 2 * m, d_model.p, d_intr.p, *opt, d_pose.p, d_sum.p);
  MM_LAUNCH_CHECK();
  std::vector<mm_ba_summary> S(B);
  MM_CUDA(cudaMemcpyAsync(S.data(), d_sum.p, sizeof(mm_ba_summary) * B, cudaMemcpyDeviceToHost, nullptr));
  MM_CUDA(cudaMemcpyAsync(poses.data(), d_pose.p, sizeof(double) * poses.size(), cudaMemcpyDeviceToHost, nullptr));
  MM_CUDA(cudaStreamSynchronize(nullptr));
  for (size_t b = 0; b < B; ++b) {
    S[b].return_value = sqrt(S[b].final_cost / (double)std::max<int64_t>(S[b].num_residuals, 1));
    S[b].ms_setup = S[b].ms_linearize = S[b].ms_schur = S[b].ms_pcg = S[b].ms_update = S[b].ms_total = 0.0;
    if (S[b].termination == MM_TERM_NUMERICAL_FAILURE && !isfinite(S[b].initial_cost)) { set_error("non-finite initial cost in problem %zu", b); return MM_ERR_NUMERICAL; }
    for (int k = 0; k < 3; ++k) { rvecs[3 * b + k] = poses[6 * b + k]; tvecs[3 * b + k] = poses[6 * b + 3 + k]; }
    if (rets) rets[b] = S[b].return_value;
    if (summaries) summaries[b] = S[b];
  }
  return MM_OK;
}

}  // extern "C"
