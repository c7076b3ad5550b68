// ba_kernels.cuh — device kernels of the bundle-adjustment inner loop (fp64, HBM-bound).
//
// What they replace in the reference (mavmap/mavmap) + Ceres:
//   K1 k_residual_jacobian   Ceres autodiff of BACostFunction<M>::operator()
//                            (src/base3d/bundle_adjustment.h:131-159) + Corrector for
//                            CauchyLoss (bundle_adjustment.cc:477-478)
//   K2 k_schur_point/_cam    Ceres SchurEliminator::Eliminate behind SPARSE_SCHUR
//                            (bundle_adjustment.cc:555)
//   K3 k_pcg_*               the reduced-camera-system solve (Ceres: sparse Cholesky)
//   K4 k_backsub / k_cost    SchurEliminator::BackSubstitute + the candidate-cost Evaluate
// Layout (DESIGN.md §Data layout): observations sorted by 3-D point; one 160-byte record per
// observation  [r0 r1 | Jc row0 (w0 w1 w2 tx ty tz) | Jc row1 | Jp row0 | Jp row1]  so that the
// point pass streams records and the camera pass gathers whole 32-byte sectors.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "camera.cuh"
#include "ba_coarse.cuh"

namespace mm {

constexpr int REC = 20;     // doubles per observation record (160 B = 5 sectors)
constexpr int PINFO = 12;                // doubles per point record (96 B = 3 sectors): V^-1 (6) | scale_p (3) | g_p (3)
constexpr int AUX = 24;     // doubles per image: R(9) t(3) Jl(9) masks(2) pad = 192 B

struct LossParams { int type; double b; double c; };   // Cauchy: b = a^2, c = 1/b

template <bool WITH_COST = true>
__device__ __forceinline__ void loss_eval(const LossParams& L, double s, double& rho0, double& sqrt_rho1) {
  if (L.type == MM_LOSS_CAUCHY) {
    const double sum = 1.0 + s * L.c, inv = 1.0 / sum;
    rho0 = WITH_COST ? L.b * log(sum) : 0.0;
    sqrt_rho1 = sqrt(fmax(inv, 2.2250738585072014e-308));
  } else { rho0 = s; sqrt_rho1 = 1.0; }
}

// ---- per-image rotation data ---------------------------------------------------------
// aux[img] = [R (9) | t (3) | Jl (9) | rvec mask, tvec mask bits | pad]: the first 96 bytes are all the cost-only pass needs
__global__ void k_pose_aux(int n_img, const double* __restrict__ poses, const double* __restrict__ pose_mask, double* __restrict__ aux) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img) return;
  double R[9], Jl[9];
  const double* p = poses + 6 * (size_t)i;
  rotation_and_left_jacobian(p, R, Jl);
  double* a = aux + AUX * (size_t)i;
#pragma unroll
  for (int k = 0; k < 9; ++k) { a[k] = R[k]; a[12 + k] = Jl[k]; }
  a[9] = p[3]; a[10] = p[4]; a[11] = p[5];
  const double* m = pose_mask + 6 * (size_t)i;       // free-parameter masks ride along: [21] rvec, [22] tx + 2 ty + 4 tz
  a[21] = m[0]; a[22] = m[3] + 2.0 * m[4] + 4.0 * m[5]; a[23] = 0;
}

// ---- K1: residual + Jacobian per observation -------------------------------------------
// reads 24 B/obs (xy, img, pt) + gathered parameters, writes one 160 B record.
// cost partial per block -> cost_part[blockIdx.x] (reduced deterministically afterwards).
//
// Observations are sorted by point, so the 32 observations of a warp see only a handful of distinct images.  A per-lane
// gather of the 192-byte image record costs 12 x 32 L1 tag look-ups per warp and made the kernel L1-bound (ncu r1d: l1tex 72 %,
// DRAM 38 %).  Instead the warp finds its distinct images (__match_any_sync), fetches each record ONCE with coalesced 16-byte
// loads into a shared-memory slot, and every lane reads its slot (conflict-free 208-byte pitch).  The same shared-memory area
// then stages the output records so that every store instruction of the warp covers one contiguous 512-byte run.
constexpr int K1_SLOT = 26;                                   // doubles per image slot (208 B pitch: 16-byte aligned, bank-conflict-free)
constexpr int K1_PITCH = 22;                                  // doubles per staged record
constexpr int K1_WARP_DOUBLES = 32 * K1_SLOT;                 // >= 32 * K1_PITCH
constexpr size_t K1_SMEM = sizeof(double) * 8 * K1_WARP_DOUBLES + sizeof(int) * 8 * 32 + sizeof(double) * 32 + sizeof(unsigned long long) * 8;
constexpr size_t K1_SMEM_COST = K1_SMEM;


// K1 fetches the per-image records with 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP): the leader lane of every distinct
// image of the warp issues one copy of the record into the image's shared-memory slot, completion on the warp's mbarrier.
__device__ __forceinline__ void k1_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void k1_mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

// WITH_COST = false: the Jacobian pass after an accepted step (the cost of the new iterate is the candidate cost that the
// cost-only pass has just computed, so the logarithm of the Cauchy loss is not evaluated again).
template <bool WITH_J, bool WITH_JI = false, bool WITH_COST = true>
__global__ void __launch_bounds__(256, WITH_J ? 2 : 3) k_residual_jacobian(
    int64_t n_obs, const double2* __restrict__ obs_xy, const int* __restrict__ obs_img, const int* __restrict__ obs_pt,
    const double* __restrict__ aux, const double* __restrict__ pts, const double* __restrict__ intr,
    const int* __restrict__ img_cam, const int* __restrict__ cam_model,
    const double* __restrict__ pose_mask, const double* __restrict__ pt_mask,
    LossParams L, double* __restrict__ rec, double* __restrict__ cost_part,
    double* __restrict__ ji = nullptr, const double* __restrict__ intr_mask = nullptr) {
  extern __shared__ __align__(16) unsigned char k1_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* wbuf = reinterpret_cast<double*>(k1_smem) + (size_t)wib * K1_WARP_DOUBLES;       // image slots, then staged records
  int* wimg = reinterpret_cast<int*>(k1_smem + sizeof(double) * 8 * K1_WARP_DOUBLES) + wib * 32;
  double* red = reinterpret_cast<double*>(k1_smem + sizeof(double) * 8 * K1_WARP_DOUBLES + sizeof(int) * 8 * 32);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(red + 32);
  const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + wib);
  unsigned phase = 0;
  if (lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  constexpr int NCH = WITH_J ? 12 : 6;                        // 16-byte chunks of the image record that this pass needs
  double cost = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i0 = blockIdx.x * (int64_t)blockDim.x + wib * 32;
  // software pipeline: the indices / observation of the next round are in flight while this one is computed
  int img_n = 0, pt_n = 0; double2 xy_n = make_double2(0.0, 0.0);
  if (i0 < n_obs) { const int64_t i = min(i0 + lane, n_obs - 1); img_n = obs_img[i]; pt_n = obs_pt[i]; xy_n = obs_xy[i]; }
  for (; i0 < n_obs; i0 += stride) {
    const bool valid = i0 + lane < n_obs;                     // lanes past the end recompute the last observation and are masked out
    const int img = img_n, pt = pt_n; const double2 xy = xy_n;
    if (i0 + stride < n_obs) { const int64_t i = min(i0 + stride + lane, n_obs - 1); img_n = obs_img[i]; pt_n = obs_pt[i]; xy_n = obs_xy[i]; }
    const double X0 = pts[3 * (size_t)pt], X1 = pts[3 * (size_t)pt + 1], X2 = pts[3 * (size_t)pt + 2];
    // distinct images of this warp -> shared-memory slots
    const unsigned grp = __match_any_sync(0xffffffffu, img);
    const int leader = __ffs(grp) - 1;
    const unsigned lead_mask = __ballot_sync(0xffffffffu, lane == leader);
    const int slot = __popc(lead_mask & ((1u << leader) - 1u));
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((unsigned)(__popc(lead_mask) * NCH * 16)) : "memory");
    if (lane == leader) k1_bulk_g2s(wbuf + slot * K1_SLOT, aux + AUX * (size_t)img, NCH * 16, bar);
    k1_mbar_wait(bar, phase); phase ^= 1;
    double a[AUX];
    {
      const double2* s2 = reinterpret_cast<const double2*>(wbuf + slot * K1_SLOT);
#pragma unroll
      for (int k = 0; k < NCH; ++k) { const double2 t = s2[k]; a[2 * k] = t.x; a[2 * k + 1] = t.y; }
    }
    if (!WITH_J) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // (the Jacobian pass fences after it has staged its records)
    __syncwarp();                                             // slots are free again: the same area stages the records below
    (void)wimg;
    const double* R = a;
    const double Y0 = R[0] * X0 + R[1] * X1 + R[2] * X2;
    const double Y1 = R[3] * X0 + R[4] * X1 + R[5] * X2;
    const double Y2 = R[6] * X0 + R[7] * X1 + R[8] * X2;
    const double xc = Y0 + a[9], yc = Y1 + a[10], zc = Y2 + a[11];
    const int cam = img_cam[img];
    const int model = cam_model[cam];
    double u, v, dX[2][3], dP[2][9];
    world2image<WITH_J>(model, intr + MM_INTR_STRIDE * (size_t)cam, xc, yc, zc, u, v, dX, WITH_JI ? dP : nullptr);
    const double r0 = u - xy.x, r1 = v - xy.y;
    double rho0, sr;
    loss_eval<WITH_COST>(L, r0 * r0 + r1 * r1, rho0, sr);
    if (WITH_COST && valid) cost += 0.5 * rho0;
    if (WITH_JI) {        // d r / d intrinsics (2 x 9), robustified and masked; 144 B per observation
      if (valid) {
        double2* j2 = reinterpret_cast<double2*>(ji + 18 * (size_t)(i0 + lane));
        double t[18];
#pragma unroll
        for (int k = 0; k < 9; ++k) { const double m = intr_mask[9 * (size_t)cam + k] * sr; t[k] = dP[0][k] * m; t[9 + k] = dP[1][k] * m; }
#pragma unroll
        for (int k = 0; k < 9; ++k) j2[k] = make_double2(t[2 * k], t[2 * k + 1]);
      }
    }
    if (WITH_J) {
      const double mp = pt_mask[pt] * sr;
      const int mbits = (int)a[22];
      const double mw = a[21] * sr, mx = (mbits & 1) ? sr : 0.0, my = (mbits & 2) ? sr : 0.0, mz = (mbits & 4) ? sr : 0.0;
      double2* o2 = reinterpret_cast<double2*>(wbuf + (size_t)lane * K1_PITCH);
      o2[0] = make_double2(sr * r0, sr * r1);
      // d(Xc)/d(w) = -[Y]x Jl : column k = Jl[:,k] x Y
      double M[3][3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double c0 = a[12 + k], c1 = a[15 + k], c2 = a[18 + k];
        M[0][k] = c1 * Y2 - c2 * Y1; M[1][k] = c2 * Y0 - c0 * Y2; M[2][k] = c0 * Y1 - c1 * Y0;
      }
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const double d0 = dX[row][0], d1 = dX[row][1], d2 = dX[row][2];
        const double jw0 = (d0 * M[0][0] + d1 * M[1][0] + d2 * M[2][0]) * mw;
        const double jw1 = (d0 * M[0][1] + d1 * M[1][1] + d2 * M[2][1]) * mw;
        const double jw2 = (d0 * M[0][2] + d1 * M[1][2] + d2 * M[2][2]) * mw;
        o2[1 + 3 * row] = make_double2(jw0, jw1);
        o2[2 + 3 * row] = make_double2(jw2, d0 * mx);
        o2[3 + 3 * row] = make_double2(d1 * my, d2 * mz);
      }
      const double p00 = (dX[0][0] * R[0] + dX[0][1] * R[3] + dX[0][2] * R[6]) * mp;
      const double p01 = (dX[0][0] * R[1] + dX[0][1] * R[4] + dX[0][2] * R[7]) * mp;
      const double p02 = (dX[0][0] * R[2] + dX[0][1] * R[5] + dX[0][2] * R[8]) * mp;
      const double p10 = (dX[1][0] * R[0] + dX[1][1] * R[3] + dX[1][2] * R[6]) * mp;
      const double p11 = (dX[1][0] * R[1] + dX[1][1] * R[4] + dX[1][2] * R[7]) * mp;
      const double p12 = (dX[1][0] * R[2] + dX[1][1] * R[5] + dX[1][2] * R[8]) * mp;
      o2[7] = make_double2(p00, p01);
      o2[8] = make_double2(p02, p10);
      o2[9] = make_double2(p11, p12);
      __syncwarp();
      const int nvalid = (int)min((int64_t)32, n_obs - i0);
      double2* g2 = reinterpret_cast<double2*>(rec + REC * (size_t)i0);
#pragma unroll
      for (int k = 0; k < 10; ++k) {
        const int c = k * 32 + lane, src = c / 10, part = c - 10 * src;
        if (src < nvalid) g2[c] = *reinterpret_cast<const double2*>(wbuf + (size_t)src * K1_PITCH + 2 * part);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic accesses to the slots before the next round's bulk copies land there
      __syncwarp();
    }
  }
  if (WITH_COST) {
    cost = block_sum(cost, red);
    if (threadIdx.x == 0) cost_part[blockIdx.x] = cost;
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// deterministic final reduction of per-block partials: out[0] = sum(part[0..n))
__global__ void k_reduce_sum(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

// ---- Jacobi column scaling (estimated once, trust_region_minimizer.cc EstimateScale) ----
__global__ void k_colnorm_point(int n_pt, const int* __restrict__ pt_start, const double* __restrict__ rec, double* __restrict__ scale_p) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pt) return;
  double s0 = 0, s1 = 0, s2 = 0;
  for (int o = pt_start[p]; o < pt_start[p + 1]; ++o) {
    const double* r = rec + REC * (size_t)o + 14;
    s0 += r[0] * r[0] + r[3] * r[3]; s1 += r[1] * r[1] + r[4] * r[4]; s2 += r[2] * r[2] + r[5] * r[5];
  }
  scale_p[3 * (size_t)p] = 1.0 / (1.0 + sqrt(s0));
  scale_p[3 * (size_t)p + 1] = 1.0 / (1.0 + sqrt(s1));
  scale_p[3 * (size_t)p + 2] = 1.0 / (1.0 + sqrt(s2));
}

// one warp per image over [cam_lo[w], cam_hi[w]) of the camera-sorted permutation (the whole image, or this rank's part
// of it when the points are sharded): raw column sums of squares; k_scale_finish turns them into 1/(1+norm)
__global__ void k_colnorm_cam(int n_img, const int* __restrict__ cam_lo, const int* __restrict__ cam_hi, const int* __restrict__ cam_perm,
                              const double* __restrict__ rec, double* __restrict__ sums) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_img) return;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int q = cam_lo[w] + lane; q < cam_hi[w]; q += 32) {
    const double* r = rec + REC * (size_t)cam_perm[q] + 2;
#pragma unroll
    for (int k = 0; k < 6; ++k) s[k] += r[k] * r[k] + r[6 + k] * r[6 + k];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) s[k] = warp_sum(s[k]);
  if (lane < 6) {
    double v = s[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) if (lane == k) v = s[k];
    sums[6 * (size_t)w + lane] = v;
  }
}
__global__ void k_scale_finish(int n, const double* __restrict__ sums, double* __restrict__ scale_c, const double* __restrict__ pr_J) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = sums[i];
  const int img = i / 6, a = i - 6 * img;
  if (pr_J && a < 3) { const double j = pr_J[3 * (size_t)img + a]; v += j * j; }       // rotation-constraint rows
  scale_c[i] = 1.0 / (1.0 + sqrt(v));
}

// ---- rotation constraints (constrain_rotation, bundle_adjustment.cc:390-446; cost functor .cc:57-111) ----------
// r = weight * sqrt(sum_i (R_a(i) - R0_b(i))^2) over the reference's nine index pairs (||R' - R0||_F with the index quirk
// of .cc:103), no loss function; d r / d rvec in closed form with d R / d w_k = [Jl e_k]x R.  One residual per image.
__device__ __forceinline__ double rot_prior_eval(const double* w, const double* w0, double weight, double* J) {
  const int PA[9] = { 0, 1, 2, 3, 4, 5, 6, 2, 8 }, PB[9] = { 0, 3, 6, 1, 4, 7, 2, 5, 8 };
  double R[9], Jl[9], R0[9], J0[9];
  rotation_and_left_jacobian(w, R, Jl); rotation_and_left_jacobian(w0, R0, J0);
  double q = 0.0, d[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { d[i] = R[PA[i]] - R0[PB[i]]; q += d[i] * d[i]; }
  const double n = sqrt(q);
  if (J) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double v0 = Jl[k], v1 = Jl[3 + k], v2 = Jl[6 + k];
      double dR[9];
#pragma unroll
      for (int c = 0; c < 3; ++c) { dR[c] = v1 * R[6 + c] - v2 * R[3 + c]; dR[3 + c] = v2 * R[c] - v0 * R[6 + c]; dR[6 + c] = v0 * R[3 + c] - v1 * R[c]; }
      double sdr = 0.0;
#pragma unroll
      for (int i = 0; i < 9; ++i) sdr += d[i] * dR[PA[i]];
      J[k] = n > 0.0 ? weight * sdr / n : 0.0;
    }
  }
  return weight * n;
}
// one block: residual, masked Jacobian (unscaled) and the cost 1/2 sum r^2 of all rotation constraints
__global__ void k_rot_prior(int n_img, const double* __restrict__ poses, const double* __restrict__ pose_mask, const double* __restrict__ rot0,
                            const double* __restrict__ weight, double* __restrict__ r_out, double* __restrict__ J_out, double* __restrict__ cost_out) {
  __shared__ double red[32];
  double c = 0.0;
  for (int i = threadIdx.x; i < n_img; i += blockDim.x) {
    double r = 0.0, J[3] = {0, 0, 0};
    if (weight[i] != 0.0) r = rot_prior_eval(poses + 6 * (size_t)i, rot0 + 3 * (size_t)i, weight[i], J_out ? J : nullptr);
    c += 0.5 * r * r;
    if (r_out) { r_out[i] = r; const double m = pose_mask[6 * (size_t)i]; J_out[3 * (size_t)i] = J[0] * m; J_out[3 * (size_t)i + 1] = J[1] * m; J_out[3 * (size_t)i + 2] = J[2] * m; }
  }
  c = block_sum(c, red);
  if (threadIdx.x == 0) cost_out[0] = c;
}

// ---- K2a: per-point blocks ------------------------------------------------------------------
// thread per point (streams its records): V' = s_p (sum Jp'Jp) s_p + D_p^2, inverse, g_p'.
struct LMDiag { double radius, min_diag, max_diag; };

__device__ __forceinline__ bool sym3_inverse(const double* V /*xx xy xz yy yz zz*/, double* I) {
  const double a = V[0], b = V[1], c = V[2], d = V[3], e = V[4], f = V[5];
  const double A = d * f - e * e, B = -(b * f - c * e), C = b * e - c * d;
  const double det = a * A + b * B + c * C;
  if (!(det > 0.0) || !isfinite(det)) return false;
  const double id = 1.0 / det;
  I[0] = A * id; I[1] = B * id; I[2] = C * id;
  I[3] = (a * f - c * c) * id; I[4] = -(a * e - b * c) * id; I[5] = (a * d - b * b) * id;
  return true;
}

__device__ __forceinline__ void load_scaled(const double* __restrict__ rec, int64_t o, const double* __restrict__ sc /*6*/,
                                            const double* sp /*3*/, double (&Jc)[2][6], double (&Jp)[2][3]) {
  const double2* r2 = reinterpret_cast<const double2*>(rec + REC * (size_t)o);
  double t[18];
#pragma unroll
  for (int k = 0; k < 9; ++k) { const double2 v = r2[1 + k]; t[2 * k] = v.x; t[2 * k + 1] = v.y; }
#pragma unroll
  for (int k = 0; k < 6; ++k) { const double s = sc[k]; Jc[0][k] = t[k] * s; Jc[1][k] = t[6 + k] * s; }
#pragma unroll
  for (int k = 0; k < 3; ++k) { Jp[0][k] = t[12 + k] * sp[k]; Jp[1][k] = t[15 + k] * sp[k]; }
}

__global__ void __launch_bounds__(128) k_schur_point(
    int n_pt, const int* __restrict__ pt_start, const double* __restrict__ rec, const double* __restrict__ scale_p, LMDiag lm,
    double* __restrict__ Vinv, double* __restrict__ gp_out, double* __restrict__ dp_out, double* __restrict__ gmax, int* __restrict__ fail,
    double* __restrict__ pinfo) {
  __shared__ double red[32];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double gm = 0.0;
  if (p < n_pt) {
    const int o0 = pt_start[p], o1 = pt_start[p + 1];
    const double sp[3] = { scale_p[3 * (size_t)p], scale_p[3 * (size_t)p + 1], scale_p[3 * (size_t)p + 2] };
    double V[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    for (int o = o0; o < o1; ++o) {
      const double2* r2 = reinterpret_cast<const double2*>(rec + REC * (size_t)o);
      const double2 rr = r2[0], a = r2[7], b = r2[8], c = r2[9];
      const double j00 = a.x * sp[0], j01 = a.y * sp[1], j02 = b.x * sp[2];
      const double j10 = b.y * sp[0], j11 = c.x * sp[1], j12 = c.y * sp[2];
      V[0] += j00 * j00 + j10 * j10; V[1] += j00 * j01 + j10 * j11; V[2] += j00 * j02 + j10 * j12;
      V[3] += j01 * j01 + j11 * j11; V[4] += j01 * j02 + j11 * j12; V[5] += j02 * j02 + j12 * j12;
      g[0] += j00 * rr.x + j10 * rr.y; g[1] += j01 * rr.x + j11 * rr.y; g[2] += j02 * rr.x + j12 * rr.y;
    }
    const double d0 = fmin(fmax(V[0], lm.min_diag), lm.max_diag) / lm.radius;
    const double d1 = fmin(fmax(V[3], lm.min_diag), lm.max_diag) / lm.radius;
    const double d2 = fmin(fmax(V[5], lm.min_diag), lm.max_diag) / lm.radius;
    V[0] += d0; V[3] += d1; V[5] += d2;
    double I[6];
    if (!sym3_inverse(V, I)) { *fail = 1; I[0] = I[3] = I[5] = 0; I[1] = I[2] = I[4] = 0; }
#pragma unroll
    for (int k = 0; k < 6; ++k) Vinv[6 * (size_t)p + k] = I[k];
    {     // the same data as one 96-byte record per point for the block and camera passes (6 whole 16-byte chunks)
      double2* pi = reinterpret_cast<double2*>(pinfo + PINFO * (size_t)p);
      pi[0] = make_double2(I[0], I[1]); pi[1] = make_double2(I[2], I[3]); pi[2] = make_double2(I[4], I[5]);
      pi[3] = make_double2(sp[0], sp[1]); pi[4] = make_double2(sp[2], g[0]); pi[5] = make_double2(g[1], g[2]);
    }
    gp_out[3 * (size_t)p] = g[0]; gp_out[3 * (size_t)p + 1] = g[1]; gp_out[3 * (size_t)p + 2] = g[2];
    dp_out[3 * (size_t)p] = d0; dp_out[3 * (size_t)p + 1] = d1; dp_out[3 * (size_t)p + 2] = d2;
    gm = fmax(fmax(fabs(g[0] / sp[0]), fabs(g[1] / sp[1])), fabs(g[2] / sp[2]));
  }
  gm = block_max(gm, red);
  if (threadIdx.x == 0 && gm > 0.0) atomic_max_nonneg(gmax, gm);
}

// ---- K2a': the blocks of the reduced camera system, one warp per stored block, no atomics ----------------
// Block (a, b), a <= b, is  - sum over the points p seen by both images of  Y_a W_b'  with  W = Jc' Jp (scaled),
// Y_a = W_a V_p^-1.  The (observation in a, observation in b, point) triples of every block were listed at setup (sorted
// by block, blocks in row-major order so that a row's records stay in L2).  Lanes take pairs round-robin with all of a
// pair's loads in flight at once (two 144-byte Jacobian records, one 80-byte point record), accumulate the 6 x 6 product
// in registers, and one shuffle reduction per block ends it: deterministic, no memset, no atomics.
// (Tried and measured slower on B200, cfg4: staging the records through shared memory with cp.async so that consecutive
// lanes read consecutive 16-byte chunks — single-buffered 2.1 ms, double-buffered 2.3 ms against 1.9 ms for this form:
// the gather is L1-tag bound, the staged forms are latency bound at the occupancy their buffers allow.)
// Diagonal blocks only receive the (rare) pairs of one point observed twice by the same image; K2b adds U - sum Y W'.

__global__ void __launch_bounds__(128, 3) k_schur_blocks(
    int n_img, int64_t nblk, const int* __restrict__ blk_a, const int* __restrict__ blk_b,
    const int* __restrict__ bp_start, const int* __restrict__ bp_end, const int* __restrict__ sp_lo, const int* __restrict__ sp_hi,
    const int* __restrict__ sp_pt, const double* __restrict__ rec, const double* __restrict__ scale_c,
    const double* __restrict__ pinfo, double* __restrict__ S) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nblk) return;
  const bool diag = b < n_img;
  const int ia = diag ? (int)b : blk_a[b - n_img], ib = diag ? (int)b : blk_b[b - n_img];
  const int q0 = bp_start[b], q1 = bp_end[b];
  double acc[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) acc[k] = 0.0;
  if (q1 > q0) {
    double sa[6], sb[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { sa[k] = scale_c[6 * (size_t)ia + k]; sb[k] = scale_c[6 * (size_t)ib + k]; }
    for (int q = q0 + lane; q < q1; q += 32) {
      const int oi = sp_lo[q], oj = sp_hi[q], p = sp_pt[q];
      double I[6], sp[3];
      {
        const double2* p2 = reinterpret_cast<const double2*>(pinfo + PINFO * (size_t)p);
        const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2], v3 = p2[3], v4 = p2[4];
        I[0] = v0.x; I[1] = v0.y; I[2] = v1.x; I[3] = v1.y; I[4] = v2.x; I[5] = v2.y; sp[0] = v3.x; sp[1] = v3.y; sp[2] = v4.x;
      }
      double Jc[2][6], Jp[2][3], Kc[2][6], Kp[2][3];
      load_scaled(rec, oi, sa, sp, Jc, Jp);
      load_scaled(rec, oj, sb, sp, Kc, Kp);
      double T[6][2];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        const double w0 = Jc[0][a] * Jp[0][0] + Jc[1][a] * Jp[1][0];
        const double w1 = Jc[0][a] * Jp[0][1] + Jc[1][a] * Jp[1][1];
        const double w2 = Jc[0][a] * Jp[0][2] + Jc[1][a] * Jp[1][2];
        const double y0 = w0 * I[0] + w1 * I[1] + w2 * I[2];
        const double y1 = w0 * I[1] + w1 * I[3] + w2 * I[4];
        const double y2 = w0 * I[2] + w1 * I[4] + w2 * I[5];
        T[a][0] = y0 * Kp[0][0] + y1 * Kp[0][1] + y2 * Kp[0][2];
        T[a][1] = y0 * Kp[1][0] + y1 * Kp[1][1] + y2 * Kp[1][2];
      }
      if (!diag) {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int c = 0; c < 6; ++c) acc[6 * a + c] -= T[a][0] * Kc[0][c] + T[a][1] * Kc[1][c];
      } else {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int c = 0; c < 6; ++c) { const double v = T[a][0] * Kc[0][c] + T[a][1] * Kc[1][c]; acc[6 * a + c] -= v; acc[6 * c + a] -= v; }
      }
    }
#pragma unroll
    for (int k = 0; k < 36; ++k) acc[k] = warp_sum(acc[k]);
  }
  double* dst = S + 36 * (size_t)b;
  double mine = acc[0];
#pragma unroll
  for (int k = 1; k < 32; ++k) if (lane == k) mine = acc[k];
  dst[lane] = mine;
  if (lane < 4) {
    double m2 = acc[32];
#pragma unroll
    for (int k = 1; k < 4; ++k) if (lane == k) m2 = acc[32 + k];
    dst[32 + lane] = m2;
  }
}

// ---- K2b: per-image diagonal block, reduced right-hand side, gradient, LM diagonal ----------
// one warp per image over its observations (camera-sorted permutation, 160 B record gathers).
__global__ void __launch_bounds__(128, 3) k_schur_cam(
    int n_img, const int* __restrict__ cam_lo, const int* __restrict__ cam_hi, const int* __restrict__ cam_perm, const int* __restrict__ obs_pt,
    const double* __restrict__ rec, const double* __restrict__ scale_c, const double* __restrict__ pinfo,
    double* __restrict__ S, double* __restrict__ rhs, double* __restrict__ gc_out, double* __restrict__ ud_out) {
  // one CTA (4 warps) per image: enough warps in flight for a 500-image problem too (one warp per image left 3 warps/SM)
  __shared__ double part[4][40];
  const int w = blockIdx.x, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double sc[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) sc[k] = scale_c[6 * (size_t)w + k];
  double acc[39];                       // [0,21) Q upper triangle | [21,27) h | [27,33) gc | [33,39) ud
#pragma unroll
  for (int k = 0; k < 39; ++k) acc[k] = 0.0;
  for (int q = cam_lo[w] + threadIdx.x; q < cam_hi[w]; q += 128) {
    const int o = cam_perm[q];
    const int p = obs_pt[o];
    double I[6], sp[3], g3[3];
    {
      const double2* p2 = reinterpret_cast<const double2*>(pinfo + PINFO * (size_t)p);
      const double2 v0 = p2[0], v1 = p2[1], v2 = p2[2], v3 = p2[3], v4 = p2[4], v5 = p2[5];
      I[0] = v0.x; I[1] = v0.y; I[2] = v1.x; I[3] = v1.y; I[4] = v2.x; I[5] = v2.y; sp[0] = v3.x; sp[1] = v3.y; sp[2] = v4.x;
      g3[0] = v4.y; g3[1] = v5.x; g3[2] = v5.y;
    }
    double Jc[2][6], Jp[2][3];
    load_scaled(rec, o, sc, sp, Jc, Jp);
    const double2 rr = *reinterpret_cast<const double2*>(rec + REC * (size_t)o);
    double W[6][3], Y[6][3];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      W[a][0] = Jc[0][a] * Jp[0][0] + Jc[1][a] * Jp[1][0];
      W[a][1] = Jc[0][a] * Jp[0][1] + Jc[1][a] * Jp[1][1];
      W[a][2] = Jc[0][a] * Jp[0][2] + Jc[1][a] * Jp[1][2];
      Y[a][0] = W[a][0] * I[0] + W[a][1] * I[1] + W[a][2] * I[2];
      Y[a][1] = W[a][0] * I[1] + W[a][1] * I[3] + W[a][2] * I[4];
      Y[a][2] = W[a][0] * I[2] + W[a][1] * I[4] + W[a][2] * I[5];
      const double ga = Jc[0][a] * rr.x + Jc[1][a] * rr.y;
      acc[27 + a] += ga;
      acc[21 + a] += ga - (Y[a][0] * g3[0] + Y[a][1] * g3[1] + Y[a][2] * g3[2]);
      acc[33 + a] += Jc[0][a] * Jc[0][a] + Jc[1][a] * Jc[1][a];
    }
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int c = a; c < 6; ++c, ++k)
        acc[k] += Jc[0][a] * Jc[0][c] + Jc[1][a] * Jc[1][c] - (Y[a][0] * W[c][0] + Y[a][1] * W[c][1] + Y[a][2] * W[c][2]);
  }
#pragma unroll
  for (int k = 0; k < 39; ++k) { const double v = warp_sum(acc[k]); if (lane == 0) part[wib][k] = v; }
  __syncthreads();
  if (threadIdx.x < 39) part[0][threadIdx.x] = part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    double* dst = S + 36 * (size_t)w;      // diagonal block id == image index
    const double* t = part[0];
    int k = 0;
    for (int a = 0; a < 6; ++a) {
      ud_out[6 * (size_t)w + a] = t[33 + a];
      gc_out[6 * (size_t)w + a] = t[27 + a];
      rhs[6 * (size_t)w + a] = t[21 + a];
      for (int c = a; c < 6; ++c, ++k) {
        if (c == a) dst[6 * a + a] += t[k];
        else { dst[6 * a + c] += t[k]; dst[6 * c + a] += t[k]; }
      }
    }
  }
}

// after the camera pass (and, when the points are sharded across GPUs, after the all-reduce of S | rhs | gc | ud | scal):
// LM diagonal of every pose parameter onto the diagonal blocks, gradient max-norm, and the reduced scalars.
// scal: [0] cost  [1] |x|^2  [2] failure count  [3 .. 3 + world) per-rank max |g_point|
__global__ void k_cam_finish(int n6, const double* __restrict__ ud, double* __restrict__ gc, double* __restrict__ rhs, const double* __restrict__ scale_c, LMDiag lm,
                             double* __restrict__ S, double* __restrict__ dc_out, const double* __restrict__ scal, int world,
                             double* __restrict__ red, int* __restrict__ fail,
                             const double* __restrict__ pr_r, const double* __restrict__ pr_J, const double* __restrict__ prior_cost) {
  __shared__ double sred[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double gm = 0.0;
  if (i < n6) {
    const int img = i / 6, a = i - 6 * img;
    double udv = ud[i], gcv = gc[i];
    if (pr_J && a < 3) {          // the image's rotation-constraint row: J'J into the rvec block, J'r into gradient and rhs
      const double ja = pr_J[3 * (size_t)img + a] * scale_c[i], r = pr_r[img];
      udv += ja * ja; gcv += ja * r; gc[i] = gcv; rhs[i] += ja * r;
#pragma unroll
      for (int c = 0; c < 3; ++c) S[36 * (size_t)img + 6 * a + c] += ja * pr_J[3 * (size_t)img + c] * scale_c[6 * (size_t)img + c];
    }
    const double d = fmin(fmax(udv, lm.min_diag), lm.max_diag) / lm.radius;
    dc_out[i] = d;
    S[36 * (size_t)img + 7 * a] += d;
    gm = fabs(gcv / scale_c[i]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int r = 0; r < world; ++r) gm = fmax(gm, scal[3 + r]);
    red[0] = scal[0] + (prior_cost ? prior_cost[0] : 0.0); red[5] = scal[1];
    if (scal[2] > 0.0) *fail = 1;
  }
  gm = block_max(gm, sred);
  if (threadIdx.x == 0 && gm > 0.0) atomic_max_nonneg(red + 2, gm);
}

// ---- block-Jacobi preconditioner: inverse of each 6x6 diagonal block --------------------
__global__ void k_precond(int n_img, const double* __restrict__ S, double* __restrict__ Minv, int* __restrict__ fail) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img) return;
  double L[6][6];
  const double* A = S + 36 * (size_t)i;
  bool ok = true;
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) L[r][c] = A[6 * r + c];
  // Cholesky
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = L[j][j];
#pragma unroll
    for (int k = 0; k < 6; ++k) if (k < j) d -= L[j][k] * L[j][k];
    if (!(d > 0.0)) { ok = false; d = 1.0; }
    const double ljj = sqrt(d);
    L[j][j] = ljj;
#pragma unroll
    for (int r = 0; r < 6; ++r) if (r > j) {
      double s = L[r][j];
#pragma unroll
      for (int k = 0; k < 6; ++k) if (k < j) s -= L[r][k] * L[j][k];
      L[r][j] = s / ljj;
    }
  }
  // invert via solving L L' X = I column by column
  double* out = Minv + 36 * (size_t)i;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    double y[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) if (k < r) s -= L[r][k] * y[k];
      y[r] = s / L[r][r];
    }
#pragma unroll
    for (int r = 5; r >= 0; --r) {
      double s = y[r];
#pragma unroll
      for (int k = 0; k < 6; ++k) if (k > r) s -= L[k][r] * y[k];
      y[r] = s / L[r][r];
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) out[6 * r + c] = y[r];
  }
  if (!ok) *fail = 1;
}

// ---- K3: PCG on the reduced camera system (block-CSR with transposed references) -----------

// stand-alone block-row SpMV (one warp per block-row; lanes split the row's blocks): p_new = z + beta p_old, Ap = S p_new.
// Only used to time one SpMV in isolation (mm_ba_session_time_kernel); the solves run in the persistent kernels below.
__global__ void __launch_bounds__(128) k_pcg_spmv(
    int n_img, const int* __restrict__ row_start, const int* __restrict__ row_col, const int* __restrict__ row_blk,
    const double* __restrict__ S, const double* __restrict__ z, const double* __restrict__ p_old, double* __restrict__ p_new,
    double* __restrict__ Ap, double* __restrict__ sc, const int* __restrict__ ic, int first) {
  if (ic[0]) return;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_img) return;
  const double beta = first ? 0.0 : sc[2] / sc[0];
  double y[6] = {0, 0, 0, 0, 0, 0};
  for (int e = row_start[w] + lane; e < row_start[w + 1]; e += 32) {
    const int col = row_col[e]; const int bid = row_blk[e];
    const bool tr = bid < 0;
    const double* B = S + 36 * (size_t)(tr ? -bid - 1 : bid);
    double pv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) pv[k] = z[6 * (size_t)col + k] + beta * p_old[6 * (size_t)col + k];
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
#pragma unroll
  for (int k = 0; k < 6; ++k) y[k] = warp_sum(y[k]);
  if (lane == 0) {
    double dot = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double pn = z[6 * (size_t)w + k] + beta * p_old[6 * (size_t)w + k];
      p_new[6 * (size_t)w + k] = pn; Ap[6 * (size_t)w + k] = y[k];
      dot += pn * y[k];
    }
    atomicAdd(sc + 1, dot);
  }
}

// ---- K3 (persistent form): the whole PCG solve in ONE cooperative launch -------------------------
// One warp per block-row of S, grid sized to the rows (<= co-resident CTAs); two grid barriers per
// iteration (after the SpMV reduction and after the residual update).  S, Minv and the CSR arrays are
// read-only inside the kernel and stay L1/L2 resident across iterations; vectors written by other CTAs
// are read with ld.global.cg after the barrier.  A single-CTA launch uses __syncthreads instead
// (local-BA sized systems).  scalars sc[]: [0,1] pAp slots, [2,3] r.z slots, [4,5] r.r slots, [6] b.b.
// Grid barrier of the persistent solvers: one arrival counter in global memory that only grows (target = arrivals so far
// + gridDim.x), release/acquire through __threadfence.  The kernels are launched cooperatively, so all CTAs are resident.
// About half the latency of cooperative_groups' grid.sync() at 60-300 CTAs, and the solvers take three per iteration.
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int seen;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory"); } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

struct PcgArgs {
  int n_img; const int* row_start; const int* row_col; const int* row_blk; const double* S; const double* Minv; const double* b;
  double *x, *r, *z, *p0, *p1, *Ap; double* sc; int* ic; double tol2; int max_iter;
  unsigned long long* dbg;   // optional: %globaltimer stamps of CTA 0 for the first 32 iterations (6 per iteration)
  // dense border of the reduced system when the (single, shared) camera's intrinsics are refined: unknowns [poses | 9 intrinsics]
  int n_intr; const double* Bm; const double* Cm; const double* Cinv; const double* bi; double *xi, *zi, *pi0, *pi1, *bt;
  // coarse level of the two-level preconditioner (ba_coarse.cuh): cm = 7 * aggregates (0 = block-Jacobi only)
  int cm; const int* agg; const double* Pc; const double* Ainv; double* rc /*[2][cm]*/; double* qc; double* yc;
};

// P_i' v added into a coarse vector: lane r < 6 holds row r of the image's 6 x 7 prolongation block (Pl) and v[r]
__device__ __forceinline__ void coarse_restrict_add(const double (&Pl)[CM], double v, int lane, bool valid, double* dst) {
  double t[CM];
#pragma unroll
  for (int k = 0; k < CM; ++k) {
    t[k] = Pl[k] * v;
    t[k] += __shfl_xor_sync(0xffffffffu, t[k], 1);
    t[k] += __shfl_xor_sync(0xffffffffu, t[k], 2);
    t[k] += __shfl_xor_sync(0xffffffffu, t[k], 4);
  }
  double mine = t[0];
#pragma unroll
  for (int k = 1; k < CM; ++k) if (lane == k) mine = t[k];
  if (valid && lane < CM && mine != 0.0) atomicAdd(dst + lane, mine);
}
// yc = Ainv (rc_old - alpha qc), rows spread over all warps of the grid; rc_new = rc_old - alpha qc (residual recursion)
__device__ __forceinline__ void coarse_solve(const PcgArgs& A, double alpha, const double* rc_old, double* rc_new, int gw, int nwt, int lane) {
  const int cm = A.cm;
  for (int j = gw; j < cm; j += nwt) {
    const double* row = A.Ainv + (size_t)j * cm;
    double s = 0.0;
    for (int k = lane; k < cm; k += 32) s += row[k] * (__ldcg(rc_old + k) - alpha * __ldcg(A.qc + k));
    s = warp_sum(s);
    if (lane == 0) { __stcg(A.yc + j, s); if (rc_new) __stcg(rc_new + j, __ldcg(rc_old + j) - alpha * __ldcg(A.qc + j)); }
  }
}
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PCG_STAMP(k) do { if (A.dbg && blockIdx.x == 0 && threadIdx.x == 0 && it < 32) A.dbg[6 * it + (k)] = gtimer(); } while (0)

__device__ __forceinline__ double block_sum_to_thread0(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(256) k_pcg_persistent(PcgArgs A) {
  namespace cg = cooperative_groups;
  unsigned int* bar_ctr = reinterpret_cast<unsigned int*>(A.ic + 3); unsigned int bar_target = 0;
  const bool single = gridDim.x == 1;
  __shared__ double red[8];
  __shared__ double bt_s[8][9];
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
  const int n = A.n_img;
  const int g = lane / 6, rr_ = lane - 6 * g;          // SpMV: 5 groups of 6 lanes, lane = (group, block row)
  const bool border = A.n_intr > 0;                     // 9 intrinsics unknowns owned by lanes 0..8 of warp 0 of CTA 0
  const bool owner = border && gw == 0;
  const bool olane = owner && lane < 9;
  double xi = 0.0, ri = 0.0, zi_ = 0.0, pi_ = 0.0, cp = 0.0;    // owner-lane state of the intrinsics part
  const bool coarse = A.cm > 0;
  // ---- init: x = 0, r = b, z = M^-1 r, p = 0
  if (coarse) {       // coarse residual rc = P' b, then yc = Ainv rc
    for (int row = gw; row < n; row += nw) {
      double Pl[CM];
#pragma unroll
      for (int k = 0; k < CM; ++k) Pl[k] = lane < 6 ? A.Pc[PCS * (size_t)row + CM * lane + k] : 0.0;
      coarse_restrict_add(Pl, lane < 6 ? A.b[6 * (size_t)row + lane] : 0.0, lane, true, A.rc + CM * A.agg[row]);
    }
    grid_barrier(bar_ctr, bar_target);
    coarse_solve(A, 0.0, A.rc, nullptr, gw, nw, lane);
    grid_barrier(bar_ctr, bar_target);
  }
  {
    double a_rz = 0.0, a_bb = 0.0;
    for (int row = gw; row < n; row += nw) {
      double rv = 0.0;
      if (lane < 6) rv = A.b[6 * (size_t)row + lane];
      double zl = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) { const double rc = __shfl_sync(0xffffffffu, rv, c); if (lane < 6) zl += A.Minv[36 * (size_t)row + 6 * lane + c] * rc; }
      if (coarse && lane < 6) {
        const double* yc = A.yc + CM * A.agg[row];
#pragma unroll
        for (int k = 0; k < CM; ++k) zl += A.Pc[PCS * (size_t)row + CM * lane + k] * __ldcg(yc + k);
      }
      if (lane < 6) {
        const size_t i = 6 * (size_t)row + lane;
        A.x[i] = 0.0; A.r[i] = rv; A.z[i] = zl; A.p0[i] = 0.0; A.p1[i] = 0.0;
        a_rz += rv * zl; a_bb += rv * rv;
      }
    }
    if (owner) {
      if (olane) ri = A.bi[lane];
#pragma unroll
      for (int m = 0; m < 9; ++m) { const double rm = __shfl_sync(0xffffffffu, ri, m); if (olane) zi_ += A.Cinv[9 * lane + m] * rm; }
      if (olane) { A.zi[lane] = zi_; A.pi0[lane] = 0.0; A.pi1[lane] = 0.0; a_rz += ri * zi_; a_bb += ri * ri; }
    }
    a_rz = block_sum_to_thread0(a_rz, red); a_bb = block_sum_to_thread0(a_bb, red);
    if (threadIdx.x == 0) { atomicAdd(A.sc + 2, a_rz); atomicAdd(A.sc + 6, a_bb); }
  }
  if (single) __syncthreads(); else grid_barrier(bar_ctr, bar_target);
  double rz = __ldcg(A.sc + 2), rz_old = 1.0;
  const double bb = __ldcg(A.sc + 6);
  int it = 0;
  if (bb > 0.0) {
    for (;;) {
      const double beta = it == 0 ? 0.0 : rz / rz_old;
      const double* p_old = (it & 1) ? A.p1 : A.p0; double* p_new = (it & 1) ? A.p0 : A.p1;
      const int nxt = (it + 1) & 1;
      if (gw == 0 && lane == 0) { A.sc[2 + nxt] = 0.0; A.sc[4 + nxt] = 0.0; }
      // intrinsics part of the search direction, formed on the fly by every warp: pin[m] in lane m
      double pin = 0.0;
      if (border && lane < 9) pin = __ldcg(A.zi + lane) + beta * __ldcg(((it & 1) ? A.pi1 : A.pi0) + lane);
      double btl = 0.0;                                 // this warp's share of B' p (lane m < 9)
      // ---- SpMV with p formed on the fly: Ap = S (z + beta p_old)
      PCG_STAMP(0);
      double acc = 0.0;
      for (int row = gw; row < n; row += nw) {
        double y = 0.0;
        {
          // four rounds of five blocks at a time: all index loads, then all gathers and block rows, are in flight together
          const int e0r = A.row_start[row], e1r = A.row_start[row + 1];
          for (int base = e0r; base < e1r; base += 20) {
            int col[4], bid[4]; bool okk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int e = base + 5 * k + g;
              okk[k] = lane < 30 && e < e1r;
              col[k] = okk[k] ? A.row_col[e] : 0; bid[k] = okk[k] ? A.row_blk[e] : 0;
            }
            double pv1[4], bk[4][6];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              pv1[k] = 0.0;
#pragma unroll
              for (int c = 0; c < 6; ++c) bk[k][c] = 0.0;
              if (okk[k]) {
                const bool tr = bid[k] < 0;
                const double* B = A.S + 36 * (size_t)(tr ? -bid[k] - 1 : bid[k]);
                const size_t ci = 6 * (size_t)col[k] + rr_;
                pv1[k] = __ldcg(A.z + ci) + beta * __ldcg(p_old + ci);
                if (!tr) {
                  const double2* B2 = reinterpret_cast<const double2*>(B + 6 * rr_);
                  const double2 b0 = B2[0], b1 = B2[1], b2 = B2[2];
                  bk[k][0] = b0.x; bk[k][1] = b0.y; bk[k][2] = b1.x; bk[k][3] = b1.y; bk[k][4] = b2.x; bk[k][5] = b2.y;
                } else {
#pragma unroll
                  for (int c = 0; c < 6; ++c) bk[k][c] = B[6 * c + rr_];
                }
              }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (base + 5 * k < e1r) {                        // warp-uniform
#pragma unroll
                for (int c = 0; c < 6; ++c) { const double xc = __shfl_sync(0xffffffffu, pv1[k], 6 * g + c); y += bk[k][c] * xc; }
              }
            }
          }
        }
        double t = y + __shfl_down_sync(0xffffffffu, y, 12);
        t += __shfl_down_sync(0xffffffffu, t, 6);
        t += __shfl_down_sync(0xffffffffu, y, 24);
        double pn = 0.0, bp = 0.0;
        const size_t i = 6 * (size_t)row + (lane < 6 ? lane : 0);
        if (lane < 6) pn = __ldcg(A.z + i) + beta * __ldcg(p_old + i);
        if (border) {
          const double* Ba = A.Bm + 54 * (size_t)row;
#pragma unroll
          for (int m = 0; m < 9; ++m) { const double pm = __shfl_sync(0xffffffffu, pin, m); if (lane < 6) bp += Ba[9 * lane + m] * pm; }
#pragma unroll
          for (int r = 0; r < 6; ++r) { const double pr = __shfl_sync(0xffffffffu, pn, r); if (lane < 9) btl += Ba[9 * r + lane] * pr; }
        }
        if (lane < 6) { p_new[i] = pn; A.Ap[i] = t + bp; acc += pn * (t + bp) + pn * bp; }
        if (coarse) {     // qc += P' (A p): the coarse residual follows the recursion r -= alpha A p without another pass
          double Pl[CM];
#pragma unroll
          for (int k = 0; k < CM; ++k) Pl[k] = lane < 6 ? A.Pc[PCS * (size_t)row + CM * lane + k] : 0.0;
          coarse_restrict_add(Pl, lane < 6 ? t + bp : 0.0, lane, true, A.qc + CM * A.agg[row]);      // (bp: border term, 0 without refined intrinsics)
        }
      }
      if (border) {
        // B' p: block-level reduction, then 9 atomics per CTA into the parity slot
        if (lane < 9) bt_s[wib][lane] = btl;
        if (owner) {      // C p_i and the intrinsics term of p'Ap
          cp = 0.0;
#pragma unroll
          for (int m = 0; m < 9; ++m) { const double pm = __shfl_sync(0xffffffffu, pin, m); if (olane) cp += A.Cm[9 * lane + m] * pm; }
          if (olane) { pi_ = pin; ((it & 1) ? A.pi0 : A.pi1)[lane] = pin; acc += pin * cp; }
        }
      }
      PCG_STAMP(1);
      acc = block_sum_to_thread0(acc, red);          // (contains __syncthreads: bt_s is complete afterwards)
      if (threadIdx.x == 0) atomicAdd(A.sc + (it & 1), acc);
      if (border && threadIdx.x < 9) { double sb = 0.0; for (int w = 0; w < wpb; ++w) sb += bt_s[w][threadIdx.x]; atomicAdd(A.bt + 9 * (it & 1) + threadIdx.x, sb); }
      PCG_STAMP(2);
      if (single) __syncthreads(); else grid_barrier(bar_ctr, bar_target);
      PCG_STAMP(3);
      const double pAp = __ldcg(A.sc + (it & 1));
      const double alpha = pAp > 0.0 ? rz / pAp : 0.0;
      if (gw == 0 && lane == 0) A.sc[nxt] = 0.0;
      // ---- x += alpha p ; r -= alpha Ap ; z = Minv r
      double a_rz = 0.0, a_rr = 0.0;
      for (int row = gw; row < n; row += nw) {
        double rv = 0.0;
        const size_t i = 6 * (size_t)row + (lane < 6 ? lane : 0);
        if (lane < 6) {
          A.x[i] += alpha * p_new[i];
          rv = A.r[i] - alpha * A.Ap[i];
          A.r[i] = rv;
        }
        if (coarse) continue;                           // z needs the coarse correction: second pass below
        double zl = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) { const double rc = __shfl_sync(0xffffffffu, rv, c); if (lane < 6) zl += A.Minv[36 * (size_t)row + 6 * lane + c] * rc; }
        if (lane < 6) { A.z[i] = zl; a_rz += rv * zl; a_rr += rv * rv; }
      }
      if (coarse) {
        coarse_solve(A, alpha, A.rc + (size_t)(it & 1) * A.cm, A.rc + (size_t)nxt * A.cm, gw, nw, lane);
        grid_barrier(bar_ctr, bar_target);
        for (int j = gw * 32 + lane; j < A.cm; j += nw * 32) A.qc[j] = 0.0;       // next accumulation starts after the barrier below
        for (int row = gw; row < n; row += nw) {
          const size_t i = 6 * (size_t)row + (lane < 6 ? lane : 0);
          const double rv = lane < 6 ? A.r[i] : 0.0;
          double zl = 0.0;
#pragma unroll
          for (int c = 0; c < 6; ++c) { const double rc = __shfl_sync(0xffffffffu, rv, c); if (lane < 6) zl += A.Minv[36 * (size_t)row + 6 * lane + c] * rc; }
          if (lane < 6) {
            const double* yc = A.yc + CM * A.agg[row];
#pragma unroll
            for (int k = 0; k < CM; ++k) zl += A.Pc[PCS * (size_t)row + CM * lane + k] * __ldcg(yc + k);
            A.z[i] = zl; a_rz += rv * zl; a_rr += rv * rv;
          }
        }
      }
      if (owner) {
        if (olane) { const double api = __ldcg(A.bt + 9 * (it & 1) + lane) + cp; xi += alpha * pi_; ri -= alpha * api; A.bt[9 * nxt + lane] = 0.0; }
        double zn = 0.0;
#pragma unroll
        for (int m = 0; m < 9; ++m) { const double rm = __shfl_sync(0xffffffffu, ri, m); if (olane) zn += A.Cinv[9 * lane + m] * rm; }
        if (olane) { zi_ = zn; A.zi[lane] = zn; a_rz += ri * zn; a_rr += ri * ri; }
      }
      a_rz = block_sum_to_thread0(a_rz, red); a_rr = block_sum_to_thread0(a_rr, red);
      if (threadIdx.x == 0) { atomicAdd(A.sc + 2 + nxt, a_rz); atomicAdd(A.sc + 4 + nxt, a_rr); }
      PCG_STAMP(4);
      if (single) __syncthreads(); else grid_barrier(bar_ctr, bar_target);
      PCG_STAMP(5);
      rz_old = rz; rz = __ldcg(A.sc + 2 + nxt);
      const double rr = __ldcg(A.sc + 4 + nxt);
      ++it;
      if (rr <= A.tol2 * bb || it >= A.max_iter || !(rr == rr)) break;
    }
  }
  if (olane) A.xi[lane] = xi;
  if (gw == 0 && lane == 0) A.ic[1] = it;
}

// ---- K3 (persistent, cached form): small/medium reduced systems --------------------------------
// Same algorithm as k_pcg_persistent, but each warp owns ONE block-row for the whole solve: the row's
// 6x6 blocks (already oriented), its column indices and its Minv block live in shared memory, and its
// slices of x, r, z, p, Ap live in registers.  Per iteration a warp only gathers z/p of its neighbour
// rows from L2 (all loads issued back to back), so an iteration costs about two grid barriers.
// Usable when every CTA's rows fit in shared memory (host checks); otherwise the streaming kernel runs.
constexpr int PCG_MAXR = 16;          // rounds of 5 blocks held in registers per chunk

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_pcg_cached(PcgArgs A, int e_cap) {
  namespace cg = cooperative_groups;
  unsigned int* bar_ctr = reinterpret_cast<unsigned int*>(A.ic + 3); unsigned int bar_target = 0;
  constexpr int NW = THREADS / 32;
  extern __shared__ double sm[];
  double* sB = sm;                                   // [e_cap][36]
  double* sM = sB + (size_t)e_cap * 36;              // [NW][36]
  int* sC = reinterpret_cast<int*>(sM + NW * 36);    // [e_cap]
  __shared__ double red[NW];
  const bool single = gridDim.x == 1;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int n = A.n_img;
  const int row = blockIdx.x * NW + wib;
  const bool valid = row < n;
  const int g = lane / 6, r6 = lane - 6 * g;
  const bool act = valid && lane < 6;
  const int first_row = min(blockIdx.x * NW, n);
  const int e_blk0 = A.row_start[first_row];
  int e0 = 0, e1 = 0;
  if (valid) { e0 = A.row_start[row] - e_blk0; e1 = A.row_start[row + 1] - e_blk0; }
  // ---- one-time staging of this row's blocks (oriented) and preconditioner block
  for (int idx = lane; idx < (e1 - e0) * 36; idx += 32) {
    const int e = idx / 36, k = idx - 36 * e, rr = k / 6, cc = k - 6 * rr;
    const int bid = A.row_blk[e_blk0 + e0 + e];
    const bool tr = bid < 0;
    const double* B = A.S + 36 * (size_t)(tr ? -bid - 1 : bid);
    sB[(size_t)(e0 + e) * 36 + k] = tr ? B[6 * cc + rr] : B[k];
  }
  for (int e = e0 + lane; e < e1; e += 32) sC[e] = A.row_col[e_blk0 + e];
  if (valid) for (int k = lane; k < 36; k += 32) sM[wib * 36 + k] = A.Minv[36 * (size_t)row + k];
  __syncwarp();
  // ---- init
  double xl = 0.0, rl = 0.0, zl = 0.0, pl = 0.0, apl = 0.0;
  const size_t gi = 6 * (size_t)(valid ? row : 0) + (lane < 6 ? lane : 0);
  const bool coarse = A.cm > 0;
  const int gw = blockIdx.x * NW + wib, nwt = gridDim.x * NW;
  double Pl[CM]; int ga = 0;                          // this row of the prolongation block, aggregate offset
#pragma unroll
  for (int k = 0; k < CM; ++k) Pl[k] = (coarse && act) ? A.Pc[PCS * (size_t)row + CM * lane + k] : 0.0;
  if (coarse && valid) ga = CM * A.agg[row];
  {
    if (act) rl = A.b[gi];
    if (coarse) {
      coarse_restrict_add(Pl, rl, lane, valid, A.rc + ga);
      grid_barrier(bar_ctr, bar_target);
      coarse_solve(A, 0.0, A.rc, nullptr, gw, nwt, lane);
      grid_barrier(bar_ctr, bar_target);
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) { const double rc = __shfl_sync(0xffffffffu, rl, c); if (act) zl += sM[wib * 36 + 6 * lane + c] * rc; }
    if (coarse && act) {
#pragma unroll
      for (int k = 0; k < CM; ++k) zl += Pl[k] * __ldcg(A.yc + ga + k);
    }
    if (act) { A.z[gi] = zl; A.p0[gi] = 0.0; A.p1[gi] = 0.0; }
    double a_rz = act ? rl * zl : 0.0, a_bb = act ? rl * rl : 0.0;
    a_rz = block_sum_to_thread0(a_rz, red); a_bb = block_sum_to_thread0(a_bb, red);
    if (threadIdx.x == 0) { atomicAdd(A.sc + 2, a_rz); atomicAdd(A.sc + 6, a_bb); }
  }
  if (single) __syncthreads(); else grid_barrier(bar_ctr, bar_target);
  double rz = __ldcg(A.sc + 2), rz_old = 1.0;
  const double bb = __ldcg(A.sc + 6);
  int it = 0;
  if (bb > 0.0) {
    for (;;) {
      const double beta = it == 0 ? 0.0 : rz / rz_old;
      const double* p_old = (it & 1) ? A.p1 : A.p0; double* p_new = (it & 1) ? A.p0 : A.p1;
      const int nxt = (it + 1) & 1;
      if (blockIdx.x == 0 && threadIdx.x == 0) { A.sc[2 + nxt] = 0.0; A.sc[4 + nxt] = 0.0; }
      PCG_STAMP(0);
      // ---- SpMV: gather all neighbour slices first (independent L2 loads), then multiply from smem
      double y = 0.0;
      for (int base = e0; base < e1; base += 5 * PCG_MAXR) {
        double pv[PCG_MAXR];
#pragma unroll
        for (int k = 0; k < PCG_MAXR; ++k) {
          const int e = base + 5 * k + g;
          pv[k] = 0.0;
          if (lane < 30 && e < e1) { const size_t ci = 6 * (size_t)sC[e] + r6; pv[k] = __ldcg(A.z + ci) + beta * __ldcg(p_old + ci); }
        }
#pragma unroll
        for (int k = 0; k < PCG_MAXR; ++k) {
          if (base + 5 * k < e1) {                       // warp-uniform
            const int e = base + 5 * k + g;
            const bool ok = lane < 30 && e < e1;
            const double* Brow = sB + (size_t)(ok ? e : e0) * 36 + 6 * r6;
#pragma unroll
            for (int c = 0; c < 6; ++c) { const double xc = __shfl_sync(0xffffffffu, pv[k], 6 * g + c); if (ok) y += Brow[c] * xc; }
          }
        }
      }
      double t = y + __shfl_down_sync(0xffffffffu, y, 12);
      t += __shfl_down_sync(0xffffffffu, t, 6);
      t += __shfl_down_sync(0xffffffffu, y, 24);
      PCG_STAMP(1);
      double acc = 0.0;
      if (act) { pl = zl + beta * pl; apl = t; p_new[gi] = pl; acc = pl * t; }
      if (coarse) coarse_restrict_add(Pl, act ? t : 0.0, lane, valid, A.qc + ga);
      acc = block_sum_to_thread0(acc, red);
      if (threadIdx.x == 0) atomicAdd(A.sc + (it & 1), acc);
      PCG_STAMP(2);
      if (single) __syncthreads(); else grid_barrier(bar_ctr, bar_target);
      PCG_STAMP(3);
      const double pAp = __ldcg(A.sc + (it & 1));
      const double alpha = pAp > 0.0 ? rz / pAp : 0.0;
      if (blockIdx.x == 0 && threadIdx.x == 0) A.sc[nxt] = 0.0;
      // ---- update (registers only)
      if (act) { xl += alpha * pl; rl -= alpha * apl; }
      if (coarse) {
        coarse_solve(A, alpha, A.rc + (size_t)(it & 1) * A.cm, A.rc + (size_t)nxt * A.cm, gw, nwt, lane);
        grid_barrier(bar_ctr, bar_target);
        for (int j = gw * 32 + lane; j < A.cm; j += nwt * 32) A.qc[j] = 0.0;     // next accumulation starts after the barrier below
      }
      double zn = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) { const double rc = __shfl_sync(0xffffffffu, rl, c); if (act) zn += sM[wib * 36 + 6 * lane + c] * rc; }
      if (coarse && act) {
#pragma unroll
        for (int k = 0; k < CM; ++k) zn += Pl[k] * __ldcg(A.yc + ga + k);
      }
      if (act) { zl = zn; A.z[gi] = zl; }
      double a_rz = act ? rl * zl : 0.0, a_rr = act ? rl * rl : 0.0;
      a_rz = block_sum_to_thread0(a_rz, red); a_rr = block_sum_to_thread0(a_rr, red);
      if (threadIdx.x == 0) { atomicAdd(A.sc + 2 + nxt, a_rz); atomicAdd(A.sc + 4 + nxt, a_rr); }
      PCG_STAMP(4);
      if (single) __syncthreads(); else grid_barrier(bar_ctr, bar_target);
      PCG_STAMP(5);
      rz_old = rz; rz = __ldcg(A.sc + 2 + nxt);
      const double rr = __ldcg(A.sc + 4 + nxt);
      ++it;
      if (rr <= A.tol2 * bb || it >= A.max_iter || !(rr == rr)) break;
    }
  }
  if (act) A.x[gi] = xl;
  if (blockIdx.x == 0 && threadIdx.x == 0) A.ic[1] = it;
}

// ---- refined intrinsics: dense border of the reduced system ----------------------------------------------
// One intrinsics block of 9 per camera (bundle_adjustment.cc:246-247: camera_params is one parameter block per camera id,
// shared by all its images; refine_camera_params is the mapper's default, mapper.cc:878-886).  Unknowns of the reduced system
// become [poses (6 n_img) | intrinsics (9 n_cam)].  With E = d r / d intr of the observation's camera (scaled):
//   A_pc   = sum_{obs of p in camera c} E' Jp                   (9 x 3, stored per point and camera for the back-substitution)
//   C_cc'  = [c = c'] (sum_obs E'E + D_i^2) - sum_p A_pc V_p^-1 A_pc''      (9 x 9 blocks)
//   b_c    = sum_obs E'r - sum_p A_pc V_p^-1 g_p
//   B_ac'  = [c' = cam(a)] sum_{obs of a} Jc'E - sum_{obs of a} Y_obs A_pc''   (6 x 9), Y_obs = Jc'Jp V_p^-1
// intr_acc layout (doubles, n9 = 9 n_cam): [0, n9^2) C row-major | b (n9) | diag(E'E) (n9) | E'r, the unreduced gradient (n9)
__global__ void __launch_bounds__(128) k_colnorm_intr_img(int n_img, const int* __restrict__ cam_lo, const int* __restrict__ cam_hi, const int* __restrict__ cam_perm,
                                                          const double* __restrict__ ji /* indexed by global observation position */, double* __restrict__ img_sq /*[n_img][9]*/) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_img) return;
  double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int q = cam_lo[w] + lane; q < cam_hi[w]; q += 32) {
    const double* j = ji + 18 * (size_t)cam_perm[q];
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] += j[k] * j[k] + j[9 + k] * j[9 + k];
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) { const double v = warp_sum(a[k]); if (lane == 0) img_sq[9 * (size_t)w + k] = v; }
}
// one CTA per camera: column sums over its images in a fixed order -> Jacobi scale 1 / (1 + norm)
__global__ void __launch_bounds__(256) k_intr_scale(int n_img, const int* __restrict__ img_cam, const double* __restrict__ img_sq, double* __restrict__ scale_i) {
  __shared__ double red[32];
  const int c = blockIdx.x;
  double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n_img; i += blockDim.x) if (img_cam[i] == c) {
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] += img_sq[9 * (size_t)i + k];
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) { const double v = block_sum(a[k], red); if (threadIdx.x == 0) scale_i[9 * (size_t)c + k] = 1.0 / (1.0 + sqrt(v)); }
}

// thread per point of this rank: A_pc for every camera that sees the point, and the point's contributions to C, b, diag, gradient.
// cam_mask[p]: bit c set = A_pc is non-zero (cameras beyond 32 are refused at session creation).
__global__ void __launch_bounds__(128) k_schur_intr_point(
    int n_pt, int n_cam, const int* __restrict__ pt_start, const int* __restrict__ obs_img, const int* __restrict__ img_cam,
    const double* __restrict__ rec, const double* __restrict__ ji,
    const double* __restrict__ scale_p, const double* __restrict__ scale_i, const double* __restrict__ Vinv, const double* __restrict__ gp,
    double* __restrict__ Apc, unsigned* __restrict__ cam_mask, double* __restrict__ intr_acc) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int n9 = 9 * n_cam;
  double* Cm = intr_acc; double* bi = intr_acc + (size_t)n9 * n9; double* dg = bi + n9; double* gr = dg + n9;
  const bool live = p < n_pt;
  const int o0 = live ? pt_start[p] : 0, o1 = live ? pt_start[p + 1] : 0;
  double sp[3] = {1, 1, 1}, I[6] = {0, 0, 0, 0, 0, 0}, g3[3] = {0, 0, 0};
  if (live) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { sp[k] = scale_p[3 * (size_t)p + k]; g3[k] = gp[3 * (size_t)p + k]; }
#pragma unroll
    for (int k = 0; k < 6; ++k) I[k] = Vinv[6 * (size_t)p + k];
  }
  unsigned mask = 0;
  for (int c = 0; c < n_cam; ++c) {
    double A[9][3], Cc[45], rh[9], ud[9], gu[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) { A[a][0] = A[a][1] = A[a][2] = 0.0; rh[a] = 0.0; ud[a] = 0.0; gu[a] = 0.0; }
#pragma unroll
    for (int k = 0; k < 45; ++k) Cc[k] = 0.0;
    bool any = false;
    double si[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) si[k] = scale_i[9 * (size_t)c + k];
    for (int o = o0; o < o1; ++o) {
      if (img_cam[obs_img[o]] != c) continue;
      any = true;
      const double2* r2 = reinterpret_cast<const double2*>(rec + REC * (size_t)o);
      const double2 rr = r2[0], q7 = r2[7], q8 = r2[8], q9 = r2[9];
      const double jp0[3] = { q7.x * sp[0], q7.y * sp[1], q8.x * sp[2] }, jp1[3] = { q8.y * sp[0], q9.x * sp[1], q9.y * sp[2] };
      const double* j = ji + 18 * (size_t)o;
      int k = 0;
#pragma unroll
      for (int a = 0; a < 9; ++a) {
        const double e0 = j[a] * si[a], e1 = j[9 + a] * si[a];
        A[a][0] += e0 * jp0[0] + e1 * jp1[0]; A[a][1] += e0 * jp0[1] + e1 * jp1[1]; A[a][2] += e0 * jp0[2] + e1 * jp1[2];
        const double ga = e0 * rr.x + e1 * rr.y;
        gu[a] += ga; rh[a] += ga; ud[a] += e0 * e0 + e1 * e1;
#pragma unroll
        for (int b = a; b < 9; ++b, ++k) Cc[k] += e0 * (j[b] * si[b]) + e1 * (j[9 + b] * si[b]);
      }
    }
    if (any) mask |= 1u << c;
    // t = A_pc V^-1;  b_c -= t g_p;  C_cc -= t A_pc';  C_cc' -= t A_pc'' (c' < c, A_pc' written by this thread in an earlier round)
    double t[9][3];
#pragma unroll
    for (int a = 0; a < 9; ++a) {
      t[a][0] = A[a][0] * I[0] + A[a][1] * I[1] + A[a][2] * I[2];
      t[a][1] = A[a][0] * I[1] + A[a][1] * I[3] + A[a][2] * I[4];
      t[a][2] = A[a][0] * I[2] + A[a][1] * I[4] + A[a][2] * I[5];
      rh[a] -= t[a][0] * g3[0] + t[a][1] * g3[1] + t[a][2] * g3[2];
    }
    {
      int k = 0;
#pragma unroll
      for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int b = a; b < 9; ++b, ++k) Cc[k] -= t[a][0] * A[b][0] + t[a][1] * A[b][1] + t[a][2] * A[b][2];
    }
    if (live && any) {
      double* dst = Apc + 27 * ((size_t)p * n_cam + c);
#pragma unroll
      for (int a = 0; a < 9; ++a) { dst[3 * a] = A[a][0]; dst[3 * a + 1] = A[a][1]; dst[3 * a + 2] = A[a][2]; }
    }
    // warp reductions, one atomic per warp and value (skipped when no lane of the warp has the camera)
    const unsigned warp_any = __ballot_sync(0xffffffffu, any);
    if (warp_any) {
      int k = 0;
#pragma unroll
      for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int b = a; b < 9; ++b, ++k) {
          const double v = warp_sum(Cc[k]);
          if (lane == 0 && v != 0.0) { atomicAdd(Cm + (size_t)(9 * c + a) * n9 + 9 * c + b, v); if (b != a) atomicAdd(Cm + (size_t)(9 * c + b) * n9 + 9 * c + a, v); }
        }
#pragma unroll
      for (int a = 0; a < 9; ++a) {
        const double v0 = warp_sum(rh[a]), v1 = warp_sum(ud[a]), v2 = warp_sum(gu[a]);
        if (lane == 0) { atomicAdd(bi + 9 * c + a, v0); atomicAdd(dg + 9 * c + a, v1); atomicAdd(gr + 9 * c + a, v2); }
      }
    }
    for (int c2 = 0; c2 < c; ++c2) {
      const bool both = any && ((mask >> c2) & 1u);
      if (!__ballot_sync(0xffffffffu, both)) continue;
      const double* A2 = Apc + 27 * ((size_t)(live ? p : 0) * n_cam + c2);
#pragma unroll
      for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int b = 0; b < 9; ++b) {
          double v = both ? -(t[a][0] * A2[3 * b] + t[a][1] * A2[3 * b + 1] + t[a][2] * A2[3 * b + 2]) : 0.0;
          v = warp_sum(v);
          if (lane == 0 && v != 0.0) { atomicAdd(Cm + (size_t)(9 * c + a) * n9 + 9 * c2 + b, v); atomicAdd(Cm + (size_t)(9 * c2 + b) * n9 + 9 * c + a, v); }
        }
    }
  }
  if (live) cam_mask[p] = mask;
}

// warp per image: B_ac' (6 x 9) for every camera c'
__global__ void __launch_bounds__(128) k_schur_cam_intr(
    int n_img, int n_cam, const int* __restrict__ cam_lo, const int* __restrict__ cam_hi, const int* __restrict__ cam_perm, const int* __restrict__ obs_pt,
    const int* __restrict__ img_cam, const double* __restrict__ rec, const double* __restrict__ ji, const double* __restrict__ scale_c, const double* __restrict__ scale_p,
    const double* __restrict__ scale_i, const double* __restrict__ Vinv, const double* __restrict__ Apc, const unsigned* __restrict__ cam_mask, double* __restrict__ Bm) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_img) return;
  const int cam = img_cam[w];
  double sc[6], si[9];
#pragma unroll
  for (int k = 0; k < 6; ++k) sc[k] = scale_c[6 * (size_t)w + k];
#pragma unroll
  for (int k = 0; k < 9; ++k) si[k] = scale_i[9 * (size_t)cam + k];
  for (int c2 = 0; c2 < n_cam; ++c2) {
    double acc[6][9];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int m = 0; m < 9; ++m) acc[a][m] = 0.0;
    for (int q = cam_lo[w] + lane; q < cam_hi[w]; q += 32) {
      const int o = cam_perm[q], p = obs_pt[o];
      const bool has = (cam_mask[p] >> c2) & 1u;
      if (!has && c2 != cam) continue;
      const double sp[3] = { scale_p[3 * (size_t)p], scale_p[3 * (size_t)p + 1], scale_p[3 * (size_t)p + 2] };
      double Jc[2][6], Jp[2][3];
      load_scaled(rec, o, sc, sp, Jc, Jp);
      if (c2 == cam) {
        const double* j = ji + 18 * (size_t)o;
#pragma unroll
        for (int m = 0; m < 9; ++m) {
          const double e0 = j[m] * si[m], e1 = j[9 + m] * si[m];
#pragma unroll
          for (int a = 0; a < 6; ++a) acc[a][m] += Jc[0][a] * e0 + Jc[1][a] * e1;
        }
      }
      if (has) {
        double I[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) I[k] = Vinv[6 * (size_t)p + k];
        const double* Ap = Apc + 27 * ((size_t)p * n_cam + c2);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          const double w0 = Jc[0][a] * Jp[0][0] + Jc[1][a] * Jp[1][0];
          const double w1 = Jc[0][a] * Jp[0][1] + Jc[1][a] * Jp[1][1];
          const double w2 = Jc[0][a] * Jp[0][2] + Jc[1][a] * Jp[1][2];
          const double y0 = w0 * I[0] + w1 * I[1] + w2 * I[2], y1 = w0 * I[1] + w1 * I[3] + w2 * I[4], y2 = w0 * I[2] + w1 * I[4] + w2 * I[5];
#pragma unroll
          for (int m = 0; m < 9; ++m) acc[a][m] -= y0 * Ap[3 * m] + y1 * Ap[3 * m + 1] + y2 * Ap[3 * m + 2];
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int m = 0; m < 9; ++m) { const double v = warp_sum(acc[a][m]); if (lane == 0) Bm[54 * ((size_t)w * n_cam + c2) + 9 * a + m] = v; }
  }
}

// after the exchange step: LM diagonal of the intrinsics onto C, gradient, its max-norm.  One thread per intrinsics unknown.
__global__ void k_intr_finalize(int n9, LMDiag lm, const double* __restrict__ scale_i, double* __restrict__ intr_acc,
                                double* __restrict__ gi_out, double* __restrict__ di_out, double* __restrict__ gmax) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n9) return;
  double* Cm = intr_acc; const double* dg = intr_acc + (size_t)n9 * n9 + n9; const double* gr = dg + n9;
  const double d = fmin(fmax(dg[a], lm.min_diag), lm.max_diag) / lm.radius;
  di_out[a] = d; gi_out[a] = gr[a];
  Cm[(size_t)a * n9 + a] += d;
  const double gm = fabs(gr[a] / scale_i[a]);
  if (gm > 0.0) atomic_max_nonneg(gmax, gm);
}

// candidate intrinsics + their share of ||step||^2 and of the model cost change (one warp, fixed order)
__global__ void k_update_intr(int n_cam, const double* __restrict__ yi, const double* __restrict__ scale_i, const double* __restrict__ gi, const double* __restrict__ di,
                              const double* __restrict__ intr, double* __restrict__ intr2, double* __restrict__ part2) {
  if (threadIdx.x != 0) return;
  double sn = 0.0, mc = 0.0;
  for (int c = 0; c < n_cam; ++c)
    for (int k = 0; k < 9; ++k) {
      const double y = yi[9 * c + k], e = -y * scale_i[9 * c + k];
      intr2[MM_INTR_STRIDE * c + k] = intr[MM_INTR_STRIDE * c + k] + e; sn += e * e; mc += 0.5 * y * (gi[9 * c + k] + di[9 * c + k] * y);
    }
  part2[0] = sn; part2[1] = mc;
}

// ---- K4a: back-substitution for the points + candidate point parameters --------------------
// y_p = V'^-1 (g_p - sum_i W_i' y_c[img_i]); delta = -y_p * s_p.  Partials: [0] step_norm2,
// [1] model cost change 1/2 y (g + D y)  (valid because the linear system is solved to pcg_tolerance).
__global__ void __launch_bounds__(128) k_backsub(
    int n_pt, const int* __restrict__ pt_start, const int* __restrict__ obs_img, const double* __restrict__ rec,
    const double* __restrict__ scale_c, const double* __restrict__ scale_p, const double* __restrict__ Vinv,
    const double* __restrict__ gp, const double* __restrict__ dp, const double* __restrict__ yc,
    const double* __restrict__ pts, double* __restrict__ pts2, double* __restrict__ part /*[2*grid]*/,
    const double* __restrict__ Apc = nullptr, const double* __restrict__ yi = nullptr, int n_cam = 0, const unsigned* __restrict__ cam_mask = nullptr) {
  __shared__ double red[32];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double sn = 0.0, mc = 0.0;
  if (p < n_pt) {
    const double sp[3] = { scale_p[3 * (size_t)p], scale_p[3 * (size_t)p + 1], scale_p[3 * (size_t)p + 2] };
    double t[3] = { gp[3 * (size_t)p], gp[3 * (size_t)p + 1], gp[3 * (size_t)p + 2] };
    const double g0 = t[0], g1 = t[1], g2 = t[2];
    for (int o = pt_start[p]; o < pt_start[p + 1]; ++o) {
      const int ia = obs_img[o];
      double Jc[2][6], Jp[2][3];
      load_scaled(rec, o, scale_c + 6 * (size_t)ia, sp, Jc, Jp);
      double q0 = 0.0, q1 = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) { const double y = yc[6 * (size_t)ia + k]; q0 += Jc[0][k] * y; q1 += Jc[1][k] * y; }
#pragma unroll
      for (int k = 0; k < 3; ++k) t[k] -= Jp[0][k] * q0 + Jp[1][k] * q1;
    }
    if (Apc) {        // - sum_c A_pc' y_c (refined intrinsics)
      const unsigned mask = cam_mask[p];
      for (int c = 0; c < n_cam; ++c) if ((mask >> c) & 1u) {
        const double* Ap = Apc + 27 * ((size_t)p * n_cam + c);
#pragma unroll
        for (int a = 0; a < 9; ++a) { const double y = yi[9 * c + a]; t[0] -= Ap[3 * a] * y; t[1] -= Ap[3 * a + 1] * y; t[2] -= Ap[3 * a + 2] * y; }
      }
    }
    const double* I = Vinv + 6 * (size_t)p;
    const double y0 = I[0] * t[0] + I[1] * t[1] + I[2] * t[2];
    const double y1 = I[1] * t[0] + I[3] * t[1] + I[4] * t[2];
    const double y2 = I[2] * t[0] + I[4] * t[1] + I[5] * t[2];
    const double e0 = -y0 * sp[0], e1 = -y1 * sp[1], e2 = -y2 * sp[2];
    pts2[3 * (size_t)p] = pts[3 * (size_t)p] + e0; pts2[3 * (size_t)p + 1] = pts[3 * (size_t)p + 1] + e1; pts2[3 * (size_t)p + 2] = pts[3 * (size_t)p + 2] + e2;
    sn = e0 * e0 + e1 * e1 + e2 * e2;
    mc = 0.5 * (y0 * (g0 + dp[3 * (size_t)p] * y0) + y1 * (g1 + dp[3 * (size_t)p + 1] * y1) + y2 * (g2 + dp[3 * (size_t)p + 2] * y2));
  }
  sn = block_sum(sn, red); mc = block_sum(mc, red);
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = sn; part[2 * blockIdx.x + 1] = mc; }
}

// ---- K4b: candidate camera parameters -------------------------------------------------------
__global__ void k_update_cam(int n, const double* __restrict__ yc, const double* __restrict__ scale_c,
                             const double* __restrict__ gc, const double* __restrict__ dc,
                             const double* __restrict__ poses, double* __restrict__ poses2, double* __restrict__ part) {
  __shared__ double red[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double sn = 0.0, mc = 0.0;
  if (i < n) {
    const double y = yc[i];
    const double e = -y * scale_c[i];
    poses2[i] = poses[i] + e;
    sn = e * e;
    mc = 0.5 * y * (gc[i] + dc[i] * y);
  }
  sn = block_sum(sn, red); mc = block_sum(mc, red);
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = sn; part[2 * blockIdx.x + 1] = mc; }
}

__global__ void k_reduce_pairs(const double* __restrict__ partA, int nA, const double* __restrict__ partB, int nB,
                               double* __restrict__ out /*[2]*/) {
  __shared__ double red[32];
  double s0 = 0.0, s1 = 0.0;
  for (int i = threadIdx.x; i < nA; i += blockDim.x) { s0 += partA[2 * i]; s1 += partA[2 * i + 1]; }
  for (int i = threadIdx.x; i < nB; i += blockDim.x) { s0 += partB[2 * i]; s1 += partB[2 * i + 1]; }
  s0 = block_sum(s0, red); s1 = block_sum(s1, red);
  if (threadIdx.x == 0) { out[0] = s0; out[1] = s1; }
}

// sum of squares of the active parameters (x_norm of the reduced program)
__global__ void k_xnorm(int n_c, const double* __restrict__ poses, const double* __restrict__ pose_mask,
                        int n_pt, const double* __restrict__ pts, const double* __restrict__ pt_mask, double* __restrict__ part,
                        const double* __restrict__ intr = nullptr, const double* __restrict__ intr_mask = nullptr, int n_cam = 0) {
  __shared__ double red[32];
  double s = 0.0;
  if (intr && blockIdx.x == 0) for (int i = threadIdx.x; i < 9 * n_cam; i += blockDim.x) { const double v = intr[MM_INTR_STRIDE * (i / 9) + i % 9]; s += intr_mask[i] * v * v; }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_c; i += (int64_t)gridDim.x * blockDim.x)
    s += pose_mask[i] * poses[i] * poses[i];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < 3 * (int64_t)n_pt; i += (int64_t)gridDim.x * blockDim.x)
    s += pt_mask[i / 3] * pts[i] * pts[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// mean raw residual norm per point (bundle_adjustment.cc:575-598)
__global__ void k_point_errors(int n_pt, const int* __restrict__ pt_start, const double2* __restrict__ obs_xy,
                               const int* __restrict__ obs_img, const double* __restrict__ aux, const double* __restrict__ pts,
                               const double* __restrict__ intr, const int* __restrict__ img_cam, const int* __restrict__ cam_model,
                               double* __restrict__ pt_err) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pt) return;
  const int o0 = pt_start[p], o1 = pt_start[p + 1];
  if (o1 == o0) return;
  const double X0 = pts[3 * (size_t)p], X1 = pts[3 * (size_t)p + 1], X2 = pts[3 * (size_t)p + 2];
  double s = 0.0;
  for (int o = o0; o < o1; ++o) {
    const int img = obs_img[o];
    const double* a = aux + AUX * (size_t)img;
    const double xc = a[0] * X0 + a[1] * X1 + a[2] * X2 + a[9];
    const double yc = a[3] * X0 + a[4] * X1 + a[5] * X2 + a[10];
    const double zc = a[6] * X0 + a[7] * X1 + a[8] * X2 + a[11];
    const int cam = img_cam[img];
    double u, v;
    world2image<false>(cam_model[cam], intr + MM_INTR_STRIDE * (size_t)cam, xc, yc, zc, u, v, nullptr, nullptr);
    const double2 xy = obs_xy[o];
    s += sqrt((u - xy.x) * (u - xy.x) + (v - xy.y) * (v - xy.y)) / (double)(o1 - o0);
  }
  pt_err[p] = s;
}

}  // namespace mm
