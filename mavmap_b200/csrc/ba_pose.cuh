// ba_pose.cuh — pose_refinement (mavmap/mavmap src/base3d/bundle_adjustment.cc:139-225) as ONE kernel launch.
//
// The reference calls it once per registered image (sequential_mapper.cc:716) on a few thousand 2D-3D pairs with the
// points (:187) and intrinsics (:193) constant: 6 unknowns, DENSE_QR.  Going through the general session (sorts, block
// structure, ~70 buffers, a dozen launches per LM iteration) costs tens of milliseconds for a problem whose arithmetic
// takes microseconds, so this path runs the WHOLE Levenberg-Marquardt loop inside one CTA: every pass over the
// observations is a block-strided loop with a shuffle/shared-memory reduction of cost | J'J (21) | J'r (6), the 6 x 6
// damped normal equations are solved by Cholesky in thread 0, and the trust-region logic is the same restatement of
// Ceres 1.8 as lm_start()/lm_iterate() in ba.cu (SURVEY 8a-a3').  Parity is tested against the oracle and against the
// general engine (tests/test_gpu_ba.py).
#pragma once
#include "ba_kernels.cuh"

namespace mm {

struct PoseLM {           // shared-memory state of the single-CTA solver
  double x[6], x2[6], xprev[6], R[9], Jl[9], scale[6], A[21], g[6], y[6], D[6], acc[28];
  double cost, cost2, radius, decrease_factor, x_norm, abs_gtol, gmax, step_norm, mcc;
  int iter, n_invalid, done, want;
};

// one pass over the observations at pose `x`: acc[0] = cost, and if WITH_J acc[1..21] = upper triangle of J'J, acc[22..27] = J'r
template <bool WITH_J>
__device__ void pose_pass(int n, const double2* __restrict__ uv, const double* __restrict__ X, int model, const double* __restrict__ intr,
                          const double* x, PoseLM* sm, LossParams L, double (*wred)[28]) {
  if (threadIdx.x == 0) rotation_and_left_jacobian(x, sm->R, sm->Jl);
  __syncthreads();
  double R[9], Jl[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { R[k] = sm->R[k]; Jl[k] = sm->Jl[k]; }
  const double t0 = x[3], t1 = x[4], t2 = x[5];
  double a[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) a[k] = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double X0 = X[3 * (size_t)i], X1 = X[3 * (size_t)i + 1], X2 = X[3 * (size_t)i + 2];
    const double2 o = uv[i];
    const double Y0 = R[0] * X0 + R[1] * X1 + R[2] * X2, Y1 = R[3] * X0 + R[4] * X1 + R[5] * X2, Y2 = R[6] * X0 + R[7] * X1 + R[8] * X2;
    double u, v, dX[2][3];
    world2image<WITH_J>(model, intr, Y0 + t0, Y1 + t1, Y2 + t2, u, v, dX, nullptr);
    const double r0 = u - o.x, r1 = v - o.y;
    double rho0, sr;
    loss_eval(L, r0 * r0 + r1 * r1, rho0, sr);
    a[0] += 0.5 * rho0;
    if (WITH_J) {
      double M[3][3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double c0 = Jl[k], c1 = Jl[3 + k], c2 = Jl[6 + k];
        M[0][k] = c1 * Y2 - c2 * Y1; M[1][k] = c2 * Y0 - c0 * Y2; M[2][k] = c0 * Y1 - c1 * Y0;
      }
      double J[2][6];
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const double d0 = dX[row][0], d1 = dX[row][1], d2 = dX[row][2];
        J[row][0] = (d0 * M[0][0] + d1 * M[1][0] + d2 * M[2][0]) * sr;
        J[row][1] = (d0 * M[0][1] + d1 * M[1][1] + d2 * M[2][1]) * sr;
        J[row][2] = (d0 * M[0][2] + d1 * M[1][2] + d2 * M[2][2]) * sr;
        J[row][3] = d0 * sr; J[row][4] = d1 * sr; J[row][5] = d2 * sr;
      }
      const double s0 = sr * r0, s1 = sr * r1;
      int k = 1;
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = p; q < 6; ++q, ++k) a[k] += J[0][p] * J[0][q] + J[1][p] * J[1][q];
#pragma unroll
      for (int p = 0; p < 6; ++p) a[22 + p] += J[0][p] * s0 + J[1][p] * s1;
    }
  }
  constexpr int NV = WITH_J ? 28 : 1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) { const double v = warp_sum(a[k]); if (lane == 0) wred[w][k] = v; }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) s += wred[ww][threadIdx.x];
    sm->acc[threadIdx.x] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ void pose_trace_push(mm_ba_summary* S, const PoseLM* sm, int accepted) {
  const int i = S->num_iterations;
  if (i < MM_BA_TRACE_MAX) { S->trace_cost[i] = sm->cost; S->trace_radius[i] = sm->radius; S->trace_gradient_max_norm[i] = sm->gmax; S->trace_accepted[i] = accepted; S->trace_linear_iterations[i] = i ? 1 : 0; }
  S->num_iterations = i + 1;
  if (sm->cost < S->final_cost) S->final_cost = sm->cost;
}
// after a linearization: unscaled J'J and J'r -> state (thread 0)
__device__ __forceinline__ void pose_take_linearization(PoseLM* sm) {
  sm->cost = sm->acc[0];
  for (int k = 0; k < 21; ++k) sm->A[k] = sm->acc[1 + k];
  double gm = 0.0, xn = 0.0;
  for (int k = 0; k < 6; ++k) { sm->g[k] = sm->acc[22 + k]; gm = fmax(gm, fabs(sm->g[k])); xn += sm->x[k] * sm->x[k]; }
  sm->gmax = gm; sm->x_norm = sqrt(xn);
}

// the whole LM loop of one pose refinement, executed by one CTA of 256 threads
__device__ __forceinline__ void pose_refine_cta(int n, const double2* __restrict__ uv, const double* __restrict__ X, int model,
                                                const double* __restrict__ intr, const mm_ba_options& O, double* __restrict__ pose_io,
                                                mm_ba_summary* __restrict__ S) {
  __shared__ PoseLM sm;
  __shared__ double wred[8][28];
  LossParams L; L.type = O.loss_type; L.b = O.loss_scale * O.loss_scale; L.c = 1.0 / L.b;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 6; ++k) sm.x[k] = pose_io[k];
    sm.radius = O.initial_trust_region_radius; sm.decrease_factor = 2.0; sm.iter = 0; sm.n_invalid = 0; sm.done = 0;
    S->num_residuals = 2 * (int64_t)n; S->final_cost = INFINITY; S->termination = MM_TERM_NO_CONVERGENCE;
    S->num_iterations = 0; S->num_successful_steps = 0; S->num_unsuccessful_steps = 0;
  }
  __syncthreads();
  pose_pass<true>(n, uv, X, model, intr, sm.x, &sm, L, wred);
  if (threadIdx.x == 0) {
    pose_take_linearization(&sm);
    // Jacobi scaling, estimated once from the column norms of the first Jacobian (diagonal of J'J)
    const int diag[6] = { 0, 6, 11, 15, 18, 20 };
    for (int k = 0; k < 6; ++k) sm.scale[k] = O.jacobi_scaling ? 1.0 / (1.0 + sqrt(sm.A[diag[k]])) : 1.0;
    S->initial_cost = sm.cost;
    pose_trace_push(S, &sm, 1);
    sm.abs_gtol = O.gradient_tolerance * fmax(sm.gmax, 1e-12);
    if (!isfinite(sm.cost)) { S->termination = MM_TERM_NUMERICAL_FAILURE; sm.done = 1; }
    else if (sm.gmax <= sm.abs_gtol) { S->termination = MM_TERM_GRADIENT_TOLERANCE; sm.done = 1; }
  }
  __syncthreads();
  // The loop flags live in shared memory and are written by thread 0 only.  Every thread copies a flag into a register right
  // after the barrier that publishes it, and a second barrier keeps thread 0 from rewriting it before all have read it.
  for (;;) {
    const int done_top = sm.done;
    __syncthreads();
    if (done_top) break;
    // ---- step: (S A S + D^2) y = S g ; candidate x2 = x - S y
    if (threadIdx.x == 0) {
      if (sm.iter >= O.max_num_iterations) { S->termination = MM_TERM_NO_CONVERGENCE; sm.done = 1; }
      else {
        sm.iter++;
        double Mx[6][6], gs[6];
        int k = 0;
        for (int p = 0; p < 6; ++p) for (int q = p; q < 6; ++q, ++k) { Mx[p][q] = Mx[q][p] = sm.A[k] * sm.scale[p] * sm.scale[q]; }
        for (int p = 0; p < 6; ++p) {
          gs[p] = sm.g[p] * sm.scale[p];
          sm.D[p] = fmin(fmax(Mx[p][p], O.min_lm_diagonal), O.max_lm_diagonal) / sm.radius;
          Mx[p][p] += sm.D[p];
        }
        bool fail = false;          // Cholesky M = L L'
        for (int j = 0; j < 6 && !fail; ++j) {
          double d = Mx[j][j];
          for (int c = 0; c < j; ++c) d -= Mx[j][c] * Mx[j][c];
          if (!(d > 0.0) || !isfinite(d)) { fail = true; break; }
          const double l = sqrt(d); Mx[j][j] = l;
          for (int i = j + 1; i < 6; ++i) { double v = Mx[i][j]; for (int c = 0; c < j; ++c) v -= Mx[i][c] * Mx[j][c]; Mx[i][j] = v / l; }
        }
        double y[6];
        if (!fail) {
          for (int i = 0; i < 6; ++i) { double v = gs[i]; for (int c = 0; c < i; ++c) v -= Mx[i][c] * y[c]; y[i] = v / Mx[i][i]; }
          for (int i = 5; i >= 0; --i) { double v = y[i]; for (int c = i + 1; c < 6; ++c) v -= Mx[c][i] * y[c]; y[i] = v / Mx[i][i]; }
        } else { for (int i = 0; i < 6; ++i) y[i] = 0.0; }
        double sn = 0.0, mc = 0.0;
        for (int p = 0; p < 6; ++p) {
          const double e = -y[p] * sm.scale[p];
          sm.x2[p] = sm.x[p] + e; sn += e * e; mc += 0.5 * y[p] * (gs[p] + sm.D[p] * y[p]);
        }
        sm.step_norm = sqrt(sn); sm.mcc = mc; sm.want = fail ? 0 : 1;
      }
    }
    __syncthreads();
    const int done_step = sm.done;
    __syncthreads();
    if (done_step) break;
    pose_pass<false>(n, uv, X, model, intr, sm.x2, &sm, L, wred);
    // ---- accept / reject (thread 0), then re-linearize if accepted
    if (threadIdx.x == 0) {
      const double new_cost = sm.acc[0];
      const bool valid = sm.want && isfinite(sm.step_norm) && isfinite(sm.mcc) && isfinite(new_cost) && !(sm.mcc < 0.0);
      int successful = 0; double rel_dec = 0.0;
      sm.want = 0;
      if (!valid) {
        if (++sm.n_invalid >= O.max_num_consecutive_invalid_steps) { S->termination = MM_TERM_NUMERICAL_FAILURE; sm.done = 1; }
      } else {
        sm.n_invalid = 0;
        const double cost_change = sm.cost - new_cost;
        if (sm.step_norm <= O.parameter_tolerance * (sm.x_norm + O.parameter_tolerance)) { S->termination = MM_TERM_PARAMETER_TOLERANCE; sm.done = 1; }
        else if (fabs(cost_change) < O.function_tolerance * sm.cost) { S->termination = MM_TERM_FUNCTION_TOLERANCE; sm.done = 1; }
        else { rel_dec = cost_change / sm.mcc; successful = rel_dec > O.min_relative_decrease; }
      }
      if (!sm.done) {
        if (successful) {
          S->num_successful_steps++;
          sm.radius = fmin(O.max_trust_region_radius, sm.radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rel_dec - 1.0, 3)));
          sm.decrease_factor = 2.0;
          for (int k = 0; k < 6; ++k) { sm.xprev[k] = sm.x[k]; sm.x[k] = sm.x2[k]; }
          sm.want = 1;                                    // linearize at the new iterate
        } else {
          S->num_unsuccessful_steps++;
          sm.radius = sm.radius / sm.decrease_factor; sm.decrease_factor *= 2.0;
          if (sm.radius < O.min_trust_region_radius) { S->termination = MM_TERM_PARAMETER_TOLERANCE; sm.done = 1; }
          else pose_trace_push(S, &sm, 0);
        }
      }
    }
    __syncthreads();
    const int want = sm.want;                             // (uniform: every thread reads the same published value)
    __syncthreads();
    if (want) {
      pose_pass<true>(n, uv, X, model, intr, sm.x, &sm, L, wred);
      if (threadIdx.x == 0) {
        pose_take_linearization(&sm);
        sm.want = 0;
        if (sm.gmax <= sm.abs_gtol) {                     // Ceres 1.8 tests the gradient before committing the iterate
          for (int k = 0; k < 6; ++k) sm.x[k] = sm.xprev[k];
          S->termination = MM_TERM_GRADIENT_TOLERANCE; sm.done = 1;
        } else if (sm.radius < O.min_trust_region_radius) { S->termination = MM_TERM_PARAMETER_TOLERANCE; sm.done = 1; }
        else pose_trace_push(S, &sm, 1);
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) for (int k = 0; k < 6; ++k) pose_io[k] = sm.x[k];
}

__global__ void __launch_bounds__(256, 1) k_pose_refine(int n, const double2* __restrict__ uv, const double* __restrict__ X, int model,
                                                         const double* __restrict__ intr, mm_ba_options O, double* __restrict__ pose_io,
                                                         mm_ba_summary* __restrict__ S) {
  pose_refine_cta(n, uv, X, model, intr, O, pose_io, S);
}

// A batch of independent pose refinements (candidate poses of one image, or the images of a re-localisation sweep) in ONE launch:
// CTA b runs the LM loop of problem b.  off[b] .. off[b+1] are its 2D-3D pairs in the concatenated arrays; intr [B][9], pose [B][6].
__global__ void __launch_bounds__(256, 1) k_pose_refine_batch(const int64_t* __restrict__ off, const double2* __restrict__ uv, const double* __restrict__ X,
                                                               const int* __restrict__ model, const double* __restrict__ intr, mm_ba_options O,
                                                               double* __restrict__ pose_io, mm_ba_summary* __restrict__ S) {
  const int b = blockIdx.x;
  const int64_t o0 = off[b];
  pose_refine_cta((int)(off[b + 1] - o0), uv + o0, X + 3 * o0, model[b], intr + MM_INTR_STRIDE * (size_t)b, O, pose_io + 6 * (size_t)b, S + b);
}

}  // namespace mm
